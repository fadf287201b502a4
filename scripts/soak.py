"""Soak of the pipelined device-resident path: `rounds` x `depth` steps over `depth` pipeline slots (replayed step graphs,
merged phases, persistent last layer), every slot on its own fixed batch; after every round each slot's container size,
container bytes and reconstruction must equal those of its first step (the pipeline is deterministic), which also
rules out time-outs of the tensor kernels' barriers. Usage: python scripts/soak.py [depth] [rounds] [math]"""
import ctypes
import os
import sys
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from autoencoder_based_image_compression_b200 import _native, synthetic                     # noqa: E402
from autoencoder_based_image_compression_b200 import codec as native_codec                  # noqa: E402
from autoencoder_based_image_compression_b200 import weights as wts                         # noqa: E402
import bench                                                                                 # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    math = sys.argv[3] if len(sys.argv) > 3 else 'mixed'
    lib = _native.lib()
    torch.cuda.set_device(0)
    (n, h, w) = (24, 512, 768)
    (table, map_mean) = bench.load_tables()
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), table, map_mean)
    native_params = params.native()
    bound = int(lib.eae_container_bound(n, h, w, params.truncated_unary_length))
    weights = wts.random_init(0, False)
    codecs = [native_codec.Codec(weights, False, device=0, math=math, own_stream=True) for _ in range(depth)]
    for c in codecs:
        c.set_coder_lanes(1)
    rng = numpy.random.default_rng(1)
    base = synthetic.synthetic_luma(rng, n, h, w)
    d_img = [torch.from_numpy(numpy.roll(base, shift=(k, 7*k, 13*k), axis=(0, 1, 2)).copy()).cuda() for k in range(depth)]
    d_rec = [torch.zeros((n, h, w), dtype=torch.uint8, device='cuda') for _ in range(depth)]
    d_cont = [torch.zeros(bound, dtype=torch.uint8, device='cuda') for _ in range(depth)]
    d_tot = [torch.zeros(1, dtype=torch.int64, device='cuda') for _ in range(depth)]
    d_stats = [torch.zeros(ctypes.sizeof(_native.BatchStats), dtype=torch.uint8, device='cuda') for _ in range(depth)]

    def step(k):
        c = codecs[k]
        _native.check(lib.eae_compress_dev(c.handle, ctypes.byref(native_params), d_img[k].data_ptr(), n, h, w,
                                           d_cont[k].data_ptr(), bound, d_tot[k].data_ptr(), d_stats[k].data_ptr(), c.stream))
        _native.check(lib.eae_decompress_dev(c.handle, ctypes.byref(native_params), d_cont[k].data_ptr(), bound, n, h, w,
                                             d_rec[k].data_ptr(), c.stream))

    def sync():
        for c in codecs:
            _native.check(lib.eae_stream_synchronize(c.stream))

    for k in range(depth):
        step(k)
    sync()
    want = [(int(d_tot[k][0]), d_cont[k][:int(d_tot[k][0])].clone(), d_rec[k].clone(), d_stats[k].clone()) for k in range(depth)]
    t0 = time.time()
    for r in range(rounds):
        for k in range(depth):
            d_rec[k].zero_()
        torch.cuda.synchronize()
        for k in range(depth):
            step(k)
        sync()
        for k in range(depth):
            size = int(d_tot[k][0])
            assert size == want[k][0], (r, k, size, want[k][0])
            assert torch.equal(d_cont[k][:size], want[k][1]), (r, k, 'container')
            assert torch.equal(d_rec[k], want[k][2]), (r, k, 'reconstruction')
            assert torch.equal(d_stats[k], want[k][3]), (r, k, 'statistics')
    print('soak ok: {} steps of {} images on {} slots ({}), every container / reconstruction / statistics block equal to the '
          'first one; {:.1f} s'.format(rounds*depth, n, depth, math, time.time() - t0))


if __name__ == '__main__':
    main()

"""Would CUDA graphs help the pipelined codec? Captures one whole step (eae_compress_dev + eae_decompress_dev, ~25
launches) per pipeline slot into a CUDA graph (torch.cuda.CUDAGraph on the slot's stream; the library's launches are
plain stream work, so stream capture records them) and replays the graphs round-robin, against the same steps enqueued
launch by launch. Each slot keeps one fixed input batch (a graph freezes its pointers). Device time, CUDA events."""
import ctypes
import os
import sys

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
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    lib = _native.lib()
    torch.cuda.set_device(0)
    (n, h, w) = (24, 512, 768)
    (table, map_mean) = bench.load_tables()
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), table, map_mean)
    native_params = params.native()
    bound = int(lib.eae_container_bound(n, h, w, params.truncated_unary_length))
    weights = wts.random_init(0, False)
    codecs = [native_codec.Codec(weights, False, device=0, math='mixed', own_stream=True) for _ in range(depth)]
    for c in codecs:
        c.set_coder_lanes(1)
    streams = [torch.cuda.ExternalStream(c.stream.value) for c in codecs]
    rng = numpy.random.default_rng(1)
    base = synthetic.synthetic_luma(rng, n, h, w)
    d_img = [torch.from_numpy(numpy.roll(base, shift=(k, 7*k, 13*k), axis=(0, 1, 2)).copy()).cuda() for k in range(depth)]
    d_rec = [torch.empty((n, h, w), dtype=torch.uint8, device='cuda') for _ in range(depth)]
    d_cont = [torch.empty(bound, dtype=torch.uint8, device='cuda') for _ in range(depth)]
    d_tot = [torch.zeros(1, dtype=torch.int64, device='cuda') for _ in range(depth)]
    d_stats = [torch.zeros(ctypes.sizeof(_native.BatchStats), dtype=torch.uint8, device='cuda') for _ in range(depth)]

    def full(k):
        c = codecs[k]
        _native.check(lib.eae_compress_dev(c.handle, ctypes.byref(native_params), d_img[k].data_ptr(), n, h, w,
                                           d_cont[k].data_ptr(), bound, d_tot[k].data_ptr(), d_stats[k].data_ptr(), c.stream))
        _native.check(lib.eae_decompress_dev(c.handle, ctypes.byref(native_params), d_cont[k].data_ptr(), bound, n, h, w,
                                             d_rec[k].data_ptr(), c.stream))

    for _ in range(2):
        for k in range(depth):
            full(k)
    torch.cuda.synchronize()
    want = [r.clone() for r in d_rec]
    graphs = []
    for k in range(depth):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(streams[k]):
            g.capture_begin()
            full(k)
            g.capture_end()
        graphs.append(g)
    torch.cuda.synchronize()

    def timed(fn):
        for i in range(2*depth):
            fn(i % depth)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(depth)]
        e0.record(streams[0])
        for i in range(steps):
            fn(i % depth)
        for k in range(depth):
            ends[k].record(streams[k])
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e) for e in ends)/steps

    def replay(k):
        with torch.cuda.stream(streams[k]):
            graphs[k].replay()

    for rep in range(2):
        ms = timed(full)
        print('launch by launch, {} slots: {:.3f} ms per 24-image step ({:.0f} images/s)'.format(depth, ms, n/ms*1e3))
        ms = timed(replay)
        print('one graph per step, {} slots: {:.3f} ms per 24-image step ({:.0f} images/s)'.format(depth, ms, n/ms*1e3))
    assert all(torch.equal(a, b) for (a, b) in zip(want, d_rec))
    print('reconstructions of the replayed graphs equal the launch-by-launch ones')


if __name__ == '__main__':
    main()

"""What does a running arithmetic-coder kernel cost the tensor kernels that share the GPU with it? The transforms of
24-image batches run back to back on one stream while K other streams run the decoder (or encoder) kernel in a loop on
a prepared batch of 3072 streams; the transforms' time per step is printed for K = 0, 1, 2, 4.
Run with EAE_CODER_LANES=1 (the pipeline's setting)."""
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
    which = sys.argv[1] if len(sys.argv) > 1 else 'decode'
    lib = _native.lib()
    (n, h, w) = (24, 512, 768)
    (table, _) = bench.load_tables()
    dev = torch.device('cuda', 0)
    rng = numpy.random.default_rng(5)
    scales = 2.0*numpy.exp(rng.normal(0., 0.7, size=128))
    lat = numpy.round(rng.laplace(0., 1., size=(n, 128, 1536))*scales.reshape(1, 128, 1)).astype(numpy.int16)
    (n_streams, size, L) = (n*128, 1536, table.shape[1])
    slot = lib.eae_coder_slot_bytes(size, L)
    d_planar = torch.from_numpy(lat.reshape(n_streams, size)).to(dev)
    d_table = torch.from_numpy(numpy.ascontiguousarray(table)).to(dev)
    d_bac = torch.zeros(n_streams*slot, dtype=torch.uint8, device=dev)
    d_byp = torch.zeros(n_streams*slot, dtype=torch.uint8, device=dev)
    d_bb = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    d_rb = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    d_err = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    d_off = (torch.arange(n_streams, dtype=torch.int64, device=dev)*slot).contiguous()
    st0 = torch.cuda.current_stream().cuda_stream
    _native.check(lib.eae_encode_streams_dev(d_planar.data_ptr(), n_streams, size, d_table.data_ptr(), 128, L, None,
                                             d_bac.data_ptr(), d_byp.data_ptr(), slot, d_bb.data_ptr(), d_rb.data_ptr(),
                                             d_err.data_ptr(), st0))
    torch.cuda.synchronize()
    assert int(d_err.max()) == 0
    bg_streams = [torch.cuda.Stream() for _ in range(4)]
    bg_out = [torch.empty((n_streams, size), dtype=torch.int16, device=dev) for _ in range(4)]
    bg_bac = [torch.zeros(n_streams*slot, dtype=torch.uint8, device=dev) for _ in range(4)]
    bg_byp = [torch.zeros(n_streams*slot, dtype=torch.uint8, device=dev) for _ in range(4)]
    bg_bits = [torch.zeros((3, n_streams), dtype=torch.int32, device=dev) for _ in range(4)]

    def background(k):
        s = bg_streams[k].cuda_stream
        if which == 'decode':
            _native.check(lib.eae_decode_streams_dev(bg_out[k].data_ptr(), n_streams, size, d_table.data_ptr(), 128, L, None,
                                                     d_bac.data_ptr(), d_off.data_ptr(), d_bb.data_ptr(), d_byp.data_ptr(),
                                                     d_off.data_ptr(), d_rb.data_ptr(), bg_bits[k][2].data_ptr(), s))
        else:
            _native.check(lib.eae_encode_streams_dev(d_planar.data_ptr(), n_streams, size, d_table.data_ptr(), 128, L, None,
                                                     bg_bac[k].data_ptr(), bg_byp[k].data_ptr(), slot, bg_bits[k][0].data_ptr(),
                                                     bg_bits[k][1].data_ptr(), bg_bits[k][2].data_ptr(), s))

    codec = native_codec.Codec(wts.random_init(0, False), False, device=0, math='mixed', own_stream=True)
    img = synthetic.synthetic_luma(rng, n, h, w)
    d_img = torch.from_numpy(img).to(dev)
    d_y = torch.empty((n, h//16, w//16, 128), dtype=torch.float32, device=dev)
    d_rec = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    ev = [ctypes.c_void_p(), ctypes.c_void_p()]
    for e in ev:
        _native.check(lib.eae_event_create(ctypes.byref(e)))

    def transforms():
        _native.check(lib.eae_encode_dev(codec.handle, d_img.data_ptr(), n, h, w, d_y.data_ptr(), codec.stream))
        _native.check(lib.eae_decode_dev(codec.handle, d_y.data_ptr(), n, h, w, d_rec.data_ptr(), codec.stream))

    for _ in range(3):
        transforms()
    torch.cuda.synchronize()
    # duration of one background kernel alone
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(bg_streams[0]):
        background(0)
        t0.record()
        background(0)
        t1.record()
    torch.cuda.synchronize()
    print('{} launch alone: {:.3f} ms'.format(which, t0.elapsed_time(t1)))
    steps = int(os.environ.get('PROBE_STEPS', '12'))
    for K in [int(k) for k in os.environ.get('PROBE_KS', '0,1,2,4').split(',')]:
        print('--- K = {}'.format(K), file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        for r in range(int(os.environ.get('PROBE_BG', '14'))):      # all of the background first: it must outlast the transforms
            for k in range(K):
                background(k)
        _native.check(lib.eae_event_record(ev[0], codec.stream))
        for _ in range(steps):
            transforms()
        _native.check(lib.eae_event_record(ev[1], codec.stream))
        ms = ctypes.c_float(0.)
        _native.check(lib.eae_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms)))
        torch.cuda.synchronize()
        print('{} x {} concurrent {} kernels of 96 warps: transforms {:.3f} ms per step'.format(K, 1, which, ms.value/steps))


if __name__ == '__main__':
    main()

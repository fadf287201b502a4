"""How much of the pipelined step is the transforms alone? Runs `steps` batches of 24 images round-robin over `depth`
CUDA streams (one codec each) three ways: transforms only (eae_encode_dev + eae_decode_dev), the full codec
(eae_compress_dev + eae_decompress_dev), and the coder's share by difference. Device time, CUDA events."""
import ctypes
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from autoencoder_based_image_compression_b200 import _native, synthetic                     # noqa: E402
from autoencoder_based_image_compression_b200 import codec as native_codec                  # noqa: E402
from autoencoder_based_image_compression_b200 import weights as wts                         # noqa: E402
import bench                                                                                 # noqa: E402


def main():
    math = sys.argv[1] if len(sys.argv) > 1 else 'mixed'
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 48
    lib = _native.lib()
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
    rotate = 15
    host = _native.pinned_empty((rotate, n, h, w), numpy.uint8)
    base = synthetic.synthetic_luma(rng, n, h, w)
    for r in range(rotate):
        host[r] = numpy.roll(base, shift=(r, 7*r, 13*r), axis=(0, 1, 2))
    d_img = lib.eae_device_alloc(rotate*n*h*w)
    _native.check(lib.eae_memcpy_h2d(d_img, _native.ptr(host), rotate*n*h*w, None))
    d_y = [lib.eae_device_alloc(n*(h//16)*(w//16)*128*4) for _ in range(depth)]
    d_rec = [lib.eae_device_alloc(n*h*w) for _ in range(depth)]
    d_cont = [lib.eae_device_alloc(bound) for _ in range(depth)]
    d_tot = [lib.eae_device_alloc(8) for _ in range(depth)]
    d_stats = [lib.eae_device_alloc(ctypes.sizeof(_native.BatchStats)) for _ in range(depth)]
    ev0 = ctypes.c_void_p()
    _native.check(lib.eae_event_create(ctypes.byref(ev0)))
    ends = []
    for _ in range(depth):
        e = ctypes.c_void_p()
        _native.check(lib.eae_event_create(ctypes.byref(e)))
        ends.append(e)

    def transforms(i, k):
        c = codecs[k]
        img = d_img + (i % rotate)*n*h*w
        _native.check(lib.eae_encode_dev(c.handle, img, n, h, w, d_y[k], c.stream))
        _native.check(lib.eae_decode_dev(c.handle, d_y[k], n, h, w, d_rec[k], c.stream))

    def full(i, k):
        c = codecs[k]
        img = d_img + (i % rotate)*n*h*w
        _native.check(lib.eae_compress_dev(c.handle, ctypes.byref(native_params), img, n, h, w, d_cont[k], bound,
                                           d_tot[k], d_stats[k], c.stream))
        _native.check(lib.eae_decompress_dev(c.handle, ctypes.byref(native_params), d_cont[k], bound, n, h, w, d_rec[k], c.stream))

    def sync():
        for c in codecs:
            _native.check(lib.eae_stream_synchronize(c.stream))
        _native.check(lib.eae_stream_synchronize(None))

    def timed(fn, active):
        for i in range(2*depth):
            fn(i, i % active)
        sync()
        _native.check(lib.eae_event_record(ev0, codecs[0].stream))
        for i in range(steps):
            fn(i, i % active)
        for k in range(active):
            _native.check(lib.eae_event_record(ends[k], codecs[k].stream))
        sync()
        worst = 0.
        ms = ctypes.c_float(0.)
        for k in range(active):
            _native.check(lib.eae_event_elapsed_ms(ev0, ends[k], ctypes.byref(ms)))
            worst = max(worst, ms.value)
        return worst/steps

    for (name, fn) in (('transforms only', transforms), ('full codec', full)):
        for active in sorted({1, 2, depth}):
            ms = timed(fn, active)
            print('{:16s} math {} streams {}: {:.3f} ms per 24-image step ({:.0f} images/s)'.format(name, math, active, ms, n/ms*1e3))


if __name__ == '__main__':
    main()

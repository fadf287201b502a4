"""Small compress / decompress runs for compute-sanitizer (memcheck, racecheck, synccheck, initcheck):

    compute-sanitizer --tool memcheck python scripts/sanitize.py

Both arithmetic modes of the tensor path, both variants of the architecture, every packing of the coder kernels, latent
grids that are and are not multiples of a tile, the device-resident entry points and the histogram kernels."""
import ctypes
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autoencoder_based_image_compression_b200 import _native, synthetic  # noqa: E402
from autoencoder_based_image_compression_b200 import codec as native_codec  # noqa: E402
from autoencoder_based_image_compression_b200 import weights as wts  # noqa: E402

rng = numpy.random.default_rng(0)
(table, map_mean) = bench.load_tables()
params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), table, map_mean)
lib = _native.lib()
for (n, h, w) in ((2, 64, 96), (1, 48, 80), (1, 128, 192)):
    for (math, learned, lanes) in (('mixed', False, 1), ('tf32x3', False, 0), ('mixed', True, 32), ('fp32', False, 2)):
        c = native_codec.Codec(wts.random_init(0, learned), learned, device=0, math=math)
        c.set_coder_lanes(lanes)
        img = synthetic.synthetic_luma(rng, n, h, w)
        for _ in range(2):
            blob = c.compress(img, params)
            rec = c.decompress(blob, params)
        idx = c.last_indices(n, h, w)
        print(n, h, w, math, learned, lanes, blob.size, float(rec.mean()), int(idx.min()), int(idx.max()))
# device-resident entry points (step graphs from the third use on) and the status poll
(n, h, w) = (2, 64, 96)
c = native_codec.Codec(wts.random_init(0, False), False, device=0, math='mixed', own_stream=True)
bound = int(lib.eae_container_bound(n, h, w, 10))
img = synthetic.synthetic_luma(rng, n, h, w)
d_img = lib.eae_device_alloc(n*h*w)
_native.check(lib.eae_memcpy_h2d(d_img, _native.ptr(img), n*h*w, None))
_native.check(lib.eae_stream_synchronize(None))
(d_cont, d_rec, d_tot) = (lib.eae_device_alloc(bound), lib.eae_device_alloc(n*h*w), lib.eae_device_alloc(8))
prm = params.native()
for _ in range(4):
    _native.check(lib.eae_compress_dev(c.handle, ctypes.byref(prm), d_img, n, h, w, d_cont, bound, d_tot, None, c.stream))
    _native.check(lib.eae_decompress_dev(c.handle, ctypes.byref(prm), d_cont, bound, n, h, w, d_rec, c.stream))
print('status', c.poll_status())
# histograms of planar streams
idx = c.last_indices(n, h, w)
d_idx = lib.eae_device_alloc(idx.size*2)
_native.check(lib.eae_memcpy_h2d(d_idx, _native.ptr(idx), idx.size*2, None))
(d_mn, d_mx) = (lib.eae_device_alloc(4*n*128), lib.eae_device_alloc(4*n*128))
cap = int(idx.max()) - int(idx.min()) + 1
d_hist = lib.eae_device_alloc(8*n*128*cap)
_native.check(lib.eae_histogram_streams_dev(d_idx, n, (h//16)*(w//16), 128, 1, d_mn, d_mx, None, None, 0, None))
_native.check(lib.eae_histogram_streams_dev(d_idx, n, (h//16)*(w//16), 128, 1, d_mn, None, None, d_hist, cap, None))
_native.check(lib.eae_stream_synchronize(None))
print('done')

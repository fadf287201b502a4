import sys, time, numpy
sys.path.insert(0, '/root/repo')
from autoencoder_based_image_compression_b200 import codec as native_codec, synthetic, weights as wts
import bench
table, mean = bench.load_tables()
w = wts.random_init(0, False)
rng = numpy.random.default_rng(9)
lum = synthetic.synthetic_luma(rng, 2, 2160, 3840)
params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), table, mean)
codec = native_codec.Codec(w, False, math='tf32x3')
t0 = time.time(); blob = codec.compress(lum, params); t1 = time.time()
idx = codec.last_indices(2, 2160, 3840).copy()
rec = codec.decompress(blob, params); t2 = time.time()
idx2 = codec.last_indices(2, 2160, 3840)
print('blob', blob.size, 'bpp', blob.size*8/lum.size, 'compress s', t1-t0, 'decompress s', t2-t1, 'indices equal', numpy.array_equal(idx, idx2), rec.shape, rec.min(), rec.max())

"""Round trip of a large batch (BASELINE config 4 per-GPU share): compress -> decompress, indices identical."""
import sys, time, numpy
sys.path.insert(0, '.')
from autoencoder_based_image_compression_b200 import codec as native_codec, synthetic, weights as wts
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
table, mean = bench.load_tables()
w = wts.random_init(0, False)
rng = numpy.random.default_rng(9)
base = synthetic.synthetic_luma(rng, 32, 512, 768)
lum = numpy.concatenate([numpy.roll(base, shift=(7*k, 13*k), axis=(1, 2)) for k in range(n//32)], axis=0)
params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), table, mean)
codec = native_codec.Codec(w, False, math='tf32x3')
codec.compress(lum[:32], params)
t0 = time.time(); blob = codec.compress(lum, params); t1 = time.time()
idx = codec.last_indices(n, 512, 768).copy()
rec = codec.decompress(blob, params); t2 = time.time()
print('n', n, 'blob MB', blob.size/1e6, 'bpp', blob.size*8/lum.size, 'compress', n/(t1-t0), 'img/s decompress', n/(t2-t1), 'img/s',
      'indices equal', numpy.array_equal(idx, codec.last_indices(n, 512, 768)), 'rec range', rec.min(), rec.max())

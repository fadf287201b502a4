"""Index agreement of the GPU encoder with the CPU oracle, and how far from a bin boundary the oracle's latent
is where they differ (a rounding-order mismatch sits within ~1e-5 of a boundary; a bug does not).

    EAE_UMMA_VERSION=3|4 python scripts/debug_agree.py
"""
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoencoder_based_image_compression_b200 import codec as native_codec  # noqa: E402
from autoencoder_based_image_compression_b200 import synthetic, weights as wts  # noqa: E402
from oracle import transforms as oracle_transforms  # noqa: E402

for (learned, seed, shape) in [(True, 3, (3, 128, 192)), (False, 3, (3, 128, 192)), (True, 4, (2, 256, 256)), (False, 5, (2, 64, 96))]:
    rng = numpy.random.default_rng(seed)
    w = wts.random_init(0, learned)
    (n, h, wd) = shape
    lum = synthetic.synthetic_luma(rng, n, h, wd)
    codec = native_codec.Codec(w, learned, math='tf32x3')
    y = codec.encode(lum)
    y_ref = oracle_transforms.encoder(lum[..., None].astype(numpy.float32), w, learned)
    import torch
    y64 = oracle_transforms.encoder(lum[..., None].astype(numpy.float64), w, learned, dtype=torch.float64)
    k = numpy.rint(y).astype(numpy.int64)
    k_ref = numpy.rint(y_ref).astype(numpy.int64)
    bad = k != k_ref
    dist = numpy.abs(numpy.abs(y_ref - numpy.floor(y_ref)) - 0.5)
    print('learned', learned, 'shape', shape, 'agree', 1. - bad.mean(), 'mismatches', int(bad.sum()),
          'max |y - y_ref|', float(numpy.abs(y - y_ref).max()), 'rel', float(numpy.abs(y - y_ref).max()/numpy.abs(y_ref).max()),
          'max boundary distance at mismatches', float(dist[bad].max()) if bad.any() else 0.)
    if y64 is not None:
        k64 = numpy.rint(y64).astype(numpy.int64)
        print('   vs float64 oracle: gpu mismatches', int((k != k64).sum()), 'fp32 oracle mismatches', int((k_ref != k64).sum()),
              'max err gpu', float(numpy.abs(y - y64).max()), 'max err fp32 oracle', float(numpy.abs(y_ref - y64).max()))

"""Diagnostic for the tcgen05 path: runs the transforms with each layer kind switched to tensor cores in
turn (EAE_UMMA_LAYERS bit mask: 1 convs, 2 transposed convs, 4 GDN/IGDN, 8 first/last layer) and prints
the deviation from the fp32 CUDA-core path. Usage: python scripts/debug_umma.py [h w n]"""
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoencoder_based_image_compression_b200 import codec as native_codec  # noqa: E402
from autoencoder_based_image_compression_b200 import synthetic  # noqa: E402
from autoencoder_based_image_compression_b200 import weights as wts  # noqa: E402


def main():
    (h, w, n) = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 96, 2)
    learned = False
    weights = wts.random_init(0, learned)
    lum = synthetic.synthetic_luma(numpy.random.default_rng(0), n, h, w, smooth=False)[..., None]
    ref = native_codec.Codec(weights, learned, math='fp32')
    y_ref = ref.encode(lum)
    q = numpy.round(y_ref)
    r_ref = ref.decode(q)
    print('reference: |y| max {:.4f}, rec range {}..{}'.format(numpy.abs(y_ref).max(), r_ref.min(), r_ref.max()), flush=True)
    for math in ('tf32x3', 'tf32'):
        for mask in (1, 2, 4, 8, 15):
            os.environ['EAE_UMMA_LAYERS'] = str(mask)
            try:
                t0 = time.time()
                c = native_codec.Codec(weights, learned, math=math)
                y = c.encode(lum)
                r = c.decode(q)
                dy = numpy.abs(y - y_ref)
                idx_bad = (numpy.round(y) != numpy.round(y_ref)).mean()
                print('{:7s} mask {:2d}: encode max|dy| {:.3e} (rel {:.3e}), index mismatch {:.3e}, '
                      'decode pixel mismatch {:.3e} max {}  [{:.2f}s]'.format(
                          math, mask, dy.max(), dy.max()/numpy.abs(y_ref).max(), idx_bad,
                          (r != r_ref).mean(), int(numpy.abs(r.astype(int) - r_ref.astype(int)).max()),
                          time.time() - t0), flush=True)
                c.close()
            except Exception as err:    # keep going: the point is to see which layer kind breaks
                print('{:7s} mask {:2d}: FAILED {}'.format(math, mask, err), flush=True)


if __name__ == '__main__':
    main()

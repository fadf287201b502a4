"""Synthetic luminance images with the shape / dtype contract of the Kodak set
(kodak_tensorflow/datasets/kodak/kodak.py:52-54,70-75: uint8, 512 x 768, values nominally 16..235)."""
import numpy


def synthetic_luma(rng, n, h, w, smooth=True):
    """uint8 [n, h, w] in 16..235. ``smooth``: low-pass filtered noise (Kodak-like statistics);
    otherwise i.i.d. uniform noise (the worst case for the rate)."""
    if not smooth:
        return rng.integers(16, 236, size=(n, h, w), dtype=numpy.uint8)
    x = rng.random((n, h, w))
    for axis in (1, 2):
        for _ in range(3):   # three box filters of width 7 ~ Gaussian blur with sigma ~ 3.5
            acc = numpy.zeros_like(x)
            for shift in range(-3, 4):
                acc += numpy.roll(x, shift, axis=axis)
            x = acc/7.
    x -= x.min(axis=(1, 2), keepdims=True)
    x /= x.max(axis=(1, 2), keepdims=True) + 1e-12
    return (16. + 219.*x).round().astype(numpy.uint8)

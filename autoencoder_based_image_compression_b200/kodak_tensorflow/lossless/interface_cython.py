"""ctypes replacement of the reference's Cython module ``lossless.interface_cython``
(kodak_tensorflow/lossless/interface_cython.pyx:13-59): same function name, argument checks and
return value; the coder runs on the GPU through ``eae_compress_lossless``."""
import ctypes

import numpy

from autoencoder_based_image_compression_b200 import _native


def compress_lossless_flattened_map(ref_map_int16, probabilities):
    """Compresses without loss a flattened map of signed integers and decodes it back.

    Returns ``(rec_map_int16, nb_bits)`` like interface_cython.pyx:59. Raises what Cython raises for
    the same misuse: ``ValueError`` for a wrong dtype / number of dimensions (buffer mismatch),
    ``OverflowError`` when ``probabilities.size`` does not fit ``uint8`` (:50-52), ``RuntimeError``
    ("Error of type N ...") for the coder's error codes (``except +``).
    """
    if not isinstance(ref_map_int16, numpy.ndarray) or not isinstance(probabilities, numpy.ndarray):
        raise TypeError('Argument has incorrect type (expected numpy.ndarray).')
    if ref_map_int16.dtype != numpy.int16 or probabilities.dtype != numpy.float64:
        raise ValueError('Buffer dtype mismatch.')
    if ref_map_int16.ndim != 1 or probabilities.ndim != 1:
        raise ValueError('Buffer has wrong number of dimensions (expected 1).')
    if probabilities.size > 255:
        raise OverflowError('value too large to convert to numpy.uint8_t')
    if probabilities.size == 0 or ref_map_int16.size == 0:
        raise IndexError('Out of bounds on buffer access (axis 0)')
    ref = numpy.ascontiguousarray(ref_map_int16)
    probs = numpy.ascontiguousarray(probabilities)
    rec = numpy.zeros(ref.size, dtype=numpy.int16)
    nb_bits = ctypes.c_uint32(0)
    _native.check(_native.lib().eae_compress_lossless(ref.size, _native.ptr(ref), _native.ptr(rec), probs.size,
                                                      _native.ptr(probs), ctypes.byref(nb_bits)))
    return (rec, nb_bits.value)

"""GPU-backed ``lossless.compression`` with the reference's API
(kodak_tensorflow/lossless/compression.py:11-154): all maps of a latent are coded as independent GPU
streams in one launch instead of 127 FFI calls."""
import ctypes
import os

import numpy

from autoencoder_based_image_compression_b200 import _native
from autoencoder_based_image_compression_b200.kodak_tensorflow.tools import tools as tls

_TABLE_CACHE = {}


def _load_table(path_to_binary_probabilities, nb_maps):
    # The reference re-loads the .npy for every image (compression.py:60). The table is tiny; it is cached per path AND
    # file identity (modification time, size), so that statistics regenerated in the same process (stats.save_statistics
    # tells the user to delete the files to recompute them) are picked up as the reference would pick them up.
    st = os.stat(path_to_binary_probabilities)
    key = (path_to_binary_probabilities, st.st_mtime_ns, st.st_size)
    table = _TABLE_CACHE.get(key)
    if table is None:
        table = numpy.load(path_to_binary_probabilities)
        for stale in [k for k in _TABLE_CACHE if k[0] == path_to_binary_probabilities]:
            del _TABLE_CACHE[stale]
        _TABLE_CACHE[key] = table
    if table.ndim != 2:
        raise ValueError('`binary_probabilities.ndim` is not equal to 2.')
    if table.shape[0] != nb_maps:
        raise ValueError('`binary_probabilities.shape[0]` is not equal to `ref_int16.shape[2]`.')
    if table.shape[1] > 255:
        raise OverflowError('value too large to convert to numpy.uint8_t')
    return numpy.ascontiguousarray(table, dtype=numpy.float64)


def _skip_mask(nb_maps, idx_map_exception):
    if idx_map_exception is None or idx_map_exception < 0 or idx_map_exception >= nb_maps:
        return None
    mask = numpy.zeros(nb_maps, dtype=numpy.uint8)
    mask[idx_map_exception] = 1
    return mask


def _exception_bits(ref_int16, idx_map_exception):
    # compression.py:73-74: ceil(h*w*H(map)), the entropy of the integer symbols with bin width 1.
    (height_map, width_map, _) = ref_int16.shape
    exception = numpy.ascontiguousarray(ref_int16[:, :, idx_map_exception]).reshape((1, height_map, width_map, 1))
    (mn, mx, hist, _) = tls._histograms(exception, False)
    entropy = tls._entropy_from_counts(hist[0, :int(mx[0]) - int(mn[0]) + 1])
    return numpy.ceil(height_map*width_map*entropy).astype(numpy.uint32)


def compress_lossless_maps(ref_int16, path_to_binary_probabilities, idx_map_exception=-1):
    """compression.py:11-82. Returns ``(rec_int16 [h,w,C], nb_bits_each_map uint32 [C])``."""
    if ref_int16.dtype != numpy.int16:
        raise TypeError('`ref_int16.dtype` is not equal to `numpy.int16`.')
    (height_map, width_map, nb_maps) = ref_int16.shape
    table = _load_table(path_to_binary_probabilities, nb_maps)
    ref = numpy.ascontiguousarray(ref_int16)
    rec_int16 = numpy.zeros((height_map, width_map, nb_maps), dtype=numpy.int16)
    nb_bits_each_map = numpy.zeros(nb_maps, dtype=numpy.uint32)
    mask = _skip_mask(nb_maps, idx_map_exception)
    _native.check(_native.lib().eae_compress_lossless_maps_host(
        _native.ptr(ref), height_map, width_map, nb_maps, _native.ptr(table), table.shape[1], _native.ptr(mask),
        _native.ptr(rec_int16), _native.ptr(nb_bits_each_map), None))
    if mask is not None:
        nb_bits_each_map[idx_map_exception] = _exception_bits(ref, idx_map_exception)
    return (rec_int16, nb_bits_each_map)


def rescale_compress_lossless_maps(centered_quantized_data, bin_widths_test, path_to_binary_probabilities,
                                   idx_map_exception=-1):
    """compression.py:84-154. Returns the number of bits (Python int)."""
    if bin_widths_test.ndim != 1:
        raise ValueError('`bin_widths_test.ndim` is not equal to 1.')
    (height_map, width_map, nb_maps) = centered_quantized_data.shape
    if bin_widths_test.size != nb_maps:
        raise ValueError('`bin_widths_test.size` is not equal to `centered_quantized_data.shape[2]`.')
    table = _load_table(path_to_binary_probabilities, nb_maps)
    data = numpy.ascontiguousarray(centered_quantized_data, dtype=numpy.float32)
    bw = numpy.ascontiguousarray(bin_widths_test, dtype=numpy.float32)
    nb_bits_each_map = numpy.zeros(nb_maps, dtype=numpy.uint32)
    mask = _skip_mask(nb_maps, idx_map_exception)
    idx = numpy.empty((height_map, width_map, nb_maps), dtype=numpy.int16) if mask is not None else None
    _native.check(_native.lib().eae_rescale_compress_lossless_maps_host(
        _native.ptr(data), height_map, width_map, nb_maps, _native.ptr(bw), _native.ptr(table), table.shape[1],
        _native.ptr(mask), _native.ptr(idx), _native.ptr(nb_bits_each_map), None))
    if mask is not None:
        nb_bits_each_map[idx_map_exception] = _exception_bits(idx, idx_map_exception)
    return numpy.sum(nb_bits_each_map).item()

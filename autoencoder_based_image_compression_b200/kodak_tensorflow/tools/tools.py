"""GPU-backed versions of the numpy glue on the hot path of the reference.

Same names, argument meaning and exceptions as kodak_tensorflow/tools/tools.py of the reference
(cited per function); the arithmetic runs in ``libeae_b200.so`` kernels. Nothing here falls back to
numpy for the computation: without the CUDA library or a device every function raises.
"""
import ctypes

import numpy

from autoencoder_based_image_compression_b200 import _native


def _is_float(dtype):
    # The reference tests `numpy.issubdtype(dtype, numpy.float)` (tools.py:91, 124); `numpy.float` was
    # removed from numpy, `numpy.floating` is the same test.
    return numpy.issubdtype(dtype, numpy.floating)


def _as_float32(array):
    return numpy.ascontiguousarray(array, dtype=numpy.float32)


def cast_bt601(array_float):
    """tools.py:61-93: clip to [16, 235], round half to even, cast to uint8."""
    if not _is_float(array_float.dtype):
        raise TypeError('`array_float.dtype` is not smaller than `numpy.float` in type hierarchy.')
    data = _as_float32(array_float)
    out = numpy.empty(data.shape, dtype=numpy.uint8)
    _native.check(_native.lib().eae_cast_bt601_host(_native.ptr(data), _native.ptr(out), data.size, None))
    return out


def cast_float_to_int16(array_float):
    """tools.py:95-133: round half to even, assert |.| < 32768, cast to int16."""
    if not _is_float(array_float.dtype):
        raise TypeError('`array_float.dtype` is not smaller than `numpy.float` in type hierarchy.')
    data = _as_float32(array_float)
    out = numpy.empty(data.shape, dtype=numpy.int16)
    _native.check(_native.lib().eae_cast_float_to_int16_host(_native.ptr(data), _native.ptr(out), data.size, None))
    return out


def _histograms(idx_nhwc, per_image):
    """Per-map symbol histograms on the GPU: (min int32[J], max int32[J], counts uint64[J, cap], abs_sum)."""
    idx = numpy.ascontiguousarray(idx_nhwc, dtype=numpy.int16)
    (n, h, w, nb_maps) = idx.shape
    nb_hist = n*nb_maps if per_image else nb_maps
    mn = numpy.empty(nb_hist, dtype=numpy.int32)
    mx = numpy.empty(nb_hist, dtype=numpy.int32)
    abs_sum = numpy.empty(nb_hist, dtype=numpy.uint64)
    cap = 1024
    while True:
        hist = numpy.empty((nb_hist, cap), dtype=numpy.uint64)
        needed = ctypes.c_uint32(0)
        code = _native.lib().eae_histogram_maps_host(_native.ptr(idx), n, h, w, nb_maps, int(per_image),
                                                     _native.ptr(mn), _native.ptr(mx), _native.ptr(hist), cap,
                                                     ctypes.byref(needed), _native.ptr(abs_sum), None)
        if code == _native.ERR_ARGUMENT and needed.value > cap:
            cap = int(needed.value)
            continue
        _native.check(code)
        return (mn, mx, hist, abs_sum)


def _entropy_from_counts(counts):
    """tools.py:523-537 on a histogram row: -sum f log2 f over the non-zero bins, float64, numpy order."""
    hist_non_zero = numpy.extract(counts != 0, counts).astype(numpy.int64)
    frequency = hist_non_zero.astype(numpy.float64)/numpy.sum(hist_non_zero)
    disc_entropy = -numpy.sum(frequency*numpy.log2(frequency))
    if disc_entropy < 0.:
        raise ValueError('The entropy is not positive.')
    if disc_entropy > numpy.log2(hist_non_zero.size):
        raise ValueError('The entropy is not smaller than its upper bound.')
    return disc_entropy


def _to_indices(quantized_samples, bin_width):
    """Quantized float samples -> int16 indices, with the reference's "quantization was omitted" check
    (tools.py:372-375) evaluated on the GPU through the exact rescale round trip."""
    if bin_width <= 0.:
        raise ValueError('The quantization bin width is not strictly positive.')
    data = numpy.asarray(quantized_samples)
    if not _is_float(data.dtype):
        data = data.astype(numpy.float64)
    # Same expression and dtype as tools.py:372-375 (float32 data stays float32).
    idx = cast_float_to_int16(numpy.asarray(data/bin_width, dtype=numpy.float32))
    numpy.testing.assert_almost_equal(bin_width*idx.astype(data.dtype), data, decimal=10,
                                      err_msg='The quantization was omitted.')
    return idx


def count_symbols(quantized_samples, bin_width):
    """tools.py:322-388: occurrences of each symbol from the smallest to the largest quantized sample."""
    idx = _to_indices(quantized_samples, bin_width).reshape((1, 1, -1, 1))
    (mn, mx, hist, _) = _histograms(idx, False)
    return hist[0, :int(mx[0]) - int(mn[0]) + 1].astype(numpy.int64)


def discrete_entropy(quantized_samples, bin_width):
    """tools.py:486-537."""
    return _entropy_from_counts(count_symbols(quantized_samples, bin_width))


def count_nb_deads(array_4d):
    """tools.py:294-320: per example, the number of maps whose coefficients are all zero."""
    if array_4d.ndim != 4:
        raise ValueError('`array_4d.ndim` is not equal to 4.')
    data = _as_float32(array_4d)
    (nb_examples, height_map, width_map, nb_maps) = data.shape
    nb_deads = numpy.empty(nb_examples, dtype=numpy.uint32)
    _native.check(_native.lib().eae_count_nb_deads_host(_native.ptr(data), nb_examples, height_map*width_map,
                                                        nb_maps, _native.ptr(nb_deads), None))
    return nb_deads.astype(numpy.int64)


def float_to_str(float_in):
    """tools.py:570-593: "." becomes "dot" for non-whole numbers, "-" becomes "minus" (used in result paths)."""
    float_in = float(float_in)
    str_in = str(int(float_in)) if float_in.is_integer() else str(float_in).replace('.', 'dot')
    return str_in.replace('-', 'minus')


def psnr_2d(reference_uint8, reconstruction_uint8):
    """tools.py:831-881: 10 log10(255^2 / mse) with the squared-error sum reduced on the GPU (exact
    integer), the division and logarithm in float64 as in the reference."""
    if reference_uint8.dtype != numpy.uint8:
        raise TypeError('`reference_uint8.dtype` is not equal to `numpy.uint8`.')
    if reconstruction_uint8.dtype != numpy.uint8:
        raise TypeError('`reconstruction_uint8.dtype` is not equal to `numpy.uint8`.')
    if reference_uint8.ndim != 2:
        raise ValueError('`reference_uint8.ndim` is not equal to 2.')
    if reference_uint8.shape != reconstruction_uint8.shape:
        raise ValueError('`reference_uint8.shape` is not equal to `reconstruction_uint8.shape`.')
    a = numpy.ascontiguousarray(reference_uint8)
    b = numpy.ascontiguousarray(reconstruction_uint8)
    sse = ctypes.c_uint64(0)
    _native.check(_native.lib().eae_sum_squared_error_u8_host(_native.ptr(a), _native.ptr(b), a.size,
                                                              ctypes.byref(sse), None))
    mse = numpy.float64(sse.value)/numpy.float64(a.size)
    if mse == 0.:
        raise ValueError('The mean squared error between the luminance image and its reconstruction is 0.')
    return 10.*numpy.log10((255.**2)/mse)


def quantize_per_map(data, bin_widths):
    """tools.py:883-929: bin_widths[i]*round(data[..., i]/bin_widths[i]), fp32, round half to even."""
    if bin_widths.ndim != 1:
        raise ValueError('`bin_widths.ndim` is not equal to 1.')
    (nb_examples, height_map, width_map, nb_maps) = data.shape
    if bin_widths.size != nb_maps:
        raise ValueError('`bin_widths.size` is not equal to `data.shape[3]`.')
    if numpy.any(bin_widths <= 0.):
        raise ValueError('A quantization bin width is not strictly positive.')
    x = _as_float32(data)
    bw = _as_float32(bin_widths)
    out = numpy.empty(x.shape, dtype=numpy.float32)
    _native.check(_native.lib().eae_quantize_per_map_host(_native.ptr(x), _native.ptr(out),
                                                          nb_examples*height_map*width_map, nb_maps,
                                                          _native.ptr(bw), None))
    return out


def rate_3d(quantized_latent_float32, bin_widths, h_in, w_in):
    """tools.py:931-989: sum over maps of entropy*h*w, divided by the number of pixels."""
    if bin_widths.ndim != 1:
        raise ValueError('`bin_widths.ndim` is not equal to 1.')
    (height_map, width_map, nb_maps) = quantized_latent_float32.shape
    if bin_widths.size != nb_maps:
        raise ValueError('`bin_widths.size` is not equal to `quantized_latent_float32.shape[2]`.')
    q = numpy.asarray(quantized_latent_float32)
    if not _is_float(q.dtype):
        q = q.astype(numpy.float64)
    bw = numpy.asarray(bin_widths).astype(q.dtype).reshape((1, 1, nb_maps))
    if numpy.any(bw <= 0.):
        raise ValueError('The quantization bin width is not strictly positive.')
    idx = cast_float_to_int16(numpy.asarray(q/bw, dtype=numpy.float32))
    numpy.testing.assert_almost_equal(bw*idx.astype(q.dtype), q, decimal=10, err_msg='The quantization was omitted.')
    (mn, mx, hist, _) = _histograms(idx.reshape((1, height_map, width_map, nb_maps)), False)
    cumulated_rate = 0.
    for i in range(nb_maps):
        disc_entropy = _entropy_from_counts(hist[i, :int(mx[i]) - int(mn[i]) + 1])
        cumulated_rate += disc_entropy*height_map*width_map
    return cumulated_rate/(h_in*w_in)


def save_image(path, array_uint8):
    """tools.py:1082-1106."""
    import PIL.Image
    if array_uint8.dtype != numpy.uint8:
        raise TypeError('`array_uint8.dtype` is not equal to `numpy.uint8`.')
    PIL.Image.fromarray(array_uint8).save(path)


def crop_repeat_2d(image_uint8, row_top_left, column_top_left):
    """tools.py:441-484: an 80 x 80 crop, every pixel repeated twice in both directions."""
    if image_uint8.dtype != numpy.uint8:
        raise TypeError('`image_uint8.dtype` is not equal to `numpy.uint8`.')
    (height_image, width_image) = image_uint8.shape
    if row_top_left + 80 >= height_image:
        raise ValueError('`image_uint8.shape[0]` is not strictly larger than `row_top_left + 80`.')
    if column_top_left + 80 >= width_image:
        raise ValueError('`image_uint8.shape[1]` is not strictly larger than `column_top_left + 80`.')
    return numpy.kron(image_uint8[row_top_left:row_top_left + 80, column_top_left:column_top_left + 80],
                      numpy.ones((2, 2), dtype=numpy.uint8))


def visualize_crops(image_uint8, positions_top_left, paths):
    """tools.py:1172-1218: one saved crop per column of `positions_top_left`."""
    if positions_top_left.ndim != 2 or positions_top_left.shape[0] != 2:
        raise ValueError('`positions_top_left.shape[0]` is not equal to 2.')
    if len(paths) != positions_top_left.shape[1]:
        raise ValueError('`len(paths)` is not equal to `positions_top_left.shape[1]`.')
    for (i, path) in enumerate(paths):
        save_image(path, crop_repeat_2d(image_uint8, positions_top_left[0, i].item(), positions_top_left[1, i].item()))


def visualize_rotated_luminance(luminance_before_rotation_uint8, is_rotated, positions_top_left, paths):
    """tools.py:1292-1330: the (possibly rotated) luminance image and its crops as PNG files. Host-side diagnostics the
    reference's evaluation loops call once per image (reconstructing_eae_kodak.py:229-232, 552-555); not on the GPU path."""
    image_uint8 = numpy.rot90(luminance_before_rotation_uint8, k=3).copy() if is_rotated \
        else luminance_before_rotation_uint8.copy()
    visualize_crops(image_uint8, positions_top_left, paths[1:])
    save_image(paths[0], image_uint8)


def subdivide_set(nb_examples, batch_size):
    """tools.py:1108-1132."""
    if nb_examples % batch_size != 0:
        raise ValueError('`nb_examples` is not divisible by `batch_size`.')
    return nb_examples//batch_size


def compute_bjontegaard(rates_0, psnrs_0, rates_1, psnrs_1):
    """tools.py:157-263: average per cent saving in bitrate between two rate-distortion curves (the figure
    reconstructing_eae_kodak.py reports next to its curves). Cubic fit of log-rate against PSNR for each curve,
    both integrated over the PSNR range the curves share. Host arithmetic (float64), same checks as the reference."""
    (rates_0, psnrs_0, rates_1, psnrs_1) = (numpy.asarray(a) for a in (rates_0, psnrs_0, rates_1, psnrs_1))
    if rates_0.ndim != 1:
        raise ValueError('`rates_0.ndim` is not equal to 1.')
    if rates_1.ndim != 1:
        raise ValueError('`rates_1.ndim` is not equal to 1.')
    if psnrs_0.shape != rates_0.shape:
        raise ValueError('`psnrs_0.shape` is not equal to `rates_0.shape`.')
    if psnrs_1.shape != rates_1.shape:
        raise ValueError('`psnrs_1.shape` is not equal to `rates_1.shape`.')
    for (name, values) in (('rates_0', rates_0), ('rates_1', rates_1), ('psnrs_0', psnrs_0), ('psnrs_1', psnrs_1)):
        numpy.testing.assert_array_less(0., values, err_msg='An element of `{}` is not strictly positive.'.format(name))
    (low, high) = (max(psnrs_0.min().item(), psnrs_1.min().item()), min(psnrs_0.max().item(), psnrs_1.max().item()))
    mean_log_rate = []
    for (rates, psnrs) in ((rates_0, psnrs_0), (rates_1, psnrs_1)):
        primitive = numpy.polyint(numpy.polyfit(psnrs, numpy.log(rates), 3))
        mean_log_rate.append(numpy.polyval(primitive, high) - numpy.polyval(primitive, low))
    return 100.*(numpy.exp((mean_log_rate[1] - mean_log_rate[0])/(high - low)).item() - 1.)

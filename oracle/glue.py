"""TEST INFRASTRUCTURE ONLY: numpy restatement of the numpy glue on the hot path.

Follows kodak_tensorflow/tools/tools.py (quantize_per_map :883-929, cast_float_to_int16 :95-133,
cast_bt601 :61-93, count_nb_deads :294-320, count_symbols :322-388, discrete_entropy :486-537,
rate_3d :931-989, psnr_2d :831-881), kodak_tensorflow/lossless/compression.py:11-154 and
kodak_tensorflow/lossless/stats.py:13-68, 136-241.

Pinned in tests/test_oracle_glue.py against the known answers in the reference's
kodak_tensorflow/test_tools.py / test_lossless.py and against the reference's own functions imported
in the build container (tests/golden/make_golden.py writes tests/golden/glue_*.npz).
"""
import numpy

from oracle import coder


def quantize_per_map(data, bin_widths):
    bw = numpy.asarray(bin_widths, dtype=numpy.float32).reshape((1,)*(data.ndim - 1) + (-1,))
    return bw*numpy.round(data/bw)


def cast_float_to_int16(array_float):
    rounded = numpy.round(array_float)
    if numpy.any(numpy.absolute(rounded) >= 32768.):
        raise AssertionError('The rounded array elements cannot be represented as 16-bit signed integers.')
    return rounded.astype(numpy.int16)


def cast_bt601(array_float):
    return numpy.round(array_float.clip(min=16., max=235.)).astype(numpy.uint8)


def count_nb_deads(array_4d):
    return numpy.sum(numpy.sum(numpy.absolute(array_4d), axis=(1, 2)) == 0, axis=1)


def count_symbols_int(indices):
    """Histogram of integer symbols from min to max (count_symbols with bin_width-scaled input)."""
    idx = numpy.asarray(indices).astype(numpy.int64).ravel()
    return numpy.bincount(idx - idx.min(), minlength=int(idx.max() - idx.min()) + 1)


def discrete_entropy_int(indices):
    hist = count_symbols_int(indices)
    hist = hist[hist != 0]
    freq = hist.astype(numpy.float64)/numpy.sum(hist)
    return -numpy.sum(freq*numpy.log2(freq))


def discrete_entropy(quantized_samples, bin_width):
    return discrete_entropy_int(numpy.round(numpy.asarray(quantized_samples, dtype=numpy.float64)/bin_width))


def rate_3d(quantized_latent_float32, bin_widths, h_in, w_in):
    (h, w, nb_maps) = quantized_latent_float32.shape
    total = 0.
    for i in range(nb_maps):
        total += discrete_entropy(quantized_latent_float32[:, :, i], float(bin_widths[i]))*h*w
    return total/(h_in*w_in)


def psnr_2d(reference_uint8, reconstruction_uint8):
    diff = reference_uint8.astype(numpy.float64) - reconstruction_uint8.astype(numpy.float64)
    mse = numpy.mean(diff**2)
    if mse == 0.:
        raise ValueError('The mean squared error between the luminance image and its reconstruction is 0.')
    return 10.*numpy.log10((255.**2)/mse)


def compress_lossless_maps(ref_int16, binary_probabilities, idx_map_exception=-1, which='port'):
    """compression.py:11-82 with the table passed as an array. Returns (rec_int16, uint32 bits[C])."""
    (h, w, nb_maps) = ref_int16.shape
    rec = numpy.zeros_like(ref_int16)
    bits = numpy.zeros(nb_maps, dtype=numpy.uint32)
    for i in range(nb_maps):
        if i == idx_map_exception:
            bits[i] = numpy.ceil(h*w*discrete_entropy_int(ref_int16[:, :, i])).astype(numpy.uint32)
            rec[:, :, i] = ref_int16[:, :, i]
        else:
            (err, out, nb) = coder.compress_lossless(ref_int16[:, :, i].flatten(), binary_probabilities[i, :],
                                                     which=which)
            if err:
                raise RuntimeError('Error of type {} during the encoding.'.format(err))
            rec[:, :, i] = out.reshape((h, w))
            bits[i] = nb
    return (rec, bits)


def rescale_compress_lossless_maps(centered_quantized_data, bin_widths_test, binary_probabilities,
                                   idx_map_exception=-1, which='port'):
    """compression.py:84-154."""
    bw = numpy.asarray(bin_widths_test, dtype=numpy.float32).reshape((1, 1, -1))
    ref_int16 = cast_float_to_int16(centered_quantized_data/bw)
    (rec, bits) = compress_lossless_maps(ref_int16, binary_probabilities, idx_map_exception, which=which)
    numpy.testing.assert_equal(centered_quantized_data, rec.astype(numpy.float32)*bw)
    return numpy.sum(bits).item()


# --- statistics (lossless/stats.py) -----------------------------------------------------------

def count_binary_decisions(abs_centered_quantized_data, bin_width_test, truncated_unary_length):
    """stats.py:136-195: zeros_j / ones_j of the truncated-unary prefix bins over |k|."""
    k = numpy.round(numpy.asarray(abs_centered_quantized_data, dtype=numpy.float64)/bin_width_test).astype(numpy.int64).ravel()
    L = truncated_unary_length
    zeros = numpy.zeros(L, dtype=numpy.int64)
    ones = numpy.zeros(L, dtype=numpy.int64)
    for j in range(L):
        zeros[j] = numpy.sum(k == j)
        ones[j] = numpy.sum(k > j)
    return (zeros, ones)


def binary_probabilities_from_counts(zeros, ones):
    """stats.py:59-67: P(bin_j = 0) with NaN -> 0.5, 0 -> 0.01, 1 -> 0.99."""
    with numpy.errstate(divide='ignore', invalid='ignore'):
        p = zeros.astype(numpy.float64)/(zeros + ones).astype(numpy.float64)
    p[numpy.isnan(p)] = 0.5
    p[p == 0.] = 0.01
    p[p == 1.] = 0.99
    return p


def probabilities_unit_intervals(data):
    """stats.py:70-134 with size_interval 1: numpy.histogram over unit bins from floor(min) to ceil(max)."""
    left = numpy.floor(numpy.amin(data)).item()
    right = numpy.ceil(numpy.amax(data)).item()
    if right - left < 1.:
        raise ValueError('The interval size exceeds the range of the data values.')
    edges = numpy.linspace(left, right, num=int(right - left) + 1)
    return (edges, numpy.histogram(data, bins=edges, density=True)[0]*1.)


def jensen_shannon_divergence(probs_0, probs_1):
    """tools.py:615-666."""
    denominator = 0.5*(probs_0 + probs_1)
    return 0.5*numpy.sum(probs_0*numpy.log2(probs_0/denominator) + probs_1*numpy.log2(probs_1/denominator))


def find_index_map_exception(y_float32):
    """stats.py:197-241: the map whose unit-interval histogram is closest (Jensen-Shannon) to uniform."""
    nb_maps = y_float32.shape[3]
    divergences = numpy.zeros(nb_maps)
    for i in range(nb_maps):
        probs = probabilities_unit_intervals(y_float32[:, :, :, i])[1]
        nz = numpy.extract(probs != 0., probs)
        divergences[i] = jensen_shannon_divergence(nz, numpy.ones(nz.size)/nz.size) if nz.size > 1 else 1.
    return numpy.argmin(divergences).item()

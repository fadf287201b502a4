"""Statistics on the latent variable feature maps: the reference's ``lossless/stats.py`` API
(kodak_tensorflow/lossless/stats.py:13-320) with the counting done on the GPU.

``save_statistics`` produces, for a calibration set, the three files the lossless coder needs for a
given model: the latent variable feature map means, the index of the map that is not compressed as the
others, and one table of binary probabilities per set of test quantization bin widths. The sums,
extrema, unit-interval histograms and truncated-unary bin counts come from
``eae_latent_statistics_host`` (one pass over the latents each, integer-exact); the float64 epilogues
below follow the reference's numpy expressions term by term, so the tables are bit-identical to the
reference's given the same latents.
"""
import ctypes
import os
import pickle

import numpy

from autoencoder_based_image_compression_b200 import _native

# The functions are sorted in alphabetic order, as in the reference.


def _as_latents(y_float32):
    y = numpy.ascontiguousarray(y_float32)
    if y.dtype != numpy.float32:
        raise TypeError('`y_float32.dtype` is not equal to `numpy.float32`.')
    if y.ndim != 4:
        raise ValueError('`y_float32.ndim` is not equal to 4.')
    return y


def _statistics(y, mean=None, delta=None, truncated_unary_length=0, unit_histograms=False):
    """One call of eae_latent_statistics_host. Returns (sums, minima, maxima, unit_hist or None, abs_counts or None)."""
    nb_maps = y.shape[-1]
    n_rows = y.size//nb_maps
    lib = _native.lib()
    sums = numpy.zeros(nb_maps, dtype=numpy.float64)
    minima = numpy.zeros(nb_maps, dtype=numpy.float32)
    maxima = numpy.zeros(nb_maps, dtype=numpy.float32)
    counts = None
    (mean_p, delta_p, counts_p) = (None, None, None)
    if mean is not None:
        mean = numpy.ascontiguousarray(mean, dtype=numpy.float32)
        delta = numpy.ascontiguousarray(delta, dtype=numpy.float32)
        if mean.shape != (nb_maps,) or delta.shape != (nb_maps,):
            raise ValueError('`map_mean` and `bin_widths_test` must hold one value per latent variable feature map.')
        counts = numpy.zeros((nb_maps, truncated_unary_length + 1), dtype=numpy.uint64)
        (mean_p, delta_p, counts_p) = (_native.ptr(mean), _native.ptr(delta), _native.ptr(counts))
    needed = ctypes.c_uint32(0)
    hist = None
    cap = 0
    if unit_histograms:
        # first call sizes the histograms, second call fills them
        _native.check(lib.eae_latent_statistics_host(_native.ptr(y), n_rows, nb_maps, _native.ptr(sums), _native.ptr(minima),
                                                     _native.ptr(maxima), None, 0, ctypes.byref(needed), None, None, 0,
                                                     None, None))
        cap = int(needed.value)
        hist = numpy.zeros((nb_maps, cap), dtype=numpy.uint64)
    _native.check(lib.eae_latent_statistics_host(_native.ptr(y), n_rows, nb_maps, _native.ptr(sums), _native.ptr(minima),
                                                 _native.ptr(maxima), None if hist is None else _native.ptr(hist), cap,
                                                 ctypes.byref(needed), mean_p, delta_p, int(truncated_unary_length),
                                                 counts_p, None))
    return (sums, minima, maxima, hist, counts)


def _decisions_from_counts(counts, truncated_unary_length):
    """counts[a], a = 0 .. L (a = L gathers every symbol >= L) -> (zeros, ones) of stats.py:184-194."""
    c = counts.astype(numpy.int64)
    zeros = c[..., :truncated_unary_length].copy()
    # ones[j] = number of symbols strictly larger than j
    tail = numpy.cumsum(c[..., ::-1], axis=-1)[..., ::-1]
    ones = tail[..., 1:truncated_unary_length + 1].copy()
    return (zeros, ones)


def compute_binary_probabilities(y_float32, bin_widths_test, map_mean, truncated_unary_length):
    """Binary probabilities of each latent variable feature map (stats.py:13-68): element [i, j] is the
    probability that the jth binary decision of the truncated unary prefix of map i is 0; decisions that never
    occur get 0.5, probabilities 0 and 1 become 0.01 and 0.99."""
    y = _as_latents(y_float32)
    bin_widths_test = numpy.asarray(bin_widths_test, dtype=numpy.float32)
    if numpy.any(bin_widths_test <= 0.):
        raise ValueError('A quantization bin width is not strictly positive.')      # tools.py:918-919
    if truncated_unary_length < 1:
        raise ValueError('`truncated_unary_length` is not strictly positive.')
    counts = _statistics(y, map_mean, bin_widths_test, truncated_unary_length)[4]
    (cumulated_zeros, cumulated_ones) = _decisions_from_counts(counts, truncated_unary_length)
    total = cumulated_zeros + cumulated_ones
    with numpy.errstate(invalid='ignore', divide='ignore'):
        binary_probabilities = cumulated_zeros.astype(numpy.float64)/total.astype(numpy.float64)
    binary_probabilities[numpy.isnan(binary_probabilities)] = 0.5
    binary_probabilities[binary_probabilities == 0.] = 0.01
    binary_probabilities[binary_probabilities == 1.] = 0.99
    return binary_probabilities


def compute_probabilities_intervals(data, size_interval):
    """Probability that a data value belongs to each axis interval of size `size_interval` between
    floor(min) and ceil(max) (stats.py:70-134). Returns (bin_edges, probabilities), both float64."""
    data = numpy.asarray(data)
    edge_left = numpy.floor(numpy.amin(data)).item()
    edge_right = numpy.ceil(numpy.amax(data)).item()
    difference_edges = edge_right - edge_left
    if difference_edges < size_interval:
        raise ValueError('The interval size exceeds the range of the data values.')
    nb_edges_minus_1_float = difference_edges/size_interval
    if float(nb_edges_minus_1_float).is_integer():
        nb_edges = int(nb_edges_minus_1_float) + 1
    else:
        raise ValueError('The range of the data values cannot be split into '
                         + 'an integer number of intervals of size {}.'.format(size_interval))
    bin_edges = numpy.linspace(edge_left, edge_right, num=nb_edges)
    if size_interval == 1. and data.dtype == numpy.float32:
        # unit intervals: the counts come from the GPU (one map = a [n, 1] latent array)
        hist = _statistics(numpy.ascontiguousarray(data).reshape((-1, 1, 1, 1)), unit_histograms=True)[3]
        counts = hist[0, :nb_edges - 1].astype(numpy.int64)
    else:
        counts = numpy.histogram(data, bins=bin_edges)[0]
    return (bin_edges, _density(counts, bin_edges)*size_interval)


def _density(counts, bin_edges):
    # numpy.histogram(..., density=True): n / diff(edges) / n.sum()
    db = numpy.array(numpy.diff(bin_edges), float)
    return counts/db/counts.sum()


def count_binary_decisions(abs_centered_quantized_data, bin_width_test, truncated_unary_length):
    """Occurrences of 0 and of 1 for each binary decision of the truncated unary prefix of the absolute
    centered-quantized data (stats.py:136-195). Returns two int64 arrays of length `truncated_unary_length`."""
    data = numpy.ascontiguousarray(abs_centered_quantized_data, dtype=numpy.float32)
    if numpy.any(data < 0.):
        raise ValueError('An element of `abs_centered_quantized_data` is not positive.')
    if data.size == 0:
        raise ValueError('`abs_centered_quantized_data` is empty.')
    y = data.reshape((-1, 1, 1, 1))
    counts = _statistics(y, numpy.zeros(1, dtype=numpy.float32), numpy.array([bin_width_test], dtype=numpy.float32),
                         truncated_unary_length)[4]
    (zeros, ones) = _decisions_from_counts(counts[0], truncated_unary_length)
    return (zeros, ones)


def jensen_shannon_divergence(probs_0, probs_1):
    """tools.py:615-666 (also exported by this package's tools.tools)."""
    if numpy.any(probs_0 <= 0.) or numpy.any(probs_0 >= 1.):
        raise ValueError('A probability in `probs_0` does not belong to ]0.0, 1.0[.')
    if numpy.any(probs_1 <= 0.) or numpy.any(probs_1 >= 1.):
        raise ValueError('A probability in `probs_1` does not belong to ]0.0, 1.0[.')
    if abs(numpy.sum(probs_0).item() - 1.) >= 1.e-9:
        raise ValueError('The probabilities in `probs_0` do not sum to 1.0.')
    if abs(numpy.sum(probs_1).item() - 1.) >= 1.e-9:
        raise ValueError('The probabilities in `probs_1` do not sum to 1.0.')
    denominator = 0.5*(probs_0 + probs_1)
    divergence = 0.5*numpy.sum(probs_0*numpy.log2(probs_0/denominator) + probs_1*numpy.log2(probs_1/denominator))
    if divergence < 0.:
        raise ValueError('The Jensen-Shannon divergence is not positive.')
    if divergence > 1.:
        raise ValueError('The Jensen-Shannon divergence is not smaller than 1.0.')
    return divergence


def find_index_map_exception(y_float32):
    """Index of the latent variable feature map whose distribution is closest to uniform (stats.py:197-241):
    the map that minimises the Jensen-Shannon divergence between its unit-interval histogram and the uniform
    distribution over the occupied intervals."""
    y = _as_latents(y_float32)
    nb_maps = y.shape[3]
    (_, minima, maxima, hist, _) = _statistics(y, unit_histograms=True)
    divergences = numpy.zeros(nb_maps)
    for i in range(nb_maps):
        edge_left = numpy.floor(minima[i]).item()
        edge_right = numpy.ceil(maxima[i]).item()
        difference_edges = edge_right - edge_left
        if difference_edges < 1.:
            raise ValueError('The interval size exceeds the range of the data values.')
        nb_edges = int(difference_edges) + 1
        bin_edges = numpy.linspace(edge_left, edge_right, num=nb_edges)
        probs = _density(hist[i, :nb_edges - 1].astype(numpy.int64), bin_edges)*1.
        probs_non_zero = numpy.extract(probs != 0., probs)
        nb_remaining_probs = probs_non_zero.size
        if nb_remaining_probs > 1:
            uniform_probs = (1./nb_remaining_probs)*numpy.ones(nb_remaining_probs)
            divergences[i] = jensen_shannon_divergence(probs_non_zero, uniform_probs)
        else:
            divergences[i] = 1.
    return numpy.argmin(divergences).item()


def save_statistics(luminances_uint8, sess, entropy_ae, batch_size, multipliers, truncated_unary_length,
                    path_to_map_mean, path_to_idx_map_exception, paths_to_binary_probabilities):
    """Saves the statistics of the latent variable feature maps of `entropy_ae` on a calibration set
    (stats.py:243-320): `map_mean.npy`, `idx_map_exception.pkl` and one `binary_probabilities_*.npy` per
    multiplier of the quantization bin widths. Existing files are kept, as in the reference."""
    from autoencoder_based_image_compression_b200.kodak_tensorflow.eae import batching

    multipliers = numpy.asarray(multipliers)
    nb_multipliers = multipliers.size
    if len(paths_to_binary_probabilities) != nb_multipliers:
        raise ValueError('`len(paths_to_binary_probabilities)` is not equal to `multipliers.size`.')
    booleans = [os.path.isfile(path) for path in paths_to_binary_probabilities]
    if os.path.isfile(path_to_map_mean) and os.path.isfile(path_to_idx_map_exception) and all(booleans):
        print('The statistics on the latent variable feature maps already exist.')
        print('Delete them manually to recompute them.')
        return
    y_float32 = batching.encode_mini_batches(luminances_uint8, sess, entropy_ae, batch_size)
    map_mean = numpy.mean(y_float32, axis=(0, 1, 2))
    numpy.save(path_to_map_mean, map_mean)
    idx_map_exception = find_index_map_exception(y_float32)
    with open(path_to_idx_map_exception, 'wb') as file:
        pickle.dump(idx_map_exception, file, protocol=2)
    for i in range(nb_multipliers):
        bin_widths_test = multipliers[i]*entropy_ae.get_bin_widths()
        binary_probabilities = compute_binary_probabilities(y_float32, bin_widths_test, map_mean,
                                                            truncated_unary_length)
        numpy.save(paths_to_binary_probabilities[i], binary_probabilities)

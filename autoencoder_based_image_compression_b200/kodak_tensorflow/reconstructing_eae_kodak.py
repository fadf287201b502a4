"""Rate / PSNR evaluation of entropy autoencoders: the two evaluation loops of the reference's
``reconstructing_eae_kodak.py`` (``fix_gamma`` :31-245, ``vary_gamma_fix_bin_widths`` :401-556) on the B200 path.

Same arguments, same result files layout, same return values (two float64 arrays ``[nb_points, nb_images]``). What
differs from the reference script:

* ``tf.Session`` is ``codec.Session`` (device + arithmetic mode); ``tf.reset_default_graph()`` has no counterpart;
* a model is restored from ``eae/results/<suffix>/model_<idx>.npz`` (``weights.save``; TensorFlow checkpoints cannot be
  read without TensorFlow), falling back to the ``.ckpt`` path's stem, and a missing file draws the reference's initial
  distributions when ``allow_random_init`` is set (tests / benchmarks without trained checkpoints);
* the PNG visualisations (``tls.visualize_rotated_luminance``, ``plot_nb_dead_feature_maps``) are not written:
  ``path_to_checking_r``, ``list_rotation`` and ``positions_top_left`` are accepted and ignored (plotting
  is out of scope).
"""
import os
import pickle

import numpy

from autoencoder_based_image_compression_b200 import codec as native_codec
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae import batching
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.EntropyAutoencoder import EntropyAutoencoder
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.IsolatedDecoder import IsolatedDecoder
from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import compression
from autoencoder_based_image_compression_b200.kodak_tensorflow.tools import tools as tls


def _path_to_restore(suffix, idx_training, allow_random_init):
    stem = 'eae/results/{0}/model_{1}'.format(suffix, idx_training)
    for ext in ('.npz', '.ckpt.npz'):
        if os.path.isfile(stem + ext):
            return stem + ext
    if allow_random_init:
        return ''
    raise IOError('{}.npz does not exist (export the TensorFlow checkpoint with weights.save).'.format(stem))


def fix_gamma(reference_uint8, bin_width_init, multipliers, idx_training, gamma_scaling, batch_size,
              are_bin_widths_learned, is_lossless, path_to_checking_r='', list_rotation=(), positions_top_left=None,
              device=0, math='tf32x3', allow_random_init=False):
    """A series of pairs (rate, PSNR) for ONE entropy autoencoder and several sets of test quantization bin widths
    (reconstructing_eae_kodak.py:31-245): element [i, j] is the rate (bits per pixel; the lossless coder's bit count if
    `is_lossless`, the empirical entropy otherwise) / the PSNR of the jth luminance image at the ith multiplier."""
    multipliers = numpy.asarray(multipliers, dtype=numpy.float32)
    nb_points = multipliers.size
    (nb_images, h_in, w_in) = reference_uint8.shape
    rate = numpy.zeros((nb_points, nb_images))
    psnr = numpy.zeros((nb_points, nb_images))
    if are_bin_widths_learned:
        suffix = 'learning_bw_{0}_{1}'.format(tls.float_to_str(bin_width_init), tls.float_to_str(gamma_scaling))
    else:
        suffix = '{0}_{1}'.format(tls.float_to_str(bin_width_init), tls.float_to_str(gamma_scaling))
    path_to_nb_itvs_per_side_load = 'eae/results/{0}/nb_itvs_per_side_{1}.pkl'.format(suffix, idx_training)
    path_to_restore = _path_to_restore(suffix, idx_training, allow_random_init)
    path_to_stats = 'lossless/results/{0}/training_index_{1}/'.format(suffix, idx_training)
    path_to_map_mean = os.path.join(path_to_stats, 'map_mean.npy')

    entropy_ae = EntropyAutoencoder(batch_size, h_in, w_in, bin_width_init, gamma_scaling, path_to_nb_itvs_per_side_load,
                                    are_bin_widths_learned)
    with native_codec.Session(device=device, math=math) as sess:
        entropy_ae.initialization(sess, path_to_restore)
        y_float32 = batching.encode_mini_batches(numpy.expand_dims(reference_uint8, axis=3), sess, entropy_ae, batch_size)
        bin_widths = entropy_ae.get_bin_widths()
    isolated_decoder = IsolatedDecoder(batch_size, h_in, w_in, are_bin_widths_learned)
    array_nb_deads = numpy.zeros((nb_points, nb_images), dtype=numpy.int32)
    map_mean = numpy.load(path_to_map_mean)
    tiled_map_mean = numpy.tile(map_mean, (nb_images, y_float32.shape[1], y_float32.shape[2], 1))
    idx_map_exception = -1
    if is_lossless:
        with open(os.path.join(path_to_stats, 'idx_map_exception.pkl'), 'rb') as file:
            idx_map_exception = pickle.load(file)
    centered_y_float32 = y_float32 - tiled_map_mean
    with native_codec.Session(device=device, math=math) as sess:
        if path_to_restore:
            isolated_decoder.initialization(sess, path_to_restore)
        else:
            isolated_decoder.set_weights(entropy_ae.weights)      # the same random draw for both halves
        for i in range(nb_points):
            multiplier = multipliers[i].item()
            str_multiplier = tls.float_to_str(multiplier)
            bin_widths_test = multiplier*bin_widths
            centered_quantized_y_float32 = tls.quantize_per_map(centered_y_float32, bin_widths_test)
            array_nb_deads[i, :] = tls.count_nb_deads(centered_quantized_y_float32)
            off_centered_quantized_y_float32 = centered_quantized_y_float32 + tiled_map_mean
            expanded_reconstruction_uint8 = batching.decode_mini_batches(off_centered_quantized_y_float32, sess,
                                                                         isolated_decoder, batch_size)
            reconstruction_uint8 = numpy.squeeze(expanded_reconstruction_uint8, axis=3)
            if is_lossless:
                path_to_binary_probabilities = os.path.join(path_to_stats,
                                                            'binary_probabilities_{}.npy'.format(str_multiplier))
            for j in range(nb_images):
                if is_lossless:
                    nb_bits = compression.rescale_compress_lossless_maps(centered_quantized_y_float32[j, :, :, :],
                                                                         bin_widths_test, path_to_binary_probabilities,
                                                                         idx_map_exception=idx_map_exception)
                    rate[i, j] = float(nb_bits)/(h_in*w_in)
                else:
                    rate[i, j] = tls.rate_3d(centered_quantized_y_float32[j, :, :, :], bin_widths_test, h_in, w_in)
                psnr[i, j] = tls.psnr_2d(reference_uint8[j, :, :], reconstruction_uint8[j, :, :])
    fix_gamma.last_nb_deads = array_nb_deads
    return (rate, psnr)


def vary_gamma_fix_bin_widths(reference_uint8, bin_width_init, idxs_training, gammas_scaling, batch_size,
                              path_to_checking_r='', list_rotation=(), positions_top_left=None, device=0, math='tf32x3',
                              allow_random_init=False):
    """A series of pairs (rate, PSNR), one entropy autoencoder per scaling coefficient, quantization bin widths fixed at
    training time (reconstructing_eae_kodak.py:401-556): rate = empirical entropy of the quantized latents."""
    gammas_scaling = numpy.asarray(gammas_scaling)
    idxs_training = numpy.asarray(idxs_training)
    nb_points = gammas_scaling.size
    if idxs_training.size != nb_points:
        raise ValueError('`gammas_scaling.size` is not equal to `idxs_training.size`.')
    (nb_images, h_in, w_in) = reference_uint8.shape
    rate = numpy.zeros((nb_points, nb_images))
    psnr = numpy.zeros((nb_points, nb_images))
    for i in range(nb_points):
        gamma_scaling = gammas_scaling[i].item()
        idx_training = idxs_training[i].item()
        suffix = '{0}_{1}'.format(tls.float_to_str(bin_width_init), tls.float_to_str(gamma_scaling))
        path_to_nb_itvs_per_side_load = 'eae/results/{0}/nb_itvs_per_side_{1}.pkl'.format(suffix, idx_training)
        path_to_restore = _path_to_restore(suffix, idx_training, allow_random_init)
        entropy_ae = EntropyAutoencoder(batch_size, h_in, w_in, bin_width_init, gamma_scaling,
                                        path_to_nb_itvs_per_side_load, False)
        with native_codec.Session(device=device, math=math) as sess:
            entropy_ae.initialization(sess, path_to_restore, seed=i)
            y_float32 = batching.encode_mini_batches(numpy.expand_dims(reference_uint8, axis=3), sess, entropy_ae,
                                                     batch_size)
            bin_widths = entropy_ae.get_bin_widths()
        isolated_decoder = IsolatedDecoder(batch_size, h_in, w_in, False)
        quantized_y_float32 = tls.quantize_per_map(y_float32, bin_widths)
        with native_codec.Session(device=device, math=math) as sess:
            if path_to_restore:
                isolated_decoder.initialization(sess, path_to_restore)
            else:
                isolated_decoder.set_weights(entropy_ae.weights)
            expanded_reconstruction_uint8 = batching.decode_mini_batches(quantized_y_float32, sess, isolated_decoder,
                                                                         batch_size)
        reconstruction_uint8 = numpy.squeeze(expanded_reconstruction_uint8, axis=3)
        for j in range(nb_images):
            rate[i, j] = tls.rate_3d(quantized_y_float32[j, :, :, :], bin_widths, h_in, w_in)
            psnr[i, j] = tls.psnr_2d(reference_uint8[j, :, :], reconstruction_uint8[j, :, :])
    return (rate, psnr)

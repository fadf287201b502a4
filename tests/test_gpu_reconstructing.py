"""The evaluation loops of reconstructing_eae_kodak.py (fix_gamma :31-245, vary_gamma_fix_bin_widths :401-556) end to end
on the GPU path: statistics collected with lossless.stats on a calibration set, model restored from an .npz, rate / PSNR
per multiplier, against the CPU oracle pipeline (torch transforms, numpy glue, C coder) on the same images and files."""
import os
import pickle

import numpy
import pytest

from autoencoder_based_image_compression_b200 import codec as native_codec
from autoencoder_based_image_compression_b200 import weights as wts
from oracle import coder as oracle_coder
from oracle import glue as oracle_glue
from oracle import transforms as oracle_transforms
from tests import util

pytestmark = pytest.mark.gpu


def visible_weights(seed, learned):
    w = wts.random_init(seed, learned)
    w['decoder/biases_5'] = (w['decoder/biases_5'] + 2.0).astype(numpy.float32)
    w['decoder/weights_6'] = (numpy.abs(w['decoder/weights_6'])*8.).astype(numpy.float32)
    return w


def test_fix_gamma_and_vary_gamma_against_the_oracle(native, tmp_path, monkeypatch):
    from autoencoder_based_image_compression_b200.kodak_tensorflow import reconstructing_eae_kodak as rek
    from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.EntropyAutoencoder import EntropyAutoencoder
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import stats
    monkeypatch.chdir(tmp_path)
    rng = numpy.random.default_rng(21)
    (h, wd) = (64, 96)
    weights = visible_weights(3, False)
    suffix = '1_10000'
    os.makedirs('eae/results/' + suffix)
    path_stats = 'lossless/results/{}/training_index_10/'.format(suffix)
    os.makedirs(path_stats)
    wts.save('eae/results/{}/model_10.npz'.format(suffix), weights)
    multipliers = numpy.array([1., 2., 4.], dtype=numpy.float32)

    # statistics on a calibration set (collecting_stats_eae_extra.py:43-90)
    extra = util.synthetic_luma(rng, 8, h, wd)[..., None]
    entropy_ae = EntropyAutoencoder(4, h, wd, 1., 10000., '', False)
    with native_codec.Session(device=0, math='tf32x3') as sess:
        entropy_ae.initialization(sess, 'eae/results/{}/model_10.npz'.format(suffix))
        stats.save_statistics(extra, sess, entropy_ae, 4, multipliers, 10, path_stats + 'map_mean.npy',
                              path_stats + 'idx_map_exception.pkl',
                              [path_stats + 'binary_probabilities_{}.npy'.format(m) for m in ('1', '2', '4')])

    kodak = util.synthetic_luma(rng, 2, h, wd)
    (rate_l, psnr_l) = rek.fix_gamma(kodak, 1., multipliers, 10, 10000., 2, False, True)
    (rate_a, psnr_a) = rek.fix_gamma(kodak, 1., multipliers, 10, 10000., 2, False, False)
    assert rate_l.shape == psnr_l.shape == (3, 2) and rate_l.dtype == numpy.float64
    assert numpy.array_equal(psnr_l, psnr_a)                       # the reconstruction does not depend on the coder

    # oracle: same files, CPU arithmetic
    map_mean = numpy.load(path_stats + 'map_mean.npy')
    with open(path_stats + 'idx_map_exception.pkl', 'rb') as f:
        idx_exc = pickle.load(f)
    y = oracle_transforms.encoder(kodak[..., None].astype(numpy.float32), weights, False)
    centered = y - map_mean.reshape((1, 1, 1, -1))
    for (i, m) in enumerate(('1', '2', '4')):
        bw = multipliers[i]*numpy.ones(128, dtype=numpy.float32)
        cq = oracle_glue.quantize_per_map(centered, bw)
        rec = oracle_glue.cast_bt601(oracle_transforms.decoder(cq + map_mean.reshape((1, 1, 1, -1)), weights, False))[..., 0]
        table = numpy.load(path_stats + 'binary_probabilities_{}.npy'.format(m))
        for j in range(2):
            bits = oracle_glue.rescale_compress_lossless_maps(cq[j], bw, table, idx_map_exception=idx_exc)
            want_rate_l = float(bits)/(h*wd)
            want_rate_a = oracle_glue.rate_3d(cq[j], bw, h, wd)
            assert abs(rate_l[i, j] - want_rate_l) <= 1e-3*want_rate_l + 2./(h*wd), (i, j, rate_l[i, j], want_rate_l)
            assert abs(rate_a[i, j] - want_rate_a) <= 1e-3*want_rate_a, (i, j, rate_a[i, j], want_rate_a)
            assert abs(psnr_l[i, j] - oracle_glue.psnr_2d(kodak[j], rec[j])) < 0.01
    assert numpy.all(numpy.diff(rate_a, axis=0) < 0) and numpy.all(rate_l > 0)        # coarser bins, fewer bits

    # one model per scaling coefficient, entropy rates (the second model is a random draw)
    (rate_v, psnr_v) = rek.vary_gamma_fix_bin_widths(kodak, 1., numpy.array([10, 10]), numpy.array([10000., 12000.]), 2,
                                                     allow_random_init=True)
    assert rate_v.shape == (2, 2)
    assert numpy.allclose(rate_v[0], [oracle_glue.rate_3d(oracle_glue.quantize_per_map(y[j], numpy.ones(128, dtype=numpy.float32)),
                                                          numpy.ones(128, dtype=numpy.float32), h, wd) for j in range(2)], rtol=1e-3)
    with pytest.raises(ValueError):
        rek.vary_gamma_fix_bin_widths(kodak, 1., numpy.array([10]), numpy.array([1., 2.]), 2)
    with pytest.raises(IOError):
        rek.fix_gamma(kodak, 1., multipliers, 11, 10000., 2, False, False)

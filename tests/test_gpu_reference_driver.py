"""The drop-in, proven with the reference's own driver: ``kodak_tensorflow/reconstructing_eae_kodak.py`` is imported
UNMODIFIED (tests/refdrivers.py) and its ``fix_gamma`` (:31-245, with and without the lossless coder) and
``vary_gamma_fix_bin_widths`` (:401-556) run on the B200 path - its own sequencing, its own file layout
(``eae/results/<suffix>/model_k.ckpt``, ``lossless/results/<suffix>/training_index_k/*``), its own PNG calls. The
statistics files are produced by the reference's ``collecting_stats_eae_extra``-style call into ``lossless.stats``.
Results are compared with the CPU oracle pipeline (torch transforms, numpy glue, the reference's compiled C++ coder
when it is there) on the same images and files."""
import os
import pickle

import numpy
import pytest

from autoencoder_based_image_compression_b200 import weights as wts
from oracle import coder as oracle_coder
from oracle import glue as oracle_glue
from oracle import transforms as oracle_transforms
from tests import refdrivers
from tests import util

pytestmark = pytest.mark.gpu


def visible_weights(seed, learned):
    w = wts.random_init(seed, learned)
    w['decoder/biases_5'] = (w['decoder/biases_5'] + 2.0).astype(numpy.float32)
    w['decoder/weights_6'] = (numpy.abs(w['decoder/weights_6'])*8.).astype(numpy.float32)
    return w


def test_the_reference_driver_runs_unmodified_on_this_package(native, tmp_path, monkeypatch):
    rek = refdrivers.load('reconstructing_eae_kodak.py')
    if rek is None:
        pytest.skip('the reference driver is neither staged under oracle/_ref/drivers nor at /root/reference')
    import tensorflow as tf                     # the stand-in of the mirror directory
    from eae.graph.EntropyAutoencoder import EntropyAutoencoder
    import lossless.stats
    assert tf.__file__.startswith(refdrivers.MIRROR) and rek.tf is tf
    assert rek.eae.batching.__file__.startswith(refdrivers.MIRROR)
    assert rek.lossless.compression.__file__.startswith(refdrivers.MIRROR)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('EAE_MATH', 'tf32x3')
    rng = numpy.random.default_rng(21)
    (h, wd) = (128, 192)
    which = 'ref' if oracle_coder.has_ref() else 'port'
    multipliers = numpy.array([1., 2., 4.], dtype=numpy.float32)
    models = {'1_10000': visible_weights(3, False), '1_12000': visible_weights(4, False)}
    for (suffix, weights) in models.items():
        os.makedirs('eae/results/' + suffix)
        wts.save('eae/results/{}/model_10.npz'.format(suffix), weights)       # restored through the '.ckpt' prefix
    path_stats = 'lossless/results/1_10000/training_index_10/'
    os.makedirs(path_stats)

    # statistics on a calibration set, as collecting_stats_eae_extra.py:43-90 does
    extra = util.synthetic_luma(rng, 8, h, wd)[..., None]
    entropy_ae = EntropyAutoencoder(4, h, wd, 1., 10000., '', False)
    with tf.Session() as sess:
        entropy_ae.initialization(sess, 'eae/results/1_10000/model_10.ckpt')
        lossless.stats.save_statistics(extra, sess, entropy_ae, 4, multipliers, 10, path_stats + 'map_mean.npy',
                                       path_stats + 'idx_map_exception.pkl',
                                       [path_stats + 'binary_probabilities_{}.npy'.format(m) for m in ('1', '2', '4')])

    kodak = util.synthetic_luma(rng, 2, h, wd)
    positions_top_left = numpy.array([[10], [20]], dtype=numpy.int32)
    checking = str(tmp_path / 'checking')
    (rate_l, psnr_l) = rek.fix_gamma(kodak, 1., multipliers, 10, 10000., 2, False, True, checking, [1], positions_top_left)
    (rate_a, psnr_a) = rek.fix_gamma(kodak, 1., multipliers, 10, 10000., 2, False, False, checking, [1], positions_top_left)
    assert rate_l.shape == psnr_l.shape == (3, 2) and rate_l.dtype == numpy.float64
    assert numpy.array_equal(psnr_l, psnr_a)                       # the reconstruction does not depend on the coder
    assert os.path.isfile(os.path.join(checking, 'reconstruction_fix_gamma', '1_10000', 'lossless', 'multiplier_2',
                                       'reconstruction_1_crop_0.png'))

    # oracle: same files, CPU arithmetic
    weights = models['1_10000']
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
            bits = oracle_glue.rescale_compress_lossless_maps(cq[j], bw, table, idx_map_exception=idx_exc, which=which)
            want_rate_l = float(bits)/(h*wd)
            want_rate_a = oracle_glue.rate_3d(cq[j], bw, h, wd)
            # north star: rate within 0.1 %, PSNR within 0.01 dB
            assert abs(rate_l[i, j] - want_rate_l) <= 1e-3*want_rate_l, (i, j, rate_l[i, j], want_rate_l)
            assert abs(rate_a[i, j] - want_rate_a) <= 1e-3*want_rate_a, (i, j, rate_a[i, j], want_rate_a)
            assert abs(psnr_l[i, j] - oracle_glue.psnr_2d(kodak[j], rec[j])) < 0.01
    assert numpy.all(numpy.diff(rate_a, axis=0) < 0) and numpy.all(rate_l > 0)        # coarser bins, fewer bits

    # one model per scaling coefficient, entropy rates, no centring (:401-556)
    (rate_v, psnr_v) = rek.vary_gamma_fix_bin_widths(kodak, 1., numpy.array([10, 10]), numpy.array([10000., 12000.]), 2,
                                                     checking, [], numpy.zeros((2, 0), dtype=numpy.int32))
    assert rate_v.shape == psnr_v.shape == (2, 2)
    ones = numpy.ones(128, dtype=numpy.float32)
    for (i, suffix) in enumerate(('1_10000', '1_12000')):
        yi = oracle_transforms.encoder(kodak[..., None].astype(numpy.float32), models[suffix], False)
        qi = oracle_glue.quantize_per_map(yi, ones)
        reci = oracle_glue.cast_bt601(oracle_transforms.decoder(qi, models[suffix], False))[..., 0]
        for j in range(2):
            want = oracle_glue.rate_3d(qi[j], ones, h, wd)
            assert abs(rate_v[i, j] - want) <= 1e-3*want
            assert abs(psnr_v[i, j] - oracle_glue.psnr_2d(kodak[j], reci[j])) < 0.01
    with pytest.raises(ValueError):
        rek.vary_gamma_fix_bin_widths(kodak, 1., numpy.array([10]), numpy.array([1., 2.]), 2, checking, [],
                                      numpy.zeros((2, 0), dtype=numpy.int32))
    with pytest.raises(IOError):
        rek.fix_gamma(kodak, 1., multipliers, 11, 10000., 2, False, False, checking, [], numpy.zeros((2, 0), dtype=numpy.int32))

"""GPU versions of the numpy glue (tools/tools.py) against the reference's own outputs (glue.npz) and
the oracle. Bit-exact for fp32 / integer results, 1e-12 for the float64 entropy / PSNR epilogues."""
import numpy
import pytest

from oracle import glue as oracle_glue

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tls(native):
    from autoencoder_based_image_compression_b200.kodak_tensorflow.tools import tools
    return tools


def test_quantize_per_map(tls, golden):
    g = golden.load('glue')
    out = tls.quantize_per_map(g['quantize__data'], g['quantize__bin_widths'])
    assert out.dtype == numpy.float32 and numpy.array_equal(out, g['quantize__out'])
    out = tls.quantize_per_map(g['quantize__halves'], numpy.ones(8, dtype=numpy.float32))
    assert numpy.array_equal(out, g['quantize__halves_out'])        # round half to even
    rng = numpy.random.default_rng(0)
    data = (rng.standard_cauchy(size=(2, 32, 48, 128))*4).astype(numpy.float32)
    bw = (0.8 + 3.2*rng.random(128)).astype(numpy.float32)
    assert numpy.array_equal(tls.quantize_per_map(data, bw), oracle_glue.quantize_per_map(data, bw))


def test_casts(tls, golden):
    g = golden.load('glue')
    assert numpy.array_equal(tls.cast_bt601(g['cast_bt601__in']), g['cast_bt601__out'])
    assert numpy.array_equal(tls.cast_float_to_int16(g['cast_int16__in']), g['cast_int16__out'])
    with pytest.raises(AssertionError):     # tools.py:126-133
        tls.cast_float_to_int16(numpy.array([1., 32768.2], dtype=numpy.float32))
    rng = numpy.random.default_rng(1)
    x = (rng.normal(120., 90., size=(3, 64, 48, 1))).astype(numpy.float32)
    assert numpy.array_equal(tls.cast_bt601(x), oracle_glue.cast_bt601(x))


def test_psnr_entropy_rate_deads(tls, golden):
    g = golden.load('glue')
    assert abs(tls.psnr_2d(g['psnr__a'], g['psnr__b']) - float(g['psnr__out'])) < 1e-12
    twelve = 12*numpy.ones((2, 2), dtype=numpy.uint8)
    assert abs(tls.psnr_2d(twelve, 15*numpy.ones((2, 2), dtype=numpy.uint8)) - 38.5883785143) < 1e-9
    with pytest.raises(ValueError):         # tools.py:879-880
        tls.psnr_2d(twelve, twelve)
    q = g['quantize__out']
    bw = g['quantize__bin_widths']
    assert numpy.array_equal(tls.count_symbols(q[0, :, :, 0], float(bw[0])), g['count_symbols__out0'])
    ent = numpy.array([tls.discrete_entropy(q[0, :, :, i], float(bw[i])) for i in range(0, 128, 8)])
    assert numpy.allclose(ent, g['entropy__out'][::8], rtol=0, atol=1e-12)
    assert abs(tls.rate_3d(q[0], bw, 192, 320) - float(g['rate_3d__out'])) < 1e-12
    assert numpy.array_equal(tls.count_nb_deads(g['nb_deads__in']), g['nb_deads__out'])
    assert tls.count_nb_deads(numpy.zeros((2, 3, 3, 4), dtype=numpy.float32)).tolist() == [4, 4]
    with pytest.raises(AssertionError):     # "The quantization was omitted." (tools.py:372-375)
        tls.rate_3d((q[0] + 0.3).astype(numpy.float32), bw, 192, 320)


@pytest.fixture(scope='module')
def stats(native):
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import stats
    return stats


def test_statistics_match_the_reference_outputs(stats, golden):
    """lossless.stats on the GPU against the reference's own outputs on the same seeded latents (glue.npz):
    tables bit-identical, same exception map, known answers of test_lossless.py:267-298."""
    g = golden.load('glue')
    data = g['quantize__data']
    bw = g['quantize__bin_widths']
    table = stats.compute_binary_probabilities(data, bw, g['binary_probabilities__mean'], 10)
    assert table.dtype == numpy.float64 and numpy.array_equal(table, g['binary_probabilities__out'])
    assert stats.find_index_map_exception(data) == int(g['idx_map_exception__out'])
    (z, o) = stats.count_binary_decisions(numpy.array([0.75, 0.05, 0.1, 0.2, 0.2, 0.15], dtype=numpy.float32), 0.05, 7)
    assert numpy.array_equal(numpy.stack([z, o]), g['binary_decisions__0'])
    (z, o) = stats.count_binary_decisions(numpy.array([210., 6., 9., 6.], dtype=numpy.float32), 3., 7)
    assert numpy.array_equal(numpy.stack([z, o]), g['binary_decisions__1'])
    with pytest.raises(ValueError):
        stats.count_binary_decisions(numpy.array([1., -1.], dtype=numpy.float32), 1., 4)


def test_statistics_against_the_oracle_on_wide_and_degenerate_maps(stats):
    rng = numpy.random.default_rng(5)
    y = (rng.laplace(0., 2., size=(5, 16, 24, 128))*rng.uniform(0.05, 30., size=128)).astype(numpy.float32)
    y[..., 3] = rng.uniform(-40., 40., size=y.shape[:3])         # the near-uniform map
    y[..., 7] = 0.25                                             # a constant map inside one unit interval
    y[..., 9] = rng.integers(-3, 4, size=y.shape[:3])            # values on the interval edges, maximum on the right edge
    mean = numpy.mean(y, axis=(0, 1, 2)).astype(numpy.float32)
    bw = rng.uniform(0.3, 4., size=128).astype(numpy.float32)
    for L in (1, 10, 40):
        cq = oracle_glue.quantize_per_map(y - mean.reshape((1, 1, 1, -1)), bw)
        want = numpy.stack([oracle_glue.binary_probabilities_from_counts(
            *oracle_glue.count_binary_decisions(numpy.absolute(cq[..., i]), float(bw[i]), L)) for i in range(128)])
        assert numpy.array_equal(stats.compute_binary_probabilities(y, bw, mean, L), want), L
    assert stats.find_index_map_exception(y) == oracle_glue.find_index_map_exception(y) == 3
    (edges, probs) = stats.compute_probabilities_intervals(y[..., 9], 1.)
    (want_edges, want_probs) = oracle_glue.probabilities_unit_intervals(y[..., 9])
    assert numpy.array_equal(edges, want_edges) and numpy.array_equal(probs, want_probs)
    with pytest.raises(ValueError):      # stats.py:105-106: all values equal to one integer
        stats.find_index_map_exception(numpy.full((1, 2, 2, 128), 2., dtype=numpy.float32))


def test_save_statistics_writes_the_three_kinds_of_files(stats, native, tmp_path):
    from autoencoder_based_image_compression_b200 import codec as native_codec
    from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.EntropyAutoencoder import EntropyAutoencoder
    from oracle import transforms as oracle_transforms
    from tests import util
    rng = numpy.random.default_rng(11)
    lum = util.synthetic_luma(rng, 4, 64, 96)[..., None]
    multipliers = numpy.array([1., 2.], dtype=numpy.float32)
    paths = [str(tmp_path/'binary_probabilities_{}.npy'.format(i)) for i in range(2)]
    (p_mean, p_idx) = (str(tmp_path/'map_mean.npy'), str(tmp_path/'idx_map_exception.pkl'))
    entropy_ae = EntropyAutoencoder(2, 64, 96, 1., 10000., '', False)
    with native_codec.Session(device=0, math='tf32x3') as sess:
        entropy_ae.initialization(sess, '')
        stats.save_statistics(lum, sess, entropy_ae, 2, multipliers, 10, p_mean, p_idx, paths)
        y = __import__('autoencoder_based_image_compression_b200.kodak_tensorflow.eae.batching', fromlist=['x']) \
            .encode_mini_batches(lum, sess, entropy_ae, 2)
    import pickle
    map_mean = numpy.load(p_mean)
    assert map_mean.dtype == numpy.float32 and numpy.array_equal(map_mean, numpy.mean(y, axis=(0, 1, 2)))
    with open(p_idx, 'rb') as f:
        assert pickle.load(f) == oracle_glue.find_index_map_exception(y)
    for (i, path) in enumerate(paths):
        table = numpy.load(path)
        assert table.shape == (128, 10) and table.dtype == numpy.float64
        assert table.min() >= 0.01 and table.max() <= 0.99
        bw = multipliers[i]*entropy_ae.get_bin_widths()
        cq = oracle_glue.quantize_per_map(y - map_mean.reshape((1, 1, 1, -1)), bw)
        want = numpy.stack([oracle_glue.binary_probabilities_from_counts(
            *oracle_glue.count_binary_decisions(numpy.absolute(cq[..., j]), float(bw[j]), 10)) for j in range(128)])
        assert numpy.array_equal(table, want)
    with pytest.raises(ValueError):      # stats.py:296-297
        stats.save_statistics(lum, None, entropy_ae, 2, multipliers, 10, p_mean, p_idx, paths[:1])


@pytest.mark.parametrize('per_image', [1, 0])
def test_warp_level_histograms_of_planar_streams(native, per_image):
    """eae_histogram_streams_dev (one warp per stream, 16-byte loads, shared-memory counters) against numpy, on stream
    lengths that are and are not multiples of the vector width (so streams start at every 2-byte alignment), a range
    wider than one shared-memory histogram (global-atomics kernel) and a constant map. tools.py:322-388."""
    import torch
    lib = native.lib()
    rng = numpy.random.default_rng(31)
    dev = torch.device('cuda', 0)
    for (n_images, size, C, spread) in ((3, 1536, 128, 6.), (2, 1537, 5, 40.), (4, 3, 7, 2.), (1, 32400, 16, 3000.), (2, 77, 4, 0.)):
        idx = numpy.round(rng.laplace(0., 1., size=(n_images, C, size))*spread).clip(-32768, 32767).astype(numpy.int16)
        d_idx = torch.from_numpy(idx).to(dev)
        n_hist = n_images*C if per_image else C
        d_mn = torch.zeros(n_hist, dtype=torch.int32, device=dev)
        d_mx = torch.zeros(n_hist, dtype=torch.int32, device=dev)
        d_abs = torch.zeros(n_hist, dtype=torch.int64, device=dev)
        native.check(lib.eae_histogram_streams_dev(d_idx.data_ptr(), n_images, size, C, per_image, d_mn.data_ptr(),
                                                   d_mx.data_ptr(), d_abs.data_ptr(), None, 0, None))
        grouped = idx.reshape(n_hist, size) if per_image else idx.transpose(1, 0, 2).reshape(C, n_images*size)
        assert numpy.array_equal(d_mn.cpu().numpy(), grouped.min(axis=1))
        assert numpy.array_equal(d_mx.cpu().numpy(), grouped.max(axis=1))
        assert numpy.array_equal(d_abs.cpu().numpy(), numpy.abs(grouped.astype(numpy.int64)).sum(axis=1))
        cap = int((grouped.max(axis=1).astype(numpy.int64) - grouped.min(axis=1)).max()) + 1
        d_hist = torch.full((n_hist, cap), -1, dtype=torch.int64, device=dev)
        native.check(lib.eae_histogram_streams_dev(d_idx.data_ptr(), n_images, size, C, per_image, d_mn.data_ptr(),
                                                   None, None, d_hist.data_ptr(), cap, None))
        hist = d_hist.cpu().numpy()
        for j in range(n_hist):
            want = numpy.bincount(grouped[j].astype(numpy.int64) - int(grouped[j].min()), minlength=cap)
            assert numpy.array_equal(hist[j], want), (n_images, size, C, j)

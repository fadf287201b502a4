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

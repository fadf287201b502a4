"""Parity of the GPU analysis / synthesis transforms with the CPU restatement (oracle/transforms.py),
through the reference-facing API (EntropyAutoencoder / IsolatedDecoder / eae.batching).

Tolerances (BASELINE.json north_star): quantization indices agree on >= 99.99 % of the coefficients
and every mismatch sits next to a bin boundary; reconstruction PSNR within 0.01 dB."""
import numpy
import pytest
import torch

from autoencoder_based_image_compression_b200 import codec as native_codec
from autoencoder_based_image_compression_b200 import weights as wts
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae import batching
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.EntropyAutoencoder import EntropyAutoencoder
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.IsolatedDecoder import IsolatedDecoder
from oracle import glue as oracle_glue
from oracle import transforms as T
from tests import util

pytestmark = pytest.mark.gpu

MATH_MODES = ['fp32']


def index_agreement(y_gpu, y_ref64, delta=1.0):
    """Fraction of equal indices and the largest distance of a mismatching coefficient to its bin edge."""
    k_gpu = numpy.round(y_gpu.astype(numpy.float64)/delta)
    k_ref = numpy.round(y_ref64/delta)
    diff = k_gpu != k_ref
    frac = 1. - diff.mean()
    edge = numpy.abs(numpy.abs(y_ref64/delta - numpy.floor(y_ref64/delta)) - 0.5)
    worst = float(edge[diff].max()) if diff.any() else 0.
    return (frac, worst, int(diff.sum()))


@pytest.mark.parametrize('math', MATH_MODES)
@pytest.mark.parametrize('learned', [False, True])
def test_encoder_decoder_small(native, learned, math):
    rng = numpy.random.default_rng(0)
    w = wts.random_init(0, learned)
    (n, h, wd) = (4, 64, 96)
    lum = util.synthetic_luma(rng, n, h, wd, smooth=False)[..., None]
    sess = native_codec.Session(0, math=math)
    ae = EntropyAutoencoder(4, h, wd, 1., 10000., '', learned)
    ae.set_weights(w)
    y = batching.encode_mini_batches(lum, sess, ae, 4)
    assert y.shape == (n, h//16, wd//16, 128) and y.dtype == numpy.float32
    y32 = T.encoder(lum.astype(numpy.float32), w, learned)
    y64 = T.encoder(lum.astype(numpy.float64), w, learned, dtype=torch.float64)
    scale = numpy.abs(y64).max()
    assert numpy.abs(y - y64).max() <= 4.*max(numpy.abs(y32 - y64).max(), 1e-6*scale)
    (frac, worst, nb) = index_agreement(y, y64)
    assert frac >= 0.9999 and worst < 1e-4, (frac, worst, nb)
    # decoder on the oracle's quantized latent
    q = oracle_glue.quantize_per_map(y32, numpy.ones(128, dtype=numpy.float32))
    dec = IsolatedDecoder(4, h, wd, learned)
    dec.set_weights(w)
    rec = batching.decode_mini_batches(q, sess, dec, 4)
    assert rec.shape == (n, h, wd, 1) and rec.dtype == numpy.uint8
    rec64 = T.decoder(q.astype(numpy.float64), w, learned, dtype=torch.float64)
    want = oracle_glue.cast_bt601(rec64)
    delta = numpy.abs(rec.astype(numpy.int32) - want.astype(numpy.int32))
    assert delta.max() <= 1 and (delta != 0).mean() < 1e-3
    assert rec.min() >= 16 and rec.max() <= 235


def test_all_zero_latent_gives_constant_reconstruction(native):
    # test_eae.py:249-294: a zero latent decodes to a constant image (no bias in the last layer)
    w = wts.random_init(1, False)
    dec = IsolatedDecoder(2, 64, 64, False)
    dec.set_weights(w)
    rec = batching.decode_mini_batches(numpy.zeros((2, 4, 4, 128), dtype=numpy.float32), native_codec.Session(0), dec, 2)
    assert rec.min() == rec.max() == 16     # reconstruction 0.0 clipped to the BT.601 floor


def test_kodak_size_image_against_oracle(native):
    """BASELINE config 1: one 512 x 768 image, delta = 1, fixed-delta variant (6 GDN/IGDN)."""
    rng = numpy.random.default_rng(1)
    w = wts.random_init(0, False)
    lum = util.synthetic_luma(rng, 1, 512, 768)[..., None]
    codec = native_codec.Codec(w, False)
    y = codec.encode(lum)
    y64 = T.encoder(lum.astype(numpy.float64), w, False, dtype=torch.float64)
    y32 = T.encoder(lum.astype(numpy.float32), w, False)
    (frac, worst, nb) = index_agreement(y, y64)
    (frac32, _, nb32) = index_agreement(y32, y64)
    assert frac >= 0.9999 and worst < 1e-4, (frac, worst, nb, frac32, nb32)
    q = oracle_glue.quantize_per_map(y32, numpy.ones(128, dtype=numpy.float32))
    rec = codec.decode(q)[..., 0]
    want = oracle_glue.cast_bt601(T.decoder(q, w, False))[..., 0]
    psnr_gpu = oracle_glue.psnr_2d(lum[0, :, :, 0], rec[0])
    psnr_ref = oracle_glue.psnr_2d(lum[0, :, :, 0], want[0])
    assert abs(psnr_gpu - psnr_ref) < 0.01
    assert (rec != want).mean() < 1e-3


def test_non_multiple_tile_sizes_and_4k_frame_shape(native):
    """Latent grids that are not multiples of the 128-row GEMM tile, incl. the 4K frame (135 x 240)."""
    rng = numpy.random.default_rng(2)
    w = wts.random_init(2, True)
    codec = native_codec.Codec(w, True)
    for (h, wd) in ((16, 16), (48, 80), (2160, 3840)):
        lum = util.synthetic_luma(rng, 1, h, wd)[..., None]
        y = codec.encode(lum)
        y32 = T.encoder(lum.astype(numpy.float32), w, True)
        assert y.shape == y32.shape == (1, h//16, wd//16, 128)
        tol = 2e-4*max(1., numpy.abs(y32).max())
        assert numpy.abs(y - y32).max() < tol
        q = numpy.round(y32)
        rec = codec.decode(q)
        want = oracle_glue.cast_bt601(T.decoder(q, w, True))
        assert (rec != want).mean() < 1e-3

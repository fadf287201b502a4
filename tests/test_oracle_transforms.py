"""Cross-checks the torch-CPU restatement of the transforms (oracle/transforms.py) against naive
direct convolutions written from the definitions, and against the few facts the reference's tests
state about the graph (shapes, gamma = 0 GDN)."""
import numpy
import torch

from autoencoder_based_image_compression_b200 import weights as wts
from oracle import transforms as T


def test_same_padding_rule():
    assert T.same_pads(512, 9, 4) == (2, 3)      # total 5 -> 2 before, 3 after
    assert T.same_pads(128, 5, 2) == (1, 2)      # total 3 -> 1 before, 2 after
    assert T.same_pads(96, 5, 2) == (1, 2)


def test_conv_and_transpose_against_naive():
    rng = numpy.random.default_rng(0)
    for (k, s, cin, cout, h, w) in ((9, 4, 1, 5, 16, 24), (5, 2, 4, 6, 10, 14), (5, 2, 3, 3, 6, 8)):
        x = rng.normal(size=(2, h, w, cin))
        f = rng.normal(size=(k, k, cin, cout))
        got = T.conv2d_same(torch.from_numpy(x), torch.from_numpy(f), s).numpy()
        assert numpy.allclose(got, T.naive_conv2d_same(x, f, s), atol=1e-10)
        # transpose: filter [k, k, out, in]
        z = rng.normal(size=(2, h//s, w//s, cout))
        ft = rng.normal(size=(k, k, cin, cout))
        got = T.conv2d_transpose_same(torch.from_numpy(z), torch.from_numpy(ft), s).numpy()
        want = T.naive_conv2d_transpose_same(z, ft, s)
        assert got.shape == (2, h, w, cin)
        assert numpy.allclose(got, want, atol=1e-10)


def test_transpose_is_the_adjoint_of_conv():
    # <conv(x), z> == <x, conv_transpose(z)> with the same filter: tf.nn.conv2d_transpose is defined
    # as the gradient of conv2d.
    rng = numpy.random.default_rng(1)
    for (k, s) in ((9, 4), (5, 2)):
        x = torch.from_numpy(rng.normal(size=(1, 16, 24, 3)))
        f = torch.from_numpy(rng.normal(size=(k, k, 3, 4)))
        z = torch.from_numpy(rng.normal(size=(1, 16//s, 24//s, 4)))
        lhs = (T.conv2d_same(x, f, s)*z).sum()
        rhs = (x*T.conv2d_transpose_same(z, f, s)).sum()
        assert abs(float(lhs - rhs)) < 1e-9


def test_gdn():
    rng = numpy.random.default_rng(2)
    x = rng.normal(size=(2, 3, 4, 8))
    gamma = rng.uniform(2e-5, 0.01, size=(8, 8))
    beta = rng.uniform(0.5, 2., size=8)
    for inverse in (False, True):
        got = T.gdn(torch.from_numpy(x), torch.from_numpy(gamma), torch.from_numpy(beta), inverse).numpy()
        assert numpy.allclose(got, T.naive_gdn(x, gamma, beta, inverse), atol=1e-12)
    # test_tfutils.py:398-423, 493-518: gamma = 0, beta = 4 -> GDN(x) = x/2, IGDN(x) = 2x
    zero = torch.zeros((8, 8), dtype=torch.float64)
    four = 4.*torch.ones(8, dtype=torch.float64)
    assert numpy.allclose(T.gdn(torch.from_numpy(x), zero, four).numpy(), x/2.)
    assert numpy.allclose(T.gdn(torch.from_numpy(x), zero, four, inverse=True).numpy(), 2.*x)


def test_graph_shapes_and_variants():
    # test_eae.py:71-139, 178-247: [2, 96, 128, 1] -> [2, 6, 8, C] -> [2, 96, 128, 1]
    for learned in (False, True):
        w = wts.random_init(3, learned)
        x = numpy.random.default_rng(3).integers(0, 256, size=(2, 96, 128, 1)).astype(numpy.float32)
        y = T.encoder(x, w, learned)
        assert y.shape == (2, 6, 8, 128) and y.dtype == numpy.float32
        r = T.decoder(y, w, learned)
        assert r.shape == (2, 96, 128, 1)
        y64 = T.encoder(x, w, learned, dtype=torch.float64)
        assert numpy.abs(y64 - y).max() < 1e-3*max(1., numpy.abs(y64).max())
    assert ('encoder/gamma_3' in wts.random_init(0, False)) and ('encoder/gamma_3' not in wts.random_init(0, True))

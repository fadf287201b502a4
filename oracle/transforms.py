"""TEST INFRASTRUCTURE ONLY: CPU restatement of the EAE analysis / synthesis transforms.

PARITY UNPINNED BY THE REFERENCE: TensorFlow is not installable offline and the trained checkpoints
are absent (/root/reference/.MISSING_LARGE_BLOBS), and the reference's own tests pin only output
SHAPES (kodak_tensorflow/test_eae.py:71-139,178-247) and the trivial gamma=0 GDN case
(test_tfutils.py:398-423,493-518). This file restates

* ``components.encoder`` / ``components.decoder`` (kodak_tensorflow/eae/graph/components.py:11-142),
* ``tfutils.gdn`` / ``tfutils.inverse_gdn`` (kodak_tensorflow/tfutils/tfutils.py:363-397, 480-509),
* TensorFlow's SAME padding rule for ``tf.nn.conv2d`` and the exact-adjoint rule for
  ``tf.nn.conv2d_transpose`` (third-party, version unpinned: README.md:12 "0.11.0 / 1.4.0"),

with PyTorch-CPU ops in fp32 or fp64, and is itself cross-checked against a naive numpy direct
convolution (``naive_*`` below) in tests/test_oracle_transforms.py.

Weights are a dict keyed by the TF variable names (EntropyAutoencoder.py:108-224), TF layouts:
conv filters ``[kh, kw, in, out]``, transposed-conv filters ``[kh, kw, out, in]``, gamma ``[in j, out i]``.
"""
import numpy
import torch
import torch.nn.functional as F

STRIDES = (4, 2, 2)  # constants.py:54-56


def same_pads(size_in, kernel, stride):
    """TF SAME: out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total//2."""
    size_out = -(-size_in//stride)
    total = max((size_out - 1)*stride + kernel - size_in, 0)
    return (total//2, total - total//2)


def _t(a, dtype):
    return torch.from_numpy(numpy.ascontiguousarray(a)).to(dtype)


def conv2d_same(x_nhwc, w_hwio, stride):
    """tf.nn.conv2d(x, w, [1,s,s,1], 'SAME') (cross-correlation)."""
    (kh, kw, _, _) = w_hwio.shape
    (pt, pb) = same_pads(x_nhwc.shape[1], kh, stride)
    (pl, pr) = same_pads(x_nhwc.shape[2], kw, stride)
    x = F.pad(x_nhwc.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(x, w_hwio.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_transpose_same(x_nhwc, w_hwoi, stride):
    """tf.nn.conv2d_transpose(x, w[kh,kw,out,in], [N, s*h, s*w, out], [1,s,s,1], 'SAME').

    Exact adjoint of ``conv2d_same`` from the (s*h, s*w) grid: full output (in-1)*s + k cropped
    at offset ``before`` of the forward padding.
    """
    (kh, kw, _, _) = w_hwoi.shape
    (h, w) = (x_nhwc.shape[1], x_nhwc.shape[2])
    full = F.conv_transpose2d(x_nhwc.permute(0, 3, 1, 2), w_hwoi.permute(3, 2, 0, 1), stride=stride)
    (pt, _) = same_pads(stride*h, kh, stride)
    (pl, _) = same_pads(stride*w, kw, stride)
    y = full[:, :, pt:pt + stride*h, pl:pl + stride*w]
    return y.permute(0, 2, 3, 1).contiguous()


def gdn(x_nhwc, gamma, beta, inverse=False):
    """x / sqrt(x^2 @ gamma + beta)  (inverse: x * sqrt(...)); gamma indexed [input j, output i]."""
    shape = x_nhwc.shape
    x2d = x_nhwc.reshape(-1, shape[3])
    norm = torch.sqrt(torch.matmul(x2d*x2d, gamma) + beta.reshape(1, -1))
    out = x2d*norm if inverse else x2d/norm
    return out.reshape(shape)


# The layer tables below ARE what encoder() / decoder() execute. tests/test_graph_structure.py compares them, entry by
# entry, with the op chain, the strides / padding / data format and the variable names and shapes of the reference's own
# serialised graphs (kodak_tensorflow/eae/results/*/model_*.ckpt.meta, parsed into tests/golden/graph_structure.json):
# the STRUCTURE of this restatement is pinned to a reference-held artefact, its arithmetic is restated.
# `optional`: the non-linearity exists only when the bin widths are NOT learned (components.py:56-58, 138-141).
ENCODER_LAYERS = (
    {'conv': 'encoder/weights_1', 'stride': STRIDES[0], 'bias': 'encoder/biases_1',
     'gdn': ('encoder/gamma_1', 'encoder/beta_1'), 'optional': False},
    {'conv': 'encoder/weights_2', 'stride': STRIDES[1], 'bias': 'encoder/biases_2',
     'gdn': ('encoder/gamma_2', 'encoder/beta_2'), 'optional': False},
    {'conv': 'encoder/weights_3', 'stride': STRIDES[2], 'bias': 'encoder/biases_3',
     'gdn': ('encoder/gamma_3', 'encoder/beta_3'), 'optional': True},
)
# The inverse GDN comes BEFORE the transposed convolution of the same index (components.py:56-84).
DECODER_LAYERS = (
    {'igdn': ('decoder/gamma_4', 'decoder/beta_4'), 'optional': True,
     'tconv': 'decoder/weights_4', 'stride': STRIDES[2], 'bias': 'decoder/biases_4'},
    {'igdn': ('decoder/gamma_5', 'decoder/beta_5'), 'optional': False,
     'tconv': 'decoder/weights_5', 'stride': STRIDES[1], 'bias': 'decoder/biases_5'},
    {'igdn': ('decoder/gamma_6', 'decoder/beta_6'), 'optional': False,
     'tconv': 'decoder/weights_6', 'stride': STRIDES[0], 'bias': None},      # no bias: components.py:79-84
)


def encoder_stages(visible_units_nhwc, weights, are_bin_widths_learned, dtype=torch.float32):
    """Per-layer outputs [gdn_1, gdn_2, y] of components.encoder (components.py:86-142). Input [B,h,w,1], raw 0..255."""
    w = {k: _t(v, dtype) for (k, v) in weights.items() if k.startswith('encoder/')}
    x = _t(visible_units_nhwc, dtype)
    stages = []
    for layer in ENCODER_LAYERS:
        x = conv2d_same(x, w[layer['conv']], layer['stride']) + w[layer['bias']]
        if not (layer['optional'] and are_bin_widths_learned):
            x = gdn(x, w[layer['gdn'][0]], w[layer['gdn'][1]])
        stages.append(x.numpy())
    return stages


def encoder(visible_units_nhwc, weights, are_bin_widths_learned, dtype=torch.float32):
    """components.encoder (components.py:86-142). Input float array [B,h,w,1], raw 0..255."""
    return encoder_stages(visible_units_nhwc, weights, are_bin_widths_learned, dtype)[-1]


def decoder(quantized_y_nhwc, weights, are_bin_widths_learned, dtype=torch.float32):
    """components.decoder (components.py:11-84). The last layer has no bias (:79-84)."""
    w = {k: _t(v, dtype) for (k, v) in weights.items() if k.startswith('decoder/')}
    x = _t(quantized_y_nhwc, dtype)
    for layer in DECODER_LAYERS:
        if not (layer['optional'] and are_bin_widths_learned):
            x = gdn(x, w[layer['igdn'][0]], w[layer['igdn'][1]], inverse=True)
        x = conv2d_transpose_same(x, w[layer['tconv']], layer['stride'])
        if layer['bias'] is not None:
            x = x + w[layer['bias']]
    return x.numpy()


# ---------------------------------------------------------------------------------------------
# Naive numpy direct forms, written from the definitions, used to validate the torch restatement.

def naive_conv2d_same(x_nhwc, w_hwio, stride):
    x = numpy.asarray(x_nhwc, dtype=numpy.float64)
    w = numpy.asarray(w_hwio, dtype=numpy.float64)
    (n, h, wd, _) = x.shape
    (kh, kw, _, co) = w.shape
    (ho, wo) = (-(-h//stride), -(-wd//stride))
    (pt, _) = same_pads(h, kh, stride)
    (pl, _) = same_pads(wd, kw, stride)
    y = numpy.zeros((n, ho, wo, co))
    for a in range(ho):
        for b in range(wo):
            for ky in range(kh):
                iy = a*stride + ky - pt
                if iy < 0 or iy >= h:
                    continue
                for kx in range(kw):
                    ix = b*stride + kx - pl
                    if ix < 0 or ix >= wd:
                        continue
                    y[:, a, b, :] += x[:, iy, ix, :] @ w[ky, kx]
    return y


def naive_conv2d_transpose_same(x_nhwc, w_hwoi, stride):
    """Adjoint written as a scatter: forward tap (a, ky) reads position a*s + ky - before."""
    x = numpy.asarray(x_nhwc, dtype=numpy.float64)
    w = numpy.asarray(w_hwoi, dtype=numpy.float64)
    (n, h, wd, _) = x.shape
    (kh, kw, co, _) = w.shape
    (hh, ww) = (stride*h, stride*wd)
    (pt, _) = same_pads(hh, kh, stride)
    (pl, _) = same_pads(ww, kw, stride)
    y = numpy.zeros((n, hh, ww, co))
    for a in range(h):
        for b in range(wd):
            for ky in range(kh):
                oy = a*stride + ky - pt
                if oy < 0 or oy >= hh:
                    continue
                for kx in range(kw):
                    ox = b*stride + kx - pl
                    if ox < 0 or ox >= ww:
                        continue
                    y[:, oy, ox, :] += x[:, a, b, :] @ w[ky, kx].T
    return y


def naive_gdn(x_nhwc, gamma, beta, inverse=False):
    x = numpy.asarray(x_nhwc, dtype=numpy.float64)
    norm = numpy.sqrt(numpy.einsum('nhwj,ji->nhwi', x*x, numpy.asarray(gamma, dtype=numpy.float64)) + beta)
    return x*norm if inverse else x/norm

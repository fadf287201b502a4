"""Inference half of the reference's ``EntropyAutoencoder`` class
(kodak_tensorflow/eae/graph/EntropyAutoencoder.py:36-251, 398-463): same constructor arguments,
``initialization`` and ``get_bin_widths``; the TensorFlow graph is replaced by the native codec."""
import numpy

from autoencoder_based_image_compression_b200 import codec as native_codec
from autoencoder_based_image_compression_b200 import weights as wts
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph import constants as csts


class _Model(object):
    def __init__(self, batch_size, h_in, w_in, are_bin_widths_learned, what):
        if h_in % csts.STRIDE_PROD != 0:
            raise ValueError('The height of the {} is not divisible by the product of the three strides.'.format(what))
        if w_in % csts.STRIDE_PROD != 0:
            raise ValueError('The width of the {} is not divisible by the product of the three strides.'.format(what))
        self.batch_size = batch_size
        self.h_in = h_in
        self.w_in = w_in
        self.are_bin_widths_learned = are_bin_widths_learned
        self.weights = None
        self._codec = None

    def check_input(self, h_in, w_in):
        # The TF placeholders have a static shape (EntropyAutoencoder.py:248-249).
        if (h_in, w_in) != (self.h_in, self.w_in):
            raise ValueError('Cannot feed images of shape ({}, {}): the graph was built for ({}, {}).'.format(
                h_in, w_in, self.h_in, self.w_in))

    def initialization(self, sess, path_to_restore, seed=0):
        """EntropyAutoencoder.py:440-463. ``path_to_restore`` names an ``.npz`` written by ``weights.save`` (keys = TF
        variable names): the file itself, or the reference's ``.../model_k.ckpt`` prefix next to which
        ``model_k.ckpt.npz`` / ``model_k.npz`` lies (so the reference's drivers run with their own paths).
        An empty string draws a random initialisation."""
        if path_to_restore:
            self.weights = wts.load(wts.resolve_path(path_to_restore))
        else:
            self.weights = wts.random_init(seed, self.are_bin_widths_learned, getattr(self, 'bin_width_init', 1.))
        self._codec = None

    def set_weights(self, weights):
        self.weights = dict(weights)
        self._codec = None

    def codec(self, sess):
        if self.weights is None:
            raise RuntimeError('Attempting to use uninitialized value: call `initialization` first.')
        device = getattr(sess, 'device', 0) if sess is not None else 0
        math = getattr(sess, 'math', 'fp32') if sess is not None else 'fp32'
        if self._codec is None or self._codec.device != device:
            full = dict(self.weights)
            # The native handle holds both transforms; a model restored with only one half gets
            # zeros for the other, which it never runs.
            for key in wts.ENCODER_KEYS + wts.DECODER_KEYS:
                if key not in full and not (key in wts.OPTIONAL_KEYS and self.are_bin_widths_learned):
                    short = key.split('/')[1]
                    shape = wts.SHAPES.get(short, (128, 128) if short.startswith('gamma') else (128,))
                    full[key] = numpy.ones(shape, dtype=numpy.float32) if short.startswith(('gamma', 'beta')) \
                        else numpy.zeros(shape, dtype=numpy.float32)
            self._codec = native_codec.Codec(full, self.are_bin_widths_learned, device=device, math=math)
        else:
            self._codec.set_math(math)
        return self._codec


class EntropyAutoencoder(_Model):
    """Entropy autoencoder (inference). See EntropyAutoencoder.py:18-35."""

    def __init__(self, batch_size, h_in, w_in, bin_width_init, gamma_scaling, path_to_nb_itvs_per_side_load,
                 are_bin_widths_learned):
        _Model.__init__(self, batch_size, h_in, w_in, are_bin_widths_learned, 'input images')
        self.bin_width_init = bin_width_init
        # Training-only arguments, kept for signature compatibility.
        self.gamma_scaling = gamma_scaling
        self.path_to_nb_itvs_per_side_load = path_to_nb_itvs_per_side_load

    def get_bin_widths(self):
        """EntropyAutoencoder.py:398-409: float32 [128]."""
        if self.weights is None:
            raise RuntimeError('Attempting to use uninitialized value: call `initialization` first.')
        if wts.BIN_WIDTHS_KEY in self.weights:
            return numpy.asarray(self.weights[wts.BIN_WIDTHS_KEY], dtype=numpy.float32).copy()
        return (self.bin_width_init*numpy.ones(csts.NB_MAPS_3)).astype(numpy.float32)

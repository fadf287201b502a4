"""The reference's ``IsolatedDecoder`` (kodak_tensorflow/eae/graph/IsolatedDecoder.py:10-129): the
decoder alone, restoring the same variable names under ``decoder/``."""
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.EntropyAutoencoder import _Model


class IsolatedDecoder(_Model):
    """Isolated decoder. ``h_in``/``w_in`` are the size of the images it returns."""

    def __init__(self, batch_size, h_in, w_in, are_bin_widths_learned):
        _Model.__init__(self, batch_size, h_in, w_in, are_bin_widths_learned,
                        'images returned by the isolated decoder')

"""``eae.batching`` with the reference's signatures (kodak_tensorflow/eae/batching.py:11-100); the
two ``sess.run`` calls become launches of the native codec held by the model object."""
import numpy

from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph import constants as csts
from autoencoder_based_image_compression_b200.kodak_tensorflow.tools import tools as tls


def decode_mini_batches(quantized_y_float32, sess, isolated_decoder, batch_size):
    """batching.py:11-54: float32 [N, h/16, w/16, 128] -> uint8 [N, h, w, 1], clipped to [16, 235].

    ``batch_size`` keeps its divisibility contract (tools.py:1130-1131) but no longer bounds the
    launch size: the native side chunks the batch to fit its workspace.
    """
    (nb_images, h_in, w_in, _) = quantized_y_float32.shape
    tls.subdivide_set(nb_images, batch_size)
    isolated_decoder.check_input(h_in*csts.STRIDE_PROD, w_in*csts.STRIDE_PROD)
    return isolated_decoder.codec(sess).decode(quantized_y_float32)


def encode_mini_batches(luminances_uint8, sess, entropy_ae, batch_size):
    """batching.py:56-100: uint8 [N, h, w, 1] -> float32 [N, h/16, w/16, 128] (no input normalisation)."""
    if luminances_uint8.dtype != numpy.uint8:
        raise TypeError('`luminances_uint8.dtype` is not equal to `numpy.uint8`.')
    (nb_images, h_in, w_in, _) = luminances_uint8.shape
    tls.subdivide_set(nb_images, batch_size)
    entropy_ae.check_input(h_in, w_in)
    return entropy_ae.codec(sess).encode(luminances_uint8)

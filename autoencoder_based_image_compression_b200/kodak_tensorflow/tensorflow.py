"""Stand-in for the three TensorFlow names the reference's drivers use around the hot path.

``kodak_tensorflow/reconstructing_eae_kodak.py`` (``fix_gamma`` :135-235, ``vary_gamma_fix_bin_widths`` :499-541) and
``collecting_stats_eae_extra.py`` only touch ``tf.Session()`` (as a context manager whose object is passed to
``eae.batching`` / ``initialization``) and ``tf.reset_default_graph()``. With this directory first on ``sys.path`` the
reference's scripts import unmodified: ``import tensorflow as tf`` resolves here, ``import eae.batching`` etc. to the
sibling packages, and nothing of TensorFlow runs (tests/test_gpu_reference_driver.py drives the reference's own file
that way). The arithmetic of the contractions is chosen by the environment variable ``EAE_MATH``
(``fp32`` | ``tf32x3`` | ``mixed`` | ``tf32``, default ``tf32x3``) and the device by ``EAE_DEVICE``.
"""
import os

from autoencoder_based_image_compression_b200 import codec as _codec

__version__ = '0.0-eae-b200-stand-in'


class Session(_codec.Session):
    """``tf.Session()``: an opaque device / arithmetic handle (codec.Session) accepted in the same positional slot."""

    def __init__(self, *args, **kwargs):
        _codec.Session.__init__(self, device=int(os.environ.get('EAE_DEVICE', '0')),
                                math=os.environ.get('EAE_MATH', 'tf32x3'))


def reset_default_graph():
    """Nothing to destroy: the models hold native codec handles, not graph nodes."""
    return None

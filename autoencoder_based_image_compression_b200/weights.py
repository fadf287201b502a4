"""Weights of the entropy autoencoder: random initialisation and an ``.npz`` format.

Keys are the TensorFlow variable names of the reference graph and arrays keep the TensorFlow
layouts (kodak_tensorflow/eae/graph/EntropyAutoencoder.py:108-224, IsolatedDecoder.py:54-97), so a
checkpoint exported from the reference with ``{v.name[:-2]: sess.run(v)}`` loads unchanged.
The reference's trained checkpoints are not shipped (.MISSING_LARGE_BLOBS), hence the random
initialiser, which reproduces the reference's initial DISTRIBUTIONS (not TensorFlow's RNG stream).
"""
import numpy

from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph import constants as csts

ENCODER_KEYS = ['encoder/weights_1', 'encoder/biases_1', 'encoder/gamma_1', 'encoder/beta_1',
                'encoder/weights_2', 'encoder/biases_2', 'encoder/gamma_2', 'encoder/beta_2',
                'encoder/weights_3', 'encoder/biases_3', 'encoder/gamma_3', 'encoder/beta_3']
DECODER_KEYS = ['decoder/gamma_4', 'decoder/beta_4', 'decoder/weights_4', 'decoder/biases_4',
                'decoder/gamma_5', 'decoder/beta_5', 'decoder/weights_5', 'decoder/biases_5',
                'decoder/gamma_6', 'decoder/beta_6', 'decoder/weights_6']
BIN_WIDTHS_KEY = 'piecewise_linear_function/bin_widths'
OPTIONAL_KEYS = ('encoder/gamma_3', 'encoder/beta_3', 'decoder/gamma_4', 'decoder/beta_4')

SHAPES = {
    'weights_1': (9, 9, 1, 128), 'weights_2': (5, 5, 128, 128), 'weights_3': (5, 5, 128, 128),
    'weights_4': (5, 5, 128, 128), 'weights_5': (5, 5, 128, 128), 'weights_6': (9, 9, 1, 128),
}


def _gamma(rng, nb_maps):
    # tfutils.initialize_weights_gdn (tfutils.py:445-478): symmetrised U[min_gamma, 0.01]
    g = rng.uniform(csts.MIN_GAMMA_BETA, 0.01, size=(nb_maps, nb_maps)).astype(numpy.float32)
    return (0.5*(g + g.T)).astype(numpy.float32)


def random_init(seed=0, are_bin_widths_learned=False, bin_width_init=1.):
    """Random weights with the reference's initial distributions (EntropyAutoencoder.py:131-224)."""
    rng = numpy.random.default_rng(seed)
    w = {}
    stddev = {'weights_1': 0.01, 'weights_2': 0.02, 'weights_3': 0.05,
              'weights_4': 0.05, 'weights_5': 0.02, 'weights_6': 0.01}
    for i in (1, 2, 3):
        name = 'weights_{}'.format(i)
        w['encoder/' + name] = rng.normal(0., stddev[name], size=SHAPES[name]).astype(numpy.float32)
        w['encoder/biases_{}'.format(i)] = numpy.zeros(128, dtype=numpy.float32)
        if i < 3 or not are_bin_widths_learned:
            w['encoder/gamma_{}'.format(i)] = _gamma(rng, 128)
            w['encoder/beta_{}'.format(i)] = numpy.ones(128, dtype=numpy.float32)
    for i in (4, 5, 6):
        name = 'weights_{}'.format(i)
        if i > 4 or not are_bin_widths_learned:
            w['decoder/gamma_{}'.format(i)] = _gamma(rng, 128)
            w['decoder/beta_{}'.format(i)] = numpy.ones(128, dtype=numpy.float32)
        w['decoder/' + name] = rng.normal(0., stddev[name], size=SHAPES[name]).astype(numpy.float32)
        if i < 6:
            w['decoder/biases_{}'.format(i)] = numpy.zeros(128, dtype=numpy.float32)
    w[BIN_WIDTHS_KEY] = (bin_width_init*numpy.ones(128)).astype(numpy.float32)
    return w


def visible_init(seed=0, are_bin_widths_learned=False, bin_width_init=1.):
    """``random_init`` with reconstructions that land inside the BT.601 range instead of being clipped to 16: a
    positive bias before the last IGDN and a larger, positive last filter (the reference's initial distributions give a
    zero-mean output, EntropyAutoencoder.py:131-224). Used wherever a parity check must see un-clipped pixels."""
    w = random_init(seed, are_bin_widths_learned, bin_width_init)
    w['decoder/biases_5'] = (w['decoder/biases_5'] + 2.0).astype(numpy.float32)
    w['decoder/weights_6'] = (numpy.abs(w['decoder/weights_6'])*8.).astype(numpy.float32)
    return w


def validate(weights, are_bin_widths_learned, need_encoder=True, need_decoder=True):
    """Checks presence, dtype and shape of every variable the inference graphs read."""
    keys = (ENCODER_KEYS if need_encoder else []) + (DECODER_KEYS if need_decoder else [])
    for key in keys:
        if key in OPTIONAL_KEYS and are_bin_widths_learned:
            continue
        if key not in weights:
            raise KeyError('missing variable `{}`'.format(key))
        a = weights[key]
        short = key.split('/')[1]
        if short.startswith('weights'):
            expected = SHAPES[short]
        elif short.startswith('gamma'):
            expected = (128, 128)
        else:
            expected = (128,)
        if a.dtype != numpy.float32 or tuple(a.shape) != expected:
            raise ValueError('`{}` must be float32 {}, got {} {}'.format(key, expected, a.dtype, a.shape))


def save(path, weights):
    numpy.savez(path, **{k.replace('/', '.'): v for (k, v) in weights.items()})


def resolve_path(path_to_restore):
    """The ``.npz`` behind a restore path: the path itself, ``<path>.npz`` or, for the reference's ``model_k.ckpt``
    checkpoint prefixes (reconstructing_eae_kodak.py:118), ``model_k.npz``."""
    import os
    candidates = [path_to_restore, path_to_restore + '.npz']
    if path_to_restore.endswith('.ckpt'):
        candidates.append(path_to_restore[:-len('.ckpt')] + '.npz')
    for candidate in candidates:
        if os.path.isfile(candidate):
            return candidate
    raise IOError('no weight file at {} (tried {})'.format(path_to_restore, ', '.join(candidates[1:])))


def load(path):
    with numpy.load(path) as data:
        return {k.replace('.', '/'): data[k] for k in data.files}

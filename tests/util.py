"""Shared helpers of the test-suite (fixtures under tests/golden/, synthetic inputs)."""
import os

import numpy

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Golden(object):
    def __init__(self):
        self._cache = {}

    def load(self, name):
        if name not in self._cache:
            with numpy.load(os.path.join(GOLDEN_DIR, name + '.npz')) as data:
                self._cache[name] = {k: data[k] for k in data.files}
        return self._cache[name]

    def table(self, model='1_10000', mult='1'):
        return self.load('tables')['{}__binary_probabilities_{}'.format(model, mult)]

    def map_mean(self, model='1_10000'):
        return self.load('tables')['{}__map_mean'.format(model)]

    def idx_map_exception(self, model='1_10000'):
        return int(self.load('tables')['{}__idx_map_exception'.format(model)])

    def kat_cases(self):
        kat = self.load('coder_kat')
        names = sorted({k.split('__')[0] for k in kat if k.endswith('__symbols')})
        return [(n, kat[n + '__symbols'], kat[n + '__probs'], kat[n + '__bac'], kat[n + '__byp'],
                 int(kat[n + '__bits'][0]), int(kat[n + '__bits'][1])) for n in names]


def laplace_latent(rng, scale, shape=(32, 48, 128)):
    scales = scale*numpy.exp(rng.normal(0., 0.7, size=shape[-1]))
    x = rng.laplace(0., 1., size=shape)*scales.reshape((1,)*(len(shape) - 1) + (-1,))
    return numpy.round(x).clip(-32767, 32767).astype(numpy.int16)


from autoencoder_based_image_compression_b200.synthetic import synthetic_luma  # noqa: E402,F401

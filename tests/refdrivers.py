"""Imports the reference's own caller scripts UNMODIFIED on top of this package (test infrastructure).

The scripts are read from ``oracle/_ref/drivers`` (staged there, untracked, by ``make -C oracle`` so that they travel to
the GPU box) or straight from ``/root/reference/kodak_tensorflow``. ``sys.path`` gets the mirror directory
``autoencoder_based_image_compression_b200/kodak_tensorflow`` in front, so the scripts' own ``import eae.batching``,
``import lossless.compression``, ``import tools.tools``, ``from eae.graph... import ...`` and ``import tensorflow as tf``
resolve to the B200-backed modules. Only what is OUT OF SCOPE is stubbed: matplotlib (absent from the image) and the
competitor-codec wrappers ``hevc.hevc`` / ``jpeg2000.jpeg2000`` and the CLI validators ``parsing.parsing``
(SURVEY.md section 2, rows 13, 14, 16).
"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIRROR = os.path.join(ROOT, 'autoencoder_based_image_compression_b200', 'kodak_tensorflow')
CANDIDATES = (os.path.join(ROOT, 'oracle', '_ref', 'drivers'), '/root/reference/kodak_tensorflow')


def find(script):
    for directory in CANDIDATES:
        path = os.path.join(directory, script)
        if os.path.isfile(path):
            return path
    return None


class _Anything(types.ModuleType):
    """A module whose every attribute is a callable that returns another such object (pyplot calls)."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Callable()


class _Callable(object):
    def __call__(self, *args, **kwargs):
        return _Callable()

    def __getattr__(self, name):
        return _Callable()


def load(script):
    """Returns the module object of the reference's ``script`` (e.g. 'reconstructing_eae_kodak.py'), or None."""
    path = find(script)
    if path is None:
        return None
    if MIRROR not in sys.path:
        sys.path.insert(0, MIRROR)
    # the mirror packages must win over any same-named top-level package already imported
    for name in ('eae', 'lossless', 'tools', 'tensorflow'):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, '__file__', '').startswith(MIRROR):
            del sys.modules[name]
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker'):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = _Anything(name)
    if isinstance(sys.modules['matplotlib'], _Anything):
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    for (pkg, sub) in (('hevc', 'hevc.hevc'), ('jpeg2000', 'jpeg2000.jpeg2000'), ('parsing', 'parsing.parsing')):
        if sub not in sys.modules:
            sys.modules.setdefault(pkg, types.ModuleType(pkg))
            sys.modules[sub] = types.ModuleType(sub)
            setattr(sys.modules[pkg], sub.split('.')[1], sys.modules[sub])
    spec = importlib.util.spec_from_file_location('reference_' + script[:-3], path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module

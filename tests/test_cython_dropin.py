"""The reference's Cython binding (lossless/interface_cython.pyx), compiled UNMODIFIED against the source-compatible
header include/compat/c++/source/compression.h and linked with libeae_b200.so (recipe: oracle/Makefile, target
_ref/cython_dropin; the generated module is untracked and travels to the GPU box like the other built files).

CPU part: the extension builds, imports and - without a device - fails loudly through the C++ exception the header
throws (Cython's ``except +`` turns std::runtime_error into RuntimeError). GPU part: the reference's own known answers
(test_lossless.py:89-101 -> 20 bits, tests.cpp:354-376 -> 104 bits) and its error behaviour
(test_lossless.py:329-375: RuntimeError "Error of type 4")."""
import importlib.util
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN_DIR = os.path.join(ROOT, 'oracle', '_ref', 'cython_dropin')


def load_dropin():
    if not os.path.isdir(DROPIN_DIR):
        return None
    for name in os.listdir(DROPIN_DIR):
        if name.startswith('interface_cython') and name.endswith('.so'):
            spec = importlib.util.spec_from_file_location('interface_cython', os.path.join(DROPIN_DIR, name))
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            return module
    return None


def test_header_declares_the_reference_signature():
    text = open(os.path.join(ROOT, 'include', 'compat', 'c++', 'source', 'compression.h')).read()
    flat = ' '.join(text.split())
    # lossless/c++/source/compression.h:41-45
    assert ('uint32_t compress_lossless(uint32_t const& size, const int16_t* const array_input, '
            'int16_t* const array_output, uint8_t const& truncated_unary_length, '
            'const double* const probabilities)') in flat
    for needle in ('std::invalid_argument', 'std::runtime_error', 'std::out_of_range', 'One of the three pointers is NULL.'):
        assert needle in text


def test_unmodified_pyx_builds_and_fails_loudly_without_a_device(native):
    module = load_dropin()
    if module is None:
        pytest.skip('oracle/_ref/cython_dropin not built (needs /root/reference at build time)')
    assert callable(module.compress_lossless_flattened_map)
    with pytest.raises(ValueError):      # Cython's own buffer dtype check (interface_cython.pyx:13-14)
        module.compress_lossless_flattened_map(numpy.zeros(4, dtype=numpy.int32), numpy.array([.5]))
    if native.device_count() > 0:
        pytest.skip('a device is present: see the GPU test')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        module.compress_lossless_flattened_map(numpy.array([0, 1, -2], dtype=numpy.int16), numpy.array([.5, .5, .5]))


@pytest.mark.gpu
def test_reference_known_answers_through_the_unmodified_binding(native, golden):
    module = load_dropin()
    if module is None:
        pytest.skip('oracle/_ref/cython_dropin not built (needs /root/reference at build time)')
    (rec, nb_bits) = module.compress_lossless_flattened_map(numpy.array([0, 1, -2, 2, 1, 0, 0, 0], dtype=numpy.int16),
                                                            numpy.array([.5, .5, .5]))
    assert nb_bits == 20 and rec.tolist() == [0, 1, -2, 2, 1, 0, 0, 0]                 # test_lossless.py:89-101
    ref = numpy.array([0, -2, 0, 765, -21, 8, -439, 0, 0, 0, 0, -9], dtype=numpy.int16)
    (rec, nb_bits) = module.compress_lossless_flattened_map(ref, 0.5*numpy.ones(8))
    assert nb_bits == 104 and numpy.array_equal(rec, ref)                              # tests.cpp:354-376
    # a shipped table row on a Laplace map, against the in-tree ctypes binding
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import interface_cython as ctypes_binding
    from tests import util
    table = golden.table()
    latent = util.laplace_latent(numpy.random.default_rng(5), 2.)
    flat = numpy.ascontiguousarray(latent[:, :, 7].flatten())
    (rec_a, bits_a) = module.compress_lossless_flattened_map(flat, numpy.ascontiguousarray(table[7]))
    (rec_b, bits_b) = ctypes_binding.compress_lossless_flattened_map(flat, numpy.ascontiguousarray(table[7]))
    assert bits_a == bits_b and numpy.array_equal(rec_a, flat) and numpy.array_equal(rec_b, flat)
    with pytest.raises(RuntimeError, match='Error of type 4'):                         # test_lossless.py:329-375
        module.compress_lossless_flattened_map(numpy.array([3, 1], dtype=numpy.int16), numpy.array([.5, 1.5, .5]))
    # interface_cython.pyx:50-52 assigns `probabilities.size` to a uint8. Its comment expects Cython to raise; Cython 3
    # reads the size of a typed buffer in C and the assignment wraps (300 -> 44), for the reference's own build as for this
    # one. Either behaviour is the binding's, not the library's: the library must see L <= 255 and answer accordingly.
    two = numpy.array([3, 1], dtype=numpy.int16)
    try:
        (rec_w, bits_w) = module.compress_lossless_flattened_map(two, 0.5*numpy.ones(300))
    except OverflowError:
        pass
    else:
        (rec_44, bits_44) = module.compress_lossless_flattened_map(two, 0.5*numpy.ones(300 & 0xFF))
        assert bits_w == bits_44 and numpy.array_equal(rec_w, rec_44) and numpy.array_equal(rec_w, two)

"""The C-ABI library loads on a machine without a GPU, exports every symbol include/eae_b200.h
declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy
import pytest

from autoencoder_based_image_compression_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'eae_b200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(eae_[a-z0-9_]+)\s*\(', text)))


def test_header_and_bindings_agree():
    names = declared_symbols()
    assert len(names) > 40
    assert sorted(_native.PROTOTYPES) == names


def test_every_declared_symbol_is_exported():
    handle = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), name


def test_no_torch_or_python_dependency():
    import subprocess
    out = subprocess.run(['ldd', _native.LIB_PATH], capture_output=True, text=True).stdout
    assert 'torch' not in out and 'python' not in out and 'libcudart' not in out   # cudart is linked statically


def test_pure_helpers_match_reference_arithmetic():
    lib = _native.lib()
    assert lib.eae_abi_version() == 1
    # compression.cpp:24: size * max(32, L) bits per buffer
    assert lib.eae_coder_capacity_bytes(1536, 10) == 1536*32//8
    assert lib.eae_coder_capacity_bytes(12, 40) == 12*40//8
    assert lib.eae_coder_capacity_bytes(3, 35) == (3*35 + 7)//8
    assert lib.eae_coder_slot_bytes(1536, 10) % 16 == 0
    assert lib.eae_container_bound(1, 512, 768, 10) == 32 + 8*128 + 2*128*6144


@pytest.mark.skipif(_native.device_count() > 0, reason='checks the behaviour WITHOUT a device')
def test_compute_fails_loudly_without_a_gpu():
    lib = _native.lib()
    x = numpy.zeros(8, dtype=numpy.int16)
    out = numpy.zeros(8, dtype=numpy.int16)
    p = numpy.full(3, 0.5)
    nb = ctypes.c_uint32(0)
    code = lib.eae_compress_lossless(8, _native.ptr(x), _native.ptr(out), 3, _native.ptr(p), ctypes.byref(nb))
    assert code == _native.ERR_CUDA
    assert 'no CPU fallback' in _native.last_error()
    with pytest.raises(RuntimeError):
        _native.check(code)
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import interface_cython
    with pytest.raises(RuntimeError):
        interface_cython.compress_lossless_flattened_map(x, p)
    from autoencoder_based_image_compression_b200.kodak_tensorflow.tools import tools as tls
    with pytest.raises(RuntimeError):
        tls.quantize_per_map(numpy.zeros((1, 2, 2, 4), dtype=numpy.float32), numpy.ones(4, dtype=numpy.float32))


def test_argument_errors_come_before_the_device_check():
    lib = _native.lib()
    nb = ctypes.c_uint32(0)
    x = numpy.zeros(8, dtype=numpy.int16)
    p = numpy.full(3, 0.5)
    assert lib.eae_compress_lossless(8, None, _native.ptr(x), 3, _native.ptr(p), ctypes.byref(nb)) == _native.ERR_NULL
    assert lib.eae_compress_lossless(8, _native.ptr(x), _native.ptr(x), 0, _native.ptr(p), ctypes.byref(nb)) == \
        _native.ERR_UNARY_LENGTH

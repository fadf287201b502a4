"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the CPU coder oracles.

Two libraries sit behind the same Python functions:

* ``port``: ``oracle/liboracle_coder.so``, our plain-C restatement (``oracle/coder_oracle.c``).
* ``ref``:  ``oracle/_ref/libref_coder.so``, the reference's own C++ coder compiled in place from
  ``/root/reference/kodak_tensorflow/lossless/c++/source`` behind ``oracle/ref_shim.cpp``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module. The product package never does (tests/test_product_isolation.py enforces it).
"""
import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_PATH = os.path.join(_HERE, 'liboracle_coder.so')
_REF_PATH = os.path.join(_HERE, '_ref', 'libref_coder.so')

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i16p = ctypes.POINTER(ctypes.c_int16)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compiles the C restatement (and the reference coder when /root/reference is present)."""
    if force or not os.path.isfile(_PORT_PATH) or \
            os.path.getmtime(_PORT_PATH) < os.path.getmtime(os.path.join(_HERE, 'coder_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, 'liboracle_coder.so'], stdout=subprocess.DEVNULL)
    if os.path.isdir('/root/reference') and (force or not os.path.isfile(_REF_PATH)):
        subprocess.check_call(['make', '-C', _HERE, 'ref'], stdout=subprocess.DEVNULL)


def has_ref():
    return os.path.isfile(_REF_PATH)


class _Lib(object):
    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        f = getattr(self.lib, prefix + '_compress_lossless')
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_uint32, _i16p, _i16p, ctypes.c_uint8 if prefix == 'ref' else ctypes.c_uint32,
                      _f64p, _u32p]
        f = getattr(self.lib, prefix + '_encode_map')
        f.restype = ctypes.c_int
        lt = ctypes.c_uint8 if prefix == 'ref' else ctypes.c_uint32
        f.argtypes = [ctypes.c_uint32, _i16p, lt, _f64p, _u8p, ctypes.c_uint32, _u32p,
                      _u8p, ctypes.c_uint32, _u32p]
        f = getattr(self.lib, prefix + '_decode_map')
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_uint32, _i16p, lt, _f64p, _u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32]
        f = getattr(self.lib, prefix + '_bac_encode_bits')
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_uint32, _u8p, _f64p, _u8p, ctypes.c_uint32, _u32p]
        for name in ('_create_divisible', '_count_nb_bits'):
            f = getattr(self.lib, prefix + name)
            f.restype = ctypes.c_uint32
        getattr(self.lib, prefix + '_create_divisible').argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        getattr(self.lib, prefix + '_count_nb_bits').argtypes = [ctypes.c_uint32]
        if prefix == 'oracle':
            f = self.lib.oracle_compress_maps
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_uint32, ctypes.c_uint32, _i16p, _i16p, ctypes.c_uint32, _f64p, _u32p]


_libs = {}


def _get(which):
    if which not in _libs:
        if which == 'port':
            build()
            _libs[which] = _Lib(_PORT_PATH, 'oracle')
        elif which == 'ref':
            build()
            if not has_ref():
                raise RuntimeError('oracle/_ref/libref_coder.so is absent (reference tree not available).')
            _libs[which] = _Lib(_REF_PATH, 'ref')
        else:
            raise ValueError(which)
    return _libs[which]


def _ptr(a, t):
    return a.ctypes.data_as(t)


def capacity_bytes(size, L):
    """Bytes of each of the two per-map buffers (compression.cpp:24, Bitstream.cpp:3-7)."""
    bits = size*max(32, L)
    return (bits + 7)//8


def compress_lossless(ref_map_int16, probabilities, which='port'):
    """``compress_lossless`` (compression.cpp:3-65). Returns (error, reconstruction, nb_bits)."""
    lib = _get(which)
    a = numpy.ascontiguousarray(ref_map_int16, dtype=numpy.int16)
    p = numpy.ascontiguousarray(probabilities, dtype=numpy.float64)
    out = numpy.zeros_like(a)
    nb = ctypes.c_uint32(0)
    err = getattr(lib.lib, lib.prefix + '_compress_lossless')(
        a.size, _ptr(a, _i16p), _ptr(out, _i16p), p.size, _ptr(p, _f64p), ctypes.byref(nb))
    return (err, out, nb.value)


def encode_map(ref_map_int16, probabilities, which='port'):
    """Encode half. Returns (error, bac_bytes, bac_bits, bypass_bytes, bypass_bits)."""
    lib = _get(which)
    a = numpy.ascontiguousarray(ref_map_int16, dtype=numpy.int16)
    p = numpy.ascontiguousarray(probabilities, dtype=numpy.float64)
    cap = capacity_bytes(a.size, p.size) + 8
    bac = numpy.zeros(cap, dtype=numpy.uint8)
    byp = numpy.zeros(cap, dtype=numpy.uint8)
    nb_bac = ctypes.c_uint32(0)
    nb_byp = ctypes.c_uint32(0)
    err = getattr(lib.lib, lib.prefix + '_encode_map')(
        a.size, _ptr(a, _i16p), p.size, _ptr(p, _f64p),
        _ptr(bac, _u8p), cap, ctypes.byref(nb_bac), _ptr(byp, _u8p), cap, ctypes.byref(nb_byp))
    if err:
        return (err, None, 0, None, 0)
    return (0, bac[:(nb_bac.value + 7)//8].copy(), nb_bac.value, byp[:(nb_byp.value + 7)//8].copy(), nb_byp.value)


def decode_map(size, probabilities, bac_bytes, bac_bits, byp_bytes, byp_bits, which='port'):
    """Decode half from external buffers. Returns (error, int16[size])."""
    lib = _get(which)
    p = numpy.ascontiguousarray(probabilities, dtype=numpy.float64)
    bac = numpy.zeros(max(1, (bac_bits + 7)//8), dtype=numpy.uint8)
    bac[:len(bac_bytes)] = bac_bytes
    byp = numpy.zeros(max(1, (byp_bits + 7)//8), dtype=numpy.uint8)
    byp[:len(byp_bytes)] = byp_bytes
    out = numpy.zeros(size, dtype=numpy.int16)
    err = getattr(lib.lib, lib.prefix + '_decode_map')(
        size, _ptr(out, _i16p), p.size, _ptr(p, _f64p), _ptr(bac, _u8p), bac_bits, _ptr(byp, _u8p), byp_bits)
    return (err, out)


def bac_encode_bits(bits, probabilities, which='port'):
    """Raw arithmetic coder, one probability per bit (tests.cpp:69-132). Returns (error, bytes, nb_bits)."""
    lib = _get(which)
    b = numpy.ascontiguousarray(bits, dtype=numpy.uint8)
    p = numpy.ascontiguousarray(probabilities, dtype=numpy.float64)
    cap = 16 + b.size*8
    out = numpy.zeros(cap, dtype=numpy.uint8)
    nb = ctypes.c_uint32(0)
    err = getattr(lib.lib, lib.prefix + '_bac_encode_bits')(
        b.size, _ptr(b, _u8p), _ptr(p, _f64p), _ptr(out, _u8p), cap, ctypes.byref(nb))
    return (err, out[:(nb.value + 7)//8].copy(), nb.value)


def compress_maps_planar(maps_int16, table, which='port'):
    """All maps of one latent, planar [C, n]. Returns (error, reconstruction [C, n], bits uint32[C])."""
    maps = numpy.ascontiguousarray(maps_int16, dtype=numpy.int16)
    p = numpy.ascontiguousarray(table, dtype=numpy.float64)
    (nb_maps, size) = maps.shape
    out = numpy.zeros_like(maps)
    bits = numpy.zeros(nb_maps, dtype=numpy.uint32)
    if which == 'port':
        lib = _get('port')
        err = lib.lib.oracle_compress_maps(nb_maps, size, _ptr(maps, _i16p), _ptr(out, _i16p),
                                           p.shape[1], _ptr(p, _f64p), _ptr(bits, _u32p))
        return (err, out, bits)
    for i in range(nb_maps):
        (err, out[i], bits[i]) = compress_lossless(maps[i], p[i], which=which)
        if err:
            return (err, out, bits)
    return (0, out, bits)


def create_divisible(x, d, which='port'):
    lib = _get(which)
    return getattr(lib.lib, lib.prefix + '_create_divisible')(x, d)


def count_nb_bits(x, which='port'):
    lib = _get(which)
    return getattr(lib.lib, lib.prefix + '_count_nb_bits')(x)

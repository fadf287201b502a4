"""The product's per-stream coder core (csrc/coder_core.cuh), compiled as host C++, against the oracle:
same bytes, same bit counts, same error codes, exact decode. Runs without a GPU; the CUDA kernels
instantiate exactly this code."""
import ctypes
import os
import subprocess

import numpy
import pytest

from oracle import coder as oracle_coder
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_harness', 'coder_core_harness.cpp')
CORE = os.path.join(HERE, '..', 'autoencoder_based_image_compression_b200', 'csrc', 'coder_core.cuh')
LIB = os.path.join(HERE, 'host_harness', 'libcoder_core_harness.so')


@pytest.fixture(scope='module')
def harness():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE))
    if not os.path.isfile(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off', '-x', 'c++', SRC,
                               '-o', LIB])
    lib = ctypes.CDLL(LIB)
    for name in ('harness_encode', 'harness_decode', 'harness_encode2', 'harness_decode2', 'harness_encode3',
                 'harness_decode3'):
        getattr(lib, name).restype = ctypes.c_int
    lib.harness_e3_exhaustive.restype = ctypes.c_uint64
    lib.harness_rescale_exhaustive.restype = ctypes.c_uint64
    return lib


class Formulation(object):
    """`encode` / `decode` entry points of one formulation of the core: 'v1' = flattened one-pass loop,
    'v2' = lean two-pass form (the kernels' path for table rows with invalid entries), 'v3-fp64' / 'v3-fixed' =
    the branch-free form the default CUDA kernels instantiate, with the reference's FP64 multiply or the
    validated 48-bit fixed-point one."""

    def __init__(self, lib, suffix, mode):
        enc = getattr(lib, 'harness_encode' + suffix)
        dec = getattr(lib, 'harness_decode' + suffix)
        extra = () if mode is None else (ctypes.c_int(mode),)
        self.harness_encode = lambda *a: enc(*(a + extra))
        self.harness_decode = lambda *a: dec(*(a + extra))


@pytest.fixture(scope='module', params=[('', None), ('2', None), ('3', 0), ('3', 1)],
                ids=['v1', 'v2', 'v3-fp64', 'v3-fixed'])
def core(request, harness):
    return Formulation(harness, *request.param)


def cap_bits(size, L):
    bits = size*max(32, L)
    return (bits + 7)//8*8


def encode(lib, x, p):
    x = numpy.ascontiguousarray(x, dtype=numpy.int16)
    p = numpy.ascontiguousarray(p, dtype=numpy.float64)
    cap = cap_bits(x.size, p.size)
    bac = numpy.zeros(cap//8 + 64, dtype=numpy.uint8)
    byp = numpy.zeros(cap//8 + 64, dtype=numpy.uint8)
    bb = ctypes.c_uint32(0)
    rb = ctypes.c_uint32(0)
    err = lib.harness_encode(ctypes.c_uint32(x.size), x.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(p.size),
                             p.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(cap),
                             bac.ctypes.data_as(ctypes.c_void_p), ctypes.byref(bb),
                             byp.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rb))
    return (err, bac[:(bb.value + 7)//8], bb.value, byp[:(rb.value + 7)//8], rb.value)


def decode(lib, size, p, bac, bb, byp, rb, misalign=0):
    p = numpy.ascontiguousarray(p, dtype=numpy.float64)
    out = numpy.zeros(size, dtype=numpy.int16)
    bac = numpy.ascontiguousarray(numpy.append(bac, numpy.zeros(4, dtype=numpy.uint8)))
    byp = numpy.ascontiguousarray(numpy.append(byp, numpy.zeros(4, dtype=numpy.uint8)))
    err = lib.harness_decode(ctypes.c_uint32(size), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(p.size),
                             p.ctypes.data_as(ctypes.c_void_p), bac.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_uint32(bb), byp.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(rb),
                             ctypes.c_uint32(misalign))
    return (err, out)


def test_known_answers(core, golden):
    for (name, x, p, bac, byp, bac_bits, byp_bits) in golden.kat_cases():
        (err, b, bb, r, rb) = encode(core, x, p)
        assert err == 0 and (bb, rb) == (bac_bits, byp_bits), name
        assert numpy.array_equal(b, bac) and numpy.array_equal(r, byp), name
        for mis in range(4):
            (err, dec) = decode(core, x.size, p, bac, bac_bits, byp, byp_bits, mis)
            assert err == 0 and numpy.array_equal(dec, x), (name, mis)


def test_random_streams_against_the_oracle(core):
    rng = numpy.random.default_rng(17)
    for trial in range(400):
        L = int(rng.integers(1, 41)) if trial % 9 else 255
        # skewed tables exercise long E3 queues and capacity errors
        p = rng.uniform(0.02, 0.98, size=L) if trial % 5 else rng.choice([0.01, 0.5, 0.99, 0.999, 0.2], size=L)
        size = int(rng.integers(0, 800))
        x = util.laplace_latent(rng, float(rng.choice([0.2, 1., 6., 50., 4000.])), shape=(max(size, 1), 1))[:size, 0].copy()
        if size and trial % 7 == 0:
            x[rng.integers(0, size)] = -32768
            x[rng.integers(0, size)] = 32767
        want = oracle_coder.encode_map(x, p, 'port')
        got = encode(core, x, p)
        assert got[0] == want[0], (trial, got[0], want[0])
        if want[0] == 0:
            assert (got[2], got[4]) == (want[2], want[4]), trial
            assert numpy.array_equal(got[1], want[1]) and numpy.array_equal(got[3], want[3]), trial
            (err, dec) = decode(core, size, p, want[1], want[2], want[3], want[4], trial % 4)
            assert err == 0 and numpy.array_equal(dec, x), trial


def test_more_than_31_follow_bits_from_one_bin(core):
    """Streams whose interval straddles the midpoint for dozens of E3 steps (BinaryArithmeticCoder.cpp:238-246), so that
    one renormalisation releases a bit and 45 / 52 / 60 queued follow bits: more than one 32-bit put. Found by hill
    climbing on a model of the coder's registers; every formulation must take its slow path and agree with the oracle."""
    cases = [([0.5059247899741813, 0.1629641289648215, 0.05125147724638375],
              [0, 1, 1, 3, 1, 1, 2, 1, 2, 1, 3, 3, 2, 2, 1, 3, 1, 3, 3, 1, 0, 0, 0, 1, 0, 3, 3, 0, 0, 1, 0, 3, 1, 0, 1, 3, 1, 0, 0, 2]),
             ([0.6068877841736549, 0.37167546978247545, 0.32194534925419743],
              [2, 0, 2, 0, 1, 0, 3, 1, 0, 2, 2, 3, 0, 2, 0, 1, 0, 0, 1, 0, 3, 1, 3, 3, 1, 3, 0, 0, 2, 0, 1, 0, 0, 0, 0, 2, 1, 1, 3, 1]),
             ([0.0479767467199812, 0.468634543253395, 0.25753457899273247],
              [2, 0, 3, 3, 2, 1, 3, 3, 2, 1, 1, 1, 1, 3, 1, 3, 3, 1, 1, 2, 1, 1, 2, 3, 3, 2, 1, 1, 3, 1, 1, 2, 1, 2, 0, 3, 3, 2, 3, 3])]
    for (probs, magnitudes) in cases:
        p = numpy.array(probs)
        x = numpy.array(magnitudes, dtype=numpy.int16)
        x[::3] *= -1
        want = oracle_coder.encode_map(x, p, 'port')
        got = encode(core, x, p)
        assert want[0] == 0 and got[0] == 0 and (got[2], got[4]) == (want[2], want[4])
        assert numpy.array_equal(got[1], want[1]) and numpy.array_equal(got[3], want[3])
        (err, dec) = decode(core, x.size, p, want[1], want[2], want[3], want[4], 1)
        assert err == 0 and numpy.array_equal(dec, x)


def test_malformed_streams_decode_like_the_oracle(core):
    """Truncated / corrupted inputs: same error code and, when both succeed, the same symbols (stale-bit
    padding of BinaryArithmeticCoder.cpp:104-122, 275-315)."""
    rng = numpy.random.default_rng(23)
    for trial in range(200):
        L = int(rng.integers(1, 12))
        p = rng.uniform(0.05, 0.95, size=L)
        size = int(rng.integers(1, 200))
        x = util.laplace_latent(rng, float(rng.choice([1., 8., 300.])), shape=(size, 1))[:, 0]
        (err, bac, bb, byp, rb) = oracle_coder.encode_map(x, p, 'port')
        assert err == 0
        mode = trial % 4
        if mode == 0:
            bb = int(rng.integers(0, bb + 1))                # arithmetic stream cut short
        elif mode == 1:
            rb = int(rng.integers(0, rb + 1))                # bypass stream cut short
        elif mode == 2:
            bac = bac.copy()
            bac[rng.integers(0, bac.size)] ^= 1 << int(rng.integers(0, 8))
        else:
            bac = rng.integers(0, 256, size=bac.size, dtype=numpy.uint8)
        want = oracle_coder.decode_map(size, p, bac[:(bb + 7)//8], bb, byp[:(rb + 7)//8], rb, 'port')
        got = decode(core, size, p, bac[:(bb + 7)//8], bb, byp[:(rb + 7)//8], rb, trial % 3)
        assert got[0] == want[0], (trial, mode, got[0], want[0])
        if want[0] == 0:
            assert numpy.array_equal(got[1], want[1]), (trial, mode)


def test_error_codes(core):
    x = numpy.array([0, 3, -2, 0], dtype=numpy.int16)
    for probs in ([0.5, numpy.nan, 0.5], [0.5, 1.0, 0.5], [0.5, 0.0, 0.5]):
        assert encode(core, x, probs)[0] == 4
    assert encode(core, numpy.array([0, 1, 0], dtype=numpy.int16), [0.5, 0.5, numpy.nan])[0] == 0
    assert encode(core, numpy.array([40], dtype=numpy.int16), numpy.full(40, 0.99))[0] == 1
    assert encode(core, numpy.zeros(0, dtype=numpy.int16), [0.5])[0] == 1      # zero-size buffer: capacity error


def test_closed_form_e3_matches_the_literal_loop_for_every_register_pair(harness):
    """e3_steps() and the closed-form register update against the reference's while loop
    (BinaryArithmeticCoder.cpp:238-246) on all 2^30 pairs low < 0x8000 <= high."""
    assert harness.harness_e3_exhaustive() == 0


def test_combined_rescaling_matches_the_literal_loop_for_every_register_pair(harness):
    """fast_rescale() (E1/E2 + E3 as one shift, quirk fix-up) against the reference's rescaling loop
    (BinaryArithmeticCoder.cpp:182-252) on all 2^31 pairs low <= high."""
    assert harness.harness_rescale_exhaustive() == 0

"""Pins the CPU coder oracle (oracle/coder_oracle.c) against the reference's known answers and the
bytes of the compiled reference coder (fixtures written by tests/golden/make_golden.py)."""
import hashlib

import numpy
import pytest

from oracle import coder
from tests import util


def test_known_answers_of_the_reference_tests(golden):
    expected_bits = {'tests_cpp_354_compress_lossless': 104,   # tests.cpp:354-376
                     'tests_cpp_280_signed_ueg0': 64 + 49,       # tests.cpp:280-352
                     'test_lossless_py_96': 20}                  # test_lossless.py:89-101
    for (name, x, p, bac, byp, bac_bits, byp_bits) in golden.kat_cases():
        (err, b, bb, r, rb) = coder.encode_map(x, p, 'port')
        assert err == 0
        assert (bb, rb) == (bac_bits, byp_bits), name
        assert numpy.array_equal(b, bac) and numpy.array_equal(r, byp), name
        if name in expected_bits:
            assert bb + rb == expected_bits[name]
        (err, rec, nb) = coder.compress_lossless(x, p, 'port')
        assert err == 0 and nb == bb + rb and numpy.array_equal(rec, x)
        (err, dec) = coder.decode_map(x.size, p, bac, bac_bits, byp, byp_bits, 'port')
        assert err == 0 and numpy.array_equal(dec, x)


def test_survey_byte_vectors(golden):
    # bytes dumped from the reference with -fno-access-control
    kat = {c[0]: c for c in golden.kat_cases()}
    assert kat['tests_cpp_354_compress_lossless'][3].tobytes().hex() == 'e6ffffff1ffe05'
    assert kat['tests_cpp_354_compress_lossless'][4].tobytes().hex() == 'fef3f6c67f0d02'
    assert kat['test_lossless_py_96'][3].tobytes().hex() == 'da82'
    assert kat['test_lossless_py_96'][4].tobytes().hex() == '0d'
    assert kat['skewed_row0'][3].tobytes().hex() == '9959b4ebff19efe1ff01'


def test_raw_arithmetic_coder(golden):
    kat = golden.load('coder_kat')
    (err, data, nb) = coder.bac_encode_bits(kat['raw_bac__bits_in'], kat['raw_bac__probs'], 'port')
    assert err == 0 and nb == 31 == int(kat['raw_bac__nb_bits'])   # tests.cpp:69-132
    assert numpy.array_equal(data, kat['raw_bac__bytes'])


def test_utils(golden):
    kat = golden.load('coder_kat')
    for (x, d, want) in kat['create_divisible']:                     # tests.cpp:5-16: 36, 45, 105
        assert coder.create_divisible(int(x), int(d), 'port') == want
    assert [int(r[2]) for r in kat['create_divisible']] == [36, 45, 105]
    for (x, want) in kat['count_nb_bits']:
        assert coder.count_nb_bits(int(x), 'port') == want


def test_random_latents_against_reference_bytes(golden):
    rnd = golden.load('coder_random')
    for k in range(4):
        x = rnd['case{}__latent'.format(k)]
        mult = str(int(rnd['case{}__seed_scale_mult'.format(k)][2]))
        table = golden.table('1_10000', mult)
        sha = hashlib.sha256()
        for i in range(128):
            (err, bac, bb, byp, rb) = coder.encode_map(x[:, :, i].flatten(), table[i], 'port')
            assert err == 0
            assert bb == rnd['case{}__bac_bits'.format(k)][i] and rb == rnd['case{}__byp_bits'.format(k)][i]
            sha.update(bac.tobytes())
            sha.update(byp.tobytes())
        assert numpy.array_equal(numpy.frombuffer(sha.digest(), dtype=numpy.uint8), rnd['case{}__sha256'.format(k)])


def test_error_codes():
    x = numpy.array([0, 3, -2, 0], dtype=numpy.int16)
    assert coder.compress_lossless(x, numpy.array([0.5, numpy.nan, 0.5]), 'port')[0] == 4
    assert coder.compress_lossless(x, numpy.array([0.5, 1.0, 0.5]), 'port')[0] == 4
    assert coder.compress_lossless(x, numpy.array([0.5, 0.0, 0.5]), 'port')[0] == 4
    # a NaN that is never reached is fine (test_lossless.py:345-352)
    assert coder.compress_lossless(numpy.array([0, 1, 0], dtype=numpy.int16), numpy.array([0.5, 0.5, numpy.nan]), 'port')[0] == 0
    # capacity: 1 symbol, L = 40 adversarial probabilities -> more than 40 bits of arithmetic code
    assert coder.compress_lossless(numpy.array([40], dtype=numpy.int16), numpy.full(40, 0.99), 'port')[0] == 1
    # L == 0: std::out_of_range in the reference
    assert coder.compress_lossless(x, numpy.zeros(0), 'port')[0] == -2


@pytest.mark.skipif(not coder.has_ref(), reason='oracle/_ref not built (reference tree absent)')
def test_port_equals_compiled_reference_on_random_maps(golden):
    rng = numpy.random.default_rng(5)
    for trial in range(40):
        L = int(rng.integers(1, 41))
        p = rng.uniform(0.02, 0.98, size=L)
        size = int(rng.integers(1, 600))
        x = util.laplace_latent(rng, float(rng.choice([0.3, 2., 30., 3000.])), shape=(size, 1))[:, 0]
        a = coder.encode_map(x, p, 'ref')
        b = coder.encode_map(x, p, 'port')
        assert a[0] == b[0]
        if a[0] == 0:
            assert a[2] == b[2] and a[4] == b[4]
            assert numpy.array_equal(a[1], b[1]) and numpy.array_equal(a[3], b[3])
            for which in ('ref', 'port'):
                (err, dec) = coder.decode_map(size, p, a[1], a[2], a[3], a[4], which)
                assert err == 0 and numpy.array_equal(dec, x)
    for probs in ([0.5, numpy.nan], [0.99]*40, [1.5, 0.5]):
        x = numpy.array([40, -3, 1], dtype=numpy.int16)
        assert coder.compress_lossless(x, numpy.array(probs), 'ref')[0] == coder.compress_lossless(x, numpy.array(probs), 'port')[0]

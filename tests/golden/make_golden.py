"""Generates the committed golden fixtures from the REFERENCE ITSELF, in the build container.

    python tests/golden/make_golden.py

Needs /root/reference (read-only). It (1) compiles the reference coder into oracle/_ref (see
oracle/Makefile), (2) builds the reference's own Cython module in a scratch copy under /tmp and
imports the reference's lossless/compression.py, lossless/stats.py and tools/tools.py (matplotlib
stubbed, numpy.float aliased) and (3) writes small .npz fixtures next to this file. Nothing under
tests/ reads /root/reference at test time; the GPU box only sees these fixtures.

Fixtures:
  tables.npz        shipped probability tables / map means / exception indices used by tests and bench
                    (reference DATA: lossless/results/**, lossless/pseudo_data/*.npy)
  coder_kat.npz     known-answer vectors of the reference's own tests with the byte buffers of the
                    compiled reference coder
  coder_random.npz  seeded random maps -> per-map bit counts and SHA-256 of both buffers (reference)
  compression.npz   lossless.compression.{compress_lossless_maps, rescale_compress_lossless_maps}
                    outputs of the reference Python + Cython path on seeded latents
  glue.npz          tools.tools / lossless.stats outputs of the reference on seeded inputs
"""
import hashlib
import os
import pickle
import shutil
import subprocess
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/kodak_tensorflow'
sys.path.insert(0, ROOT)

from oracle import coder  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    numpy.savez_compressed(path, **arrays)
    print('wrote {} ({} bytes)'.format(path, os.path.getsize(path)))


def import_reference_python():
    scratch = '/tmp/eae_ref_build'
    if os.path.isdir(scratch):
        shutil.rmtree(scratch)
    os.makedirs(scratch)
    for sub in ('lossless', 'tools', 'eae'):
        shutil.copytree(os.path.join(REF, sub), os.path.join(scratch, sub),
                        ignore=shutil.ignore_patterns('results', 'visualization', 'pseudo_visualization'))
    env = dict(os.environ, CXXFLAGS='-include cstdint -include math.h')
    subprocess.check_call([sys.executable, 'setup.py', 'build_ext', '--inplace'],
                          cwd=os.path.join(scratch, 'lossless'), env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].use = lambda *a, **k: None
    numpy.float = numpy.floating
    sys.path.insert(0, scratch)
    import lossless.compression
    import lossless.stats
    import tools.tools
    return (lossless.compression, lossless.stats, tools.tools)


def make_tables():
    out = {}
    for model in ('1_10000', 'learning_bw_0dot5_10000'):
        base = os.path.join(REF, 'lossless/results', model, 'training_index_10')
        for mult in ('1', '2', '4', '8'):
            out['{}/binary_probabilities_{}'.format(model, mult)] = numpy.load(
                os.path.join(base, 'binary_probabilities_{}.npy'.format(mult)))
        out['{}/map_mean'.format(model)] = numpy.load(os.path.join(base, 'map_mean.npy'))
        with open(os.path.join(base, 'idx_map_exception.pkl'), 'rb') as f:
            out['{}/idx_map_exception'.format(model)] = numpy.array(pickle.load(f), dtype=numpy.int64)
    pseudo = os.path.join(REF, 'lossless/pseudo_data')
    for name in sorted(os.listdir(pseudo)):
        if name.endswith('.npy'):
            out['pseudo_data/' + name[:-4]] = numpy.load(os.path.join(pseudo, name))
    save('tables.npz', **{k.replace('/', '__'): v for (k, v) in out.items()})
    return out


def make_coder_kat():
    cases = [
        # (name, symbols, probabilities)  -- sources in the reference's own tests
        ('tests_cpp_354_compress_lossless', [0, -2, 0, 765, -21, 8, -439, 0, 0, 0, 0, -9], [0.5]*8),
        ('tests_cpp_280_signed_ueg0', [0, 1, -2, -7, 8, -8, 9, -9, 127, -523], [0.5]*8),
        ('test_lossless_py_96', [0, 1, -2, 2, 1, 0, 0, 0], [0.5]*3),
        ('skewed_row0', [0, 0, 1, 0, -1, 3, 0, 0, -12, 0, 2, 0, 0, -1, 25, 0],
         [.805, .802, .75, .753, .729, .741, .759, .786, .667, .99]),
        ('extremes', [32767, -32768, -32767, 1, 0, 255, -256, 16384], [0.3, 0.6, 0.9, 0.1, 0.5]),
        ('single_zero', [0], [0.5]),
        ('L1', [0, 1, -1, 2, 5, 0], [0.7]),
    ]
    out = {}
    for (name, symbols, probs) in cases:
        x = numpy.array(symbols, dtype=numpy.int16)
        p = numpy.array(probs, dtype=numpy.float64)
        (err, bac, bac_bits, byp, byp_bits) = coder.encode_map(x, p, 'ref')
        assert err == 0
        (err2, rec, nb) = coder.compress_lossless(x, p, 'ref')
        assert err2 == 0 and numpy.array_equal(rec, x) and nb == bac_bits + byp_bits
        out[name + '__symbols'] = x
        out[name + '__probs'] = p
        out[name + '__bac'] = bac
        out[name + '__byp'] = byp
        out[name + '__bits'] = numpy.array([bac_bits, byp_bits], dtype=numpy.uint32)
    # raw BAC known answer (tests.cpp:69-132): 31 bits
    probs = [0.01, 0.99, 0.9, 0.76, 0.1, 0.01, 0.99, 0.5, 0.51, 0.2, 0.52, 0.01, 0.1, 0.01, 0.2, 0.90, 0.05, 0.5,
             0.53, 0.2]
    bits = [1 if 8 <= i <= 14 else 0 for i in range(20)]
    (err, data, nb) = coder.bac_encode_bits(bits, probs, 'ref')
    assert err == 0 and nb == 31
    out['raw_bac__bits_in'] = numpy.array(bits, dtype=numpy.uint8)
    out['raw_bac__probs'] = numpy.array(probs)
    out['raw_bac__bytes'] = data
    out['raw_bac__nb_bits'] = numpy.array(nb, dtype=numpy.uint32)
    # utils.cpp known answers (tests.cpp:5-16)
    out['create_divisible'] = numpy.array([[31, 9, coder.create_divisible(31, 9, 'ref')],
                                           [45, 5, coder.create_divisible(45, 5, 'ref')],
                                           [101, 5, coder.create_divisible(101, 5, 'ref')]], dtype=numpy.uint32)
    out['count_nb_bits'] = numpy.array([[v, coder.count_nb_bits(v, 'ref')]
                                        for v in (0, 1, 2, 3, 4, 255, 256, 32768, 32769, 65535, 65536)],
                                       dtype=numpy.uint32)
    save('coder_kat.npz', **out)


def latent(rng, scale, shape=(32, 48, 128)):
    """Laplace-distributed int16 latent with per-map scales spread around `scale`."""
    scales = scale*numpy.exp(rng.normal(0., 0.7, size=shape[2]))
    x = rng.laplace(0., 1., size=shape)*scales.reshape((1, 1, -1))
    return numpy.round(x).clip(-32767, 32767).astype(numpy.int16)


def make_coder_random(tables):
    out = {}
    rng = numpy.random.default_rng(1234)
    k = 0
    for (mult, scale) in (('1', 2.0), ('2', 1.0), ('4', 0.5), ('8', 0.25)):
        table = tables['1_10000/binary_probabilities_' + mult]
        x = latent(rng, scale)
        bac_bits = numpy.zeros(128, dtype=numpy.uint32)
        byp_bits = numpy.zeros(128, dtype=numpy.uint32)
        sha = hashlib.sha256()
        for i in range(128):
            (err, bac, bb, byp, rb) = coder.encode_map(x[:, :, i].flatten(), table[i], 'ref')
            assert err == 0
            (bac_bits[i], byp_bits[i]) = (bb, rb)
            sha.update(bac.tobytes())
            sha.update(byp.tobytes())
        out['case{}__seed_scale_mult'.format(k)] = numpy.array([1234, scale, float(mult)])
        out['case{}__latent'.format(k)] = x
        out['case{}__bac_bits'.format(k)] = bac_bits
        out['case{}__byp_bits'.format(k)] = byp_bits
        out['case{}__sha256'.format(k)] = numpy.frombuffer(sha.digest(), dtype=numpy.uint8)
        k += 1
    save('coder_random.npz', **out)


def make_compression(ref_compression, tables):
    out = {}
    rng = numpy.random.default_rng(99)
    table_path = os.path.join(REF, 'lossless/results/1_10000/training_index_10/binary_probabilities_1.npy')
    idx_exc = int(tables['1_10000/idx_map_exception'])
    x = latent(rng, 1.5)
    x[:, :, idx_exc] = rng.integers(-40, 41, size=(32, 48)).astype(numpy.int16)   # near-uniform exception map
    (rec, bits) = ref_compression.compress_lossless_maps(x, table_path, idx_map_exception=idx_exc)
    assert numpy.array_equal(rec, x)
    (rec2, bits2) = ref_compression.compress_lossless_maps(x, table_path)
    out['maps__latent'] = x
    out['maps__idx_exc'] = numpy.array(idx_exc)
    out['maps__bits_exc'] = bits
    out['maps__bits_noexc'] = bits2
    bw = (0.8 + 3.2*rng.random(128)).astype(numpy.float32)
    cq = (x.astype(numpy.float32)*bw.reshape((1, 1, -1))).astype(numpy.float32)
    total = ref_compression.rescale_compress_lossless_maps(cq, bw, table_path, idx_map_exception=idx_exc)
    out['rescale__bin_widths'] = bw
    out['rescale__total_bits'] = numpy.array(total, dtype=numpy.int64)
    # error behaviour (test_lossless.py:329-375): both invalid tables must raise "Error of type 4"
    # same construction as test_lossless.py:343-368: N(0, 5), N(0, 0.2), N(0, 0.5) maps, bin widths 1.5;
    # the NaN entries of the "valid" table are never reached by maps 1 and 2.
    bw3 = numpy.array([1.5, 1.5, 1.5], dtype=numpy.float32)
    centered = numpy.stack([rng.normal(0., s, size=(96, 48)) for s in (5., 0.2, 0.5)], axis=2).astype(numpy.float32)
    small_q = (bw3.reshape((1, 1, 3))*numpy.round(centered/bw3.reshape((1, 1, 3)))).astype(numpy.float32)
    pseudo = os.path.join(REF, 'lossless/pseudo_data')
    out['invalid__centered_quantized'] = small_q
    out['invalid__bin_widths'] = bw3
    ok_bits = ref_compression.rescale_compress_lossless_maps(
        small_q, bw3, os.path.join(pseudo, 'binary_probabilities_scale_compress_valid.npy'))
    out['invalid__valid_total_bits'] = numpy.array(ok_bits, dtype=numpy.int64)
    messages = []
    for k in (0, 1):
        try:
            ref_compression.rescale_compress_lossless_maps(
                small_q, bw3,
                os.path.join(pseudo, 'binary_probabilities_scale_compress_invalid_{}.npy'.format(k)))
            messages.append('no error')
        except RuntimeError as err:
            messages.append(str(err))
    out['invalid__messages'] = numpy.array(messages)
    save('compression.npz', **out)


def make_glue(ref_stats, tls):
    out = {}
    rng = numpy.random.default_rng(7)
    data = (rng.laplace(0., 3., size=(3, 12, 20, 128))).astype(numpy.float32)
    bw = (0.5 + 3.5*rng.random(128)).astype(numpy.float32)
    q = tls.quantize_per_map(data, bw)
    out['quantize__data'] = data
    out['quantize__bin_widths'] = bw
    out['quantize__out'] = q
    halves = numpy.array([[[[0.5, 1.5, 2.5, -0.5, -1.5, 3.5000002, 2.4999998, -2.5]]]], dtype=numpy.float32)
    out['quantize__halves'] = halves
    out['quantize__halves_out'] = tls.quantize_per_map(halves, numpy.ones(8, dtype=numpy.float32))
    x = numpy.array([15.431, -0.001, 0., 235.678, 143.18, 1.111, 16.5, 17.5, 234.5, 100.49999], dtype=numpy.float32)
    out['cast_bt601__in'] = x
    out['cast_bt601__out'] = tls.cast_bt601(x)
    y = numpy.array([0.49, -0.51, 2.5, 3.5, -2.5, 32767.4, -32767.49, 1e-3], dtype=numpy.float32)
    out['cast_int16__in'] = y
    out['cast_int16__out'] = tls.cast_float_to_int16(y)
    a = rng.integers(16, 236, size=(64, 96), dtype=numpy.uint8)
    b = numpy.clip(a.astype(numpy.int64) + rng.integers(-9, 10, size=a.shape), 0, 255).astype(numpy.uint8)
    out['psnr__a'] = a
    out['psnr__b'] = b
    out['psnr__out'] = numpy.array(tls.psnr_2d(a, b))
    out['psnr__known'] = numpy.array(tls.psnr_2d(12*numpy.ones((2, 2), dtype=numpy.uint8),
                                                 15*numpy.ones((2, 2), dtype=numpy.uint8)))
    out['nb_deads__out'] = tls.count_nb_deads(q*(rng.random((1, 1, 1, 128)) > 0.2))
    out['nb_deads__in'] = (q*(numpy.random.default_rng(7).random((1, 1, 1, 128)) > -1)).astype(numpy.float32)
    mask = (rng.random((1, 1, 1, 128)) > 0.2).astype(numpy.float32)
    out['nb_deads__in'] = (q*mask).astype(numpy.float32)
    out['nb_deads__out'] = tls.count_nb_deads(out['nb_deads__in'])
    out['entropy__out'] = numpy.array([tls.discrete_entropy(q[0, :, :, i], bw[i].item()) for i in range(128)])
    out['count_symbols__out0'] = tls.count_symbols(q[0, :, :, 0], bw[0].item())
    out['rate_3d__out'] = numpy.array(tls.rate_3d(q[0], bw, 192, 320))
    # stats.py known answers (test_lossless.py:267-298)
    (z0, o0) = ref_stats.count_binary_decisions(numpy.array([0.75, 0.05, 0.1, 0.2, 0.2, 0.15], dtype=numpy.float32), 0.05, 7)
    (z1, o1) = ref_stats.count_binary_decisions(numpy.array([210., 6., 9., 6.], dtype=numpy.float32), 3., 7)
    out['binary_decisions__0'] = numpy.stack([z0, o0])
    out['binary_decisions__1'] = numpy.stack([z1, o1])
    mean = numpy.mean(data, axis=(0, 1, 2)).astype(numpy.float32)
    out['binary_probabilities__out'] = ref_stats.compute_binary_probabilities(data, bw, mean, 10)
    out['binary_probabilities__mean'] = mean
    out['idx_map_exception__out'] = numpy.array(ref_stats.find_index_map_exception(data))
    save('glue.npz', **out)


def make_bjontegaard(tls):
    """tools.py:157-263 on the reference's own test curves (test_tools.py:120-131) and on two 9-point curves."""
    out = {}
    rates_0 = numpy.linspace(0.15, 2.05, num=191)
    rates_1 = numpy.linspace(0.1, 1.7, num=321)
    out['case0__in'] = numpy.array([0.])
    out['case0__out'] = numpy.array(tls.compute_bjontegaard(rates_0, 40.*numpy.sqrt(rates_0), rates_1, 20.*numpy.sqrt(rates_1) + 10.))
    rng = numpy.random.default_rng(11)
    r0 = numpy.sort(0.1 + 1.9*rng.random(9))
    r1 = numpy.sort(0.1 + 1.9*rng.random(9))
    p0 = 28. + 6.*numpy.log2(1. + 4.*r0) + 0.05*rng.standard_normal(9)
    p1 = 27. + 6.2*numpy.log2(1. + 4.*r1) + 0.05*rng.standard_normal(9)
    out['case1__in'] = numpy.stack([r0, p0, r1, p1])
    out['case1__out'] = numpy.array(tls.compute_bjontegaard(r0, p0, r1, p1))
    save('bjontegaard.npz', **out)


if __name__ == '__main__':
    coder.build(force=True)
    assert coder.has_ref(), 'reference tree not available'
    tables = make_tables()
    make_coder_kat()
    make_coder_random(tables)
    (ref_compression, ref_stats, tls) = import_reference_python()
    make_compression(ref_compression, tables)
    make_glue(ref_stats, tls)
    make_bjontegaard(tls)

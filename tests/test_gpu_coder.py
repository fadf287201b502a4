"""Parity of the GPU lossless coder with the CPU oracle and the reference's golden bytes, through the
C ABI. Bit-exact: identical byte buffers, identical bit counts, exact round trip."""
import ctypes
import hashlib

import numpy
import pytest

from oracle import coder as oracle_coder
from tests import util

# The checker on the GPU box is the reference's own C++ coder (oracle/_ref travels with the snapshot); the C port only
# where the reference tree was never compiled.
WHICH = 'ref' if oracle_coder.has_ref() else 'port'

pytestmark = pytest.mark.gpu


def gpu_encode(native, x, p):
    lib = native.lib()
    cap = lib.eae_coder_capacity_bytes(x.size, p.size) + 16
    bac = numpy.zeros(cap, dtype=numpy.uint8)
    byp = numpy.zeros(cap, dtype=numpy.uint8)
    bb = ctypes.c_uint32(0)
    rb = ctypes.c_uint32(0)
    code = lib.eae_encode_map_host(x.size, native.ptr(x), p.size, native.ptr(p), native.ptr(bac), ctypes.byref(bb),
                                   native.ptr(byp), ctypes.byref(rb))
    return (code, bac[:(bb.value + 7)//8], bb.value, byp[:(rb.value + 7)//8], rb.value)


def gpu_decode(native, size, p, bac, bac_bits, byp, byp_bits):
    out = numpy.zeros(size, dtype=numpy.int16)
    bac = numpy.ascontiguousarray(numpy.append(bac, numpy.zeros(8, dtype=numpy.uint8)))
    byp = numpy.ascontiguousarray(numpy.append(byp, numpy.zeros(8, dtype=numpy.uint8)))
    code = native.lib().eae_decode_map_host(size, native.ptr(out), p.size, native.ptr(p), native.ptr(bac), bac_bits,
                                            native.ptr(byp), byp_bits)
    return (code, out)


def test_known_answers(native, golden):
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import interface_cython
    for (name, x, p, bac, byp, bac_bits, byp_bits) in golden.kat_cases():
        (code, b, bb, r, rb) = gpu_encode(native, x, p)
        assert code == 0, name
        assert (bb, rb) == (bac_bits, byp_bits), name
        assert numpy.array_equal(b, bac) and numpy.array_equal(r, byp), name
        (code, dec) = gpu_decode(native, x.size, p, bac, bac_bits, byp, byp_bits)
        assert code == 0 and numpy.array_equal(dec, x), name
        (rec, nb_bits) = interface_cython.compress_lossless_flattened_map(x, p)
        assert nb_bits == bac_bits + byp_bits and numpy.array_equal(rec, x)
    # test_lossless.py:89-101 / tests.cpp:354-376
    (rec, nb_bits) = interface_cython.compress_lossless_flattened_map(
        numpy.array([0, 1, -2, 2, 1, 0, 0, 0], dtype=numpy.int16), numpy.array([0.5, 0.5, 0.5]))
    assert nb_bits == 20
    (rec, nb_bits) = interface_cython.compress_lossless_flattened_map(
        numpy.array([0, -2, 0, 765, -21, 8, -439, 0, 0, 0, 0, -9], dtype=numpy.int16), numpy.full(8, 0.5))
    assert nb_bits == 104


def test_golden_latents_byte_identical_to_the_reference(native, golden):
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import compression
    rnd = golden.load('coder_random')
    lib = native.lib()
    for k in range(4):
        x = rnd['case{}__latent'.format(k)]
        table = golden.table('1_10000', str(int(rnd['case{}__seed_scale_mult'.format(k)][2])))
        sha = hashlib.sha256()
        for i in range(128):
            (code, bac, bb, byp, rb) = gpu_encode(native, numpy.ascontiguousarray(x[:, :, i].flatten()), table[i])
            assert code == 0
            assert bb == rnd['case{}__bac_bits'.format(k)][i] and rb == rnd['case{}__byp_bits'.format(k)][i]
            sha.update(bac.tobytes())
            sha.update(byp.tobytes())
        assert numpy.array_equal(numpy.frombuffer(sha.digest(), dtype=numpy.uint8), rnd['case{}__sha256'.format(k)])
        # batched entry point: all 128 maps in one launch, same bit counts, exact round trip
        rec = numpy.zeros_like(x)
        bits = numpy.zeros(128, dtype=numpy.uint32)
        code = lib.eae_compress_lossless_maps_host(native.ptr(x), 32, 48, 128, native.ptr(table), table.shape[1], None,
                                                   native.ptr(rec), native.ptr(bits), None)
        assert code == 0
        assert numpy.array_equal(rec, x)
        assert numpy.array_equal(bits, rnd['case{}__bac_bits'.format(k)] + rnd['case{}__byp_bits'.format(k)])


def test_random_maps_against_the_oracle(native):
    rng = numpy.random.default_rng(11)
    for trial in range(60):
        L = int(rng.integers(1, 41)) if trial % 10 else 255
        p = rng.uniform(0.02, 0.98, size=L)
        size = int(rng.integers(1, 3000))
        x = util.laplace_latent(rng, float(rng.choice([0.2, 1., 6., 50., 4000.])), shape=(size, 1))[:, 0].copy()
        if trial % 7 == 0:
            x[rng.integers(0, size)] = -32768
            x[rng.integers(0, size)] = 32767
        want = oracle_coder.encode_map(x, p, WHICH)
        got = gpu_encode(native, x, p)
        assert got[0] == want[0]
        if want[0] == 0:
            assert got[2] == want[2] and got[4] == want[4]
            assert numpy.array_equal(got[1], want[1]) and numpy.array_equal(got[3], want[3])
            (code, dec) = gpu_decode(native, size, p, want[1], want[2], want[3], want[4])
            assert code == 0 and numpy.array_equal(dec, x)


def test_error_codes_match_the_reference(native):
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import interface_cython
    lib = native.lib()
    x = numpy.array([0, 3, -2, 0], dtype=numpy.int16)
    out = numpy.zeros_like(x)
    nb = ctypes.c_uint32(0)
    for probs in ([0.5, numpy.nan, 0.5], [0.5, 1.0, 0.5], [0.5, 0.0, 0.5], [0.5, -0.1, 0.5]):
        p = numpy.array(probs)
        assert lib.eae_compress_lossless(4, native.ptr(x), native.ptr(out), 3, native.ptr(p), ctypes.byref(nb)) == 4
        assert oracle_coder.compress_lossless(x, p, WHICH)[0] == 4
        with pytest.raises(RuntimeError, match='Error of type 4'):
            interface_cython.compress_lossless_flattened_map(x, p)
    # unreached NaN is not an error
    (rec, _) = interface_cython.compress_lossless_flattened_map(numpy.array([0, 1, 0], dtype=numpy.int16),
                                                                 numpy.array([0.5, 0.5, numpy.nan]))
    assert rec.tolist() == [0, 1, 0]
    # capacity error: one symbol, 40 adversarial bins (compression.cpp:24 sizes the buffer at 40 bits)
    p = numpy.full(40, 0.99)
    one = numpy.array([40], dtype=numpy.int16)
    assert lib.eae_compress_lossless(1, native.ptr(one), native.ptr(out), 40, native.ptr(p), ctypes.byref(nb)) == 1
    assert oracle_coder.compress_lossless(one, p, WHICH)[0] == 1
    # resource error of the standalone decoder: bypass stream cut short
    x = numpy.array([100, -200, 300], dtype=numpy.int16)
    p = numpy.full(4, 0.5)
    (err, bac, bb, byp, rb) = oracle_coder.encode_map(x, p, WHICH)
    assert gpu_decode(native, 3, p, bac, bb, byp, rb - 5)[0] == 2
    assert oracle_coder.decode_map(3, p, bac, bb, byp, rb - 5, WHICH)[0] == 2


def test_python_compression_module_against_reference_outputs(native, golden, tmp_path):
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import compression
    c = golden.load('compression')
    path = str(tmp_path/'binary_probabilities_1.npy')
    numpy.save(path, golden.table('1_10000', '1'))
    x = c['maps__latent']
    idx_exc = int(c['maps__idx_exc'])
    (rec, bits) = compression.compress_lossless_maps(x, path, idx_map_exception=idx_exc)
    assert rec.dtype == numpy.int16 and bits.dtype == numpy.uint32
    assert numpy.array_equal(rec, x) and numpy.array_equal(bits, c['maps__bits_exc'])
    (rec, bits) = compression.compress_lossless_maps(x, path)
    assert numpy.array_equal(rec, x) and numpy.array_equal(bits, c['maps__bits_noexc'])
    bw = c['rescale__bin_widths']
    cq = (x.astype(numpy.float32)*bw.reshape((1, 1, -1))).astype(numpy.float32)
    assert compression.rescale_compress_lossless_maps(cq, bw, path, idx_map_exception=idx_exc) == int(c['rescale__total_bits'])
    # altered data must trip the reference's round-trip assertion (compression.py:146-153)
    bad = cq.copy()
    bad[3, 4, 5] += 0.25*bw[5]
    with pytest.raises(AssertionError):
        compression.rescale_compress_lossless_maps(bad, bw, path, idx_map_exception=idx_exc)
    # test_lossless.py:329-375
    tables = golden.load('tables')
    for (name, want) in (('valid', None), ('invalid_0', 4), ('invalid_1', 4)):
        p = str(tmp_path/(name + '.npy'))
        numpy.save(p, tables['pseudo_data__binary_probabilities_scale_compress_' + name])
        if want is None:
            total = compression.rescale_compress_lossless_maps(c['invalid__centered_quantized'], c['invalid__bin_widths'], p)
            assert total == int(c['invalid__valid_total_bits'])
        else:
            with pytest.raises(RuntimeError, match='Error of type 4 during the encoding.'):
                compression.rescale_compress_lossless_maps(c['invalid__centered_quantized'], c['invalid__bin_widths'], p)


def test_large_batch_of_streams_round_trip(native):
    """BASELINE config 2 shape: 24 latents x 128 maps coded as 3072 streams in one launch (device API)."""
    import torch
    lib = native.lib()
    rng = numpy.random.default_rng(21)
    table = numpy.ascontiguousarray(rng.uniform(0.05, 0.95, size=(128, 10)))
    lat = util.laplace_latent(rng, 2.0, shape=(24, 1536, 128))          # [n, hw, C]
    planar = numpy.ascontiguousarray(lat.transpose(0, 2, 1)).reshape(24*128, 1536)
    (n_streams, size, L) = (24*128, 1536, 10)
    slot = lib.eae_coder_slot_bytes(size, L)
    dev = torch.device('cuda', 0)
    d_nhwc = torch.from_numpy(lat).to(dev)
    d_planar = torch.empty((n_streams, size), dtype=torch.int16, device=dev)
    assert lib.eae_nhwc_to_planar_i16_dev(d_nhwc.data_ptr(), d_planar.data_ptr(), 24, size, 128, None) == 0
    assert numpy.array_equal(d_planar.cpu().numpy(), planar)
    d_table = torch.from_numpy(table).to(dev)
    d_bac = torch.zeros(n_streams*slot, dtype=torch.uint8, device=dev)
    d_byp = torch.zeros(n_streams*slot, dtype=torch.uint8, device=dev)
    d_bb = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    d_rb = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    d_err = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.eae_encode_streams_dev(d_planar.data_ptr(), n_streams, size, d_table.data_ptr(), 128, L, None,
                                      d_bac.data_ptr(), d_byp.data_ptr(), slot, d_bb.data_ptr(), d_rb.data_ptr(),
                                      d_err.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert int(d_err.abs().sum()) == 0
    (bb, rb) = (d_bb.cpu().numpy(), d_rb.cpu().numpy())
    bac = d_bac.cpu().numpy().reshape(n_streams, slot)
    byp = d_byp.cpu().numpy().reshape(n_streams, slot)
    for s in range(n_streams):          # every one of the 24 x 128 streams
        want = oracle_coder.encode_map(planar[s], table[s % 128], WHICH)
        assert want[0] == 0 and (bb[s], rb[s]) == (want[2], want[4])
        assert numpy.array_equal(bac[s, :(bb[s] + 7)//8], want[1]) and numpy.array_equal(byp[s, :(rb[s] + 7)//8], want[3])
    off = (torch.arange(n_streams, dtype=torch.int64, device=dev)*slot)
    d_out = torch.zeros((n_streams, size), dtype=torch.int16, device=dev)
    assert lib.eae_decode_streams_dev(d_out.data_ptr(), n_streams, size, d_table.data_ptr(), 128, L, None,
                                      d_bac.data_ptr(), off.data_ptr(), d_bb.data_ptr(), d_byp.data_ptr(),
                                      off.data_ptr(), d_rb.data_ptr(), d_err.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert int(d_err.abs().sum()) == 0
    assert torch.equal(d_out, d_planar)
    d_back = torch.empty_like(d_nhwc)
    assert lib.eae_planar_to_nhwc_i16_dev(d_out.data_ptr(), d_back.data_ptr(), 24, size, 128, None) == 0
    assert torch.equal(d_back, d_nhwc)

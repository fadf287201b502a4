"""The fused pipeline (encode -> quantize -> lossless code -> container -> decode) against the oracle:
bitstreams byte-identical given identical indices, indices >= 99.99 % identical, PSNR within 0.01 dB,
rate within 0.1 % (BASELINE.json north_star; configs 1-3 at test scale)."""
import numpy
import pytest
import torch

from autoencoder_based_image_compression_b200 import codec as native_codec
from autoencoder_based_image_compression_b200 import weights as wts
from oracle import coder as oracle_coder
from oracle import glue as oracle_glue
from oracle import transforms as T
from tests import util
from tests.test_gpu_coder import WHICH
from tests.test_gpu_transforms import PARITY_MODES, visible_weights

pytestmark = pytest.mark.gpu

# The codec-level bars (byte-identical bitstreams, >= 99.99 % identical indices, PSNR within 0.01 dB, rate within 0.1 %)
# also hold for the bench's default arithmetic: 3xTF32 analysis, single-pass rounded-TF32 synthesis.
CODEC_MODES = PARITY_MODES + ['mixed']


def oracle_pipeline(lum, w, learned, params):
    """CPU restatement of reconstructing_eae_kodak.py:144-224 for one batch."""
    y = T.encoder(lum[..., None].astype(numpy.float32), w, learned)
    mean = params.map_mean if params.map_mean is not None else numpy.zeros(128, dtype=numpy.float32)
    centered = y - mean.reshape((1, 1, 1, -1))
    cq = oracle_glue.quantize_per_map(centered, params.bin_widths)
    idx = oracle_glue.cast_float_to_int16(cq/params.bin_widths.reshape((1, 1, 1, -1)))
    rec = oracle_glue.cast_bt601(T.decoder(cq + mean.reshape((1, 1, 1, -1)), w, learned))[..., 0]
    return (y, idx, rec)


@pytest.mark.parametrize('math', CODEC_MODES)
@pytest.mark.parametrize('learned', [False, True])
def test_round_trip_and_byte_identity(native, golden, learned, math):
    rng = numpy.random.default_rng(3)
    w = visible_weights(0, learned)
    (n, h, wd) = (6, 256, 384)
    lum = util.synthetic_luma(rng, n, h, wd)
    model = 'learning_bw_0dot5_10000' if learned else '1_10000'
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), golden.table(model, '1'),
                                       0.05*golden.map_mean(model))
    codec = native_codec.Codec(w, learned, math=math)
    (blob, stats) = codec.compress(lum, params, return_stats=True)
    idx_gpu = codec.last_indices(n, h, wd)                       # [n, 128, hw]
    (info, streams) = native_codec.parse_container(blob)
    assert info['n'] == n and info['h'] == h and info['w'] == wd and info['L'] == 10 and info['bytes'] == blob.size
    # (1) bitstreams byte-identical to the CPU coder on the same indices, every stream
    total = 0
    for s in range(n*128):
        want = oracle_coder.encode_map(idx_gpu.reshape(n*128, -1)[s], params.table[s % 128], WHICH)
        assert want[0] == 0
        (bb, rb, bac, byp) = streams[s]
        assert (bb, rb) == (want[2], want[4])
        assert numpy.array_equal(bac, want[1]) and numpy.array_equal(byp, want[3])
        total += bb + rb
    assert stats['total_bits'] == total
    assert numpy.array_equal(stats['bits_per_map'],
                             numpy.array([sum(streams[i*128 + m][0] + streams[i*128 + m][1] for i in range(n))
                                          for m in range(128)], dtype=numpy.uint64))
    # (2) indices vs the oracle's
    (y_ref, idx_ref, rec_ref) = oracle_pipeline(lum, w, learned, params)
    idx_ref_planar = idx_ref.reshape(n, -1, 128).transpose(0, 2, 1)
    # North star: at least 99.99 % of the indices agree (294 912 coefficients here: at most 29 may differ)
    mismatches = int((idx_gpu != idx_ref_planar).sum())
    assert mismatches <= 1e-4*idx_gpu.size, (mismatches, idx_gpu.size)
    assert numpy.abs(idx_gpu.astype(numpy.int32) - idx_ref_planar).max() <= 1
    dead_ref = int(oracle_glue.count_nb_deads(idx_ref.astype(numpy.float32)).sum())
    assert abs(stats['nb_dead_maps'] - dead_ref) <= 1
    # (3) decompress: exact inverse of the coder, reconstruction close to the oracle's
    rec = codec.decompress(blob, params)
    assert rec.shape == (n, h, wd) and rec.dtype == numpy.uint8
    assert numpy.array_equal(codec.last_indices(n, h, wd), idx_gpu)
    for i in range(n):
        assert abs(oracle_glue.psnr_2d(lum[i], rec[i]) - oracle_glue.psnr_2d(lum[i], rec_ref[i])) < 0.01
    # rate within 0.1 % of the oracle's coder on the oracle's indices
    bits_ref = 0
    for i in range(n):
        for m in range(128):
            bits_ref += oracle_coder.compress_lossless(idx_ref[i, :, :, m].flatten(), params.table[m], WHICH)[2]
    assert abs(total - bits_ref) <= 1e-3*bits_ref


def edge_distance(y64, mean, delta):
    """Distance of every coefficient of the float64 latent to the nearest bin edge, in units of its bin width."""
    t = (y64 - mean.reshape((1, 1, 1, -1)))/delta.reshape((1, 1, 1, -1))
    return numpy.abs(numpy.abs(t - numpy.floor(t)) - 0.5)


def test_config2_every_stream_equals_the_compiled_reference(native, golden):
    """BASELINE config 2 at full size: 24 synthetic 512 x 768 images through Codec.compress in the bench's
    arithmetic; BOTH byte buffers and bit counts of ALL 24 x 128 coded streams equal the reference's own C++ coder
    (oracle/_ref, compression.cpp:24-49) on the same int16 indices, and the coder round-trips. The north star's index
    clause is measured on the same batch: the GPU indices against the float64 oracle, with the float32 oracle's own
    count against float64 beside it (no float32 evaluation order can do better than that)."""
    rng = numpy.random.default_rng(1)
    w = visible_weights(0, False)
    (n, h, wd) = (24, 512, 768)
    lum = util.synthetic_luma(rng, n, h, wd)
    mean = golden.map_mean('1_10000').astype(numpy.float32)
    delta = numpy.ones(128, dtype=numpy.float32)
    params = native_codec.CodingParams(delta, golden.table('1_10000', '1'), mean)
    codec = native_codec.Codec(w, False, math='mixed')
    (blob, stats) = codec.compress(lum, params, return_stats=True)
    idx = codec.last_indices(n, h, wd).reshape(n*128, -1)
    (info, streams) = native_codec.parse_container(blob)
    assert info['bytes'] == blob.size and len(streams) == n*128
    total = 0
    for s in range(n*128):
        want = oracle_coder.encode_map(idx[s], params.table[s % 128], WHICH)
        assert want[0] == 0 and (streams[s][0], streams[s][1]) == (want[2], want[4]), s
        assert numpy.array_equal(streams[s][2], want[1]) and numpy.array_equal(streams[s][3], want[3]), s
        total += want[2] + want[4]
    assert stats['total_bits'] == total
    rec = codec.decompress(blob, params)
    assert numpy.array_equal(codec.last_indices(n, h, wd).reshape(n*128, -1), idx)
    # ---- the index clause, at batch scale
    y64 = numpy.concatenate([T.encoder(lum[i:i + 4, :, :, None].astype(numpy.float64), w, False, dtype=torch.float64)
                             for i in range(0, n, 4)])
    (y32, idx32, rec32) = oracle_pipeline(lum, w, False, params)
    k64 = numpy.rint((y64 - mean.astype(numpy.float64).reshape((1, 1, 1, -1)))/delta.reshape((1, 1, 1, -1)))
    k_gpu = idx.reshape(n, 128, -1).transpose(0, 2, 1).reshape(k64.shape)
    edge = edge_distance(y64, mean.astype(numpy.float64), delta.astype(numpy.float64))
    bad_gpu = k_gpu != k64
    bad_32 = idx32 != k64
    report = {'coefficients': int(k64.size), 'gpu_vs_fp64': int(bad_gpu.sum()), 'fp32_oracle_vs_fp64': int(bad_32.sum()),
              'gpu_vs_fp32_oracle': int((k_gpu != idx32).sum()),
              'gpu_worst_edge_distance': float(edge[bad_gpu].max()) if bad_gpu.any() else 0.,
              'fp32_oracle_worst_edge_distance': float(edge[bad_32].max()) if bad_32.any() else 0.}
    print('config 2 index clause:', report)
    assert report['gpu_vs_fp32_oracle'] <= 1e-4*k64.size and report['gpu_vs_fp64'] <= 1e-4*k64.size, report
    assert numpy.abs(k_gpu - k64).max() <= 1
    assert report['gpu_worst_edge_distance'] < 1e-4, report
    # PSNR within 0.01 dB and rate within 0.1 % of the oracle pipeline (its own indices, its own decoder)
    for i in range(n):
        assert abs(oracle_glue.psnr_2d(lum[i], rec[i]) - oracle_glue.psnr_2d(lum[i], rec32[i])) < 0.01
    bits_ref = sum(oracle_coder.compress_lossless(idx32[i, :, :, m].flatten(), params.table[m], WHICH)[2]
                   for i in range(n) for m in range(128))
    assert abs(total - bits_ref) <= 1e-3*bits_ref


@pytest.mark.parametrize('math', CODEC_MODES)
def test_quantization_sweep(native, golden, math):
    """BASELINE config 3 at test scale: one EAE, bin widths delta x {1, 2, 4, 8}."""
    rng = numpy.random.default_rng(4)
    w = visible_weights(0, False)
    lum = util.synthetic_luma(rng, 2, 128, 128)
    codec = native_codec.Codec(w, False, math=math)
    rates = []
    psnrs = []
    for mult in (1, 2, 4, 8):
        params = native_codec.CodingParams(mult*numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', str(mult)))
        (blob, stats) = codec.compress(lum, params, return_stats=True)
        rec = codec.decompress(blob, params)
        (_, idx_ref, rec_ref) = oracle_pipeline(lum, w, False, params)
        bits_ref = sum(oracle_coder.compress_lossless(idx_ref[i, :, :, m].flatten(), params.table[m], WHICH)[2]
                       for i in range(2) for m in range(128))
        assert abs(stats['total_bits'] - bits_ref) <= 1e-3*bits_ref
        for i in range(2):
            assert abs(oracle_glue.psnr_2d(lum[i], rec[i]) - oracle_glue.psnr_2d(lum[i], rec_ref[i])) < 0.01
        rates.append(stats['total_bits'])
        psnrs.append(oracle_glue.psnr_2d(lum[0], rec[0]))
    assert rates == sorted(rates, reverse=True)       # coarser bins never cost more bits


def test_4k_frame_through_the_whole_codec(native, golden):
    """BASELINE config 5 at test scale: one 2160 x 3840 frame (latent 135 x 240, not a multiple of any tile; 128 streams
    of 32 400 symbols) through compress -> container -> decompress in the bench's default arithmetic."""
    rng = numpy.random.default_rng(6)
    w = visible_weights(1, False)
    (h, wd) = (2160, 3840)
    lum = util.synthetic_luma(rng, 1, h, wd)
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1'))
    codec = native_codec.Codec(w, False, math='mixed')
    (blob, stats) = codec.compress(lum, params, return_stats=True)
    idx = codec.last_indices(1, h, wd).reshape(128, -1)
    assert idx.shape[1] == 135*240
    (info, streams) = native_codec.parse_container(blob)
    assert (info['n'], info['h'], info['w']) == (1, h, wd)
    total = 0
    for s in range(0, 128, 7):          # every 7th stream byte for byte against the CPU coder
        want = oracle_coder.encode_map(idx[s], params.table[s], WHICH)
        assert want[0] == 0 and (streams[s][0], streams[s][1]) == (want[2], want[4])
        assert numpy.array_equal(streams[s][2], want[1]) and numpy.array_equal(streams[s][3], want[3])
    for s in range(128):
        total += streams[s][0] + streams[s][1]
    assert stats['total_bits'] == total
    y32 = T.encoder(lum[..., None].astype(numpy.float32), w, False)
    idx_ref = numpy.round(y32).astype(numpy.int16).reshape(-1, 128).T
    assert (idx != idx_ref).mean() <= 1e-4
    rec = codec.decompress(blob, params)
    assert numpy.array_equal(codec.last_indices(1, h, wd).reshape(128, -1), idx)       # the coder inverts exactly
    q = idx.T.reshape(1, 135, 240, 128).astype(numpy.float32)
    want_rec = oracle_glue.cast_bt601(T.decoder(q, w, False))[0, :, :, 0]
    diff = numpy.abs(rec[0].astype(numpy.int32) - want_rec.astype(numpy.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-2
    assert abs(oracle_glue.psnr_2d(lum[0], rec[0]) - oracle_glue.psnr_2d(lum[0], want_rec)) < 0.01


def test_repeated_steps_replay_a_graph_with_identical_results(native, golden):
    """From its second use with the same buffers, a step of the device-resident entry points (eae_compress_dev /
    eae_decompress_dev) is captured into a CUDA graph and replayed (csrc/codec.cu, run_as_step_graph): same bytes, same
    pixels, same launch count as the host entry points, which launch directly; a new image or new coding parameters behind
    the same pointers must show up in the replayed step."""
    import ctypes
    rng = numpy.random.default_rng(8)
    w = visible_weights(2, False)
    (n, h, wd) = (2, 128, 192)
    images = [util.synthetic_luma(rng, n, h, wd) for _ in range(2)]
    prms = [native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1')),
            native_codec.CodingParams(2*numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '2'))]
    lib = native.lib()
    codec = native_codec.Codec(w, False, math='mixed', own_stream=True)
    plain = native_codec.Codec(w, False, math='mixed')            # host entry points: no graphs
    bound = int(lib.eae_container_bound(n, h, wd, 10))
    dev = torch.device('cuda', 0)
    d_img = torch.zeros((n, h, wd), dtype=torch.uint8, device=dev)
    d_cont = torch.zeros(bound, dtype=torch.uint8, device=dev)
    d_rec = torch.zeros((n, h, wd), dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_stats = torch.zeros(ctypes.sizeof(native.BatchStats), dtype=torch.uint8, device=dev)
    launches = []
    for (k, (i, j)) in enumerate([(0, 0)]*5 + [(1, 0), (1, 1), (0, 1), (0, 0)]):
        d_img.copy_(torch.from_numpy(images[i]))
        torch.cuda.synchronize()
        prm = prms[j].native()
        before = lib.eae_launch_count()
        native.check(lib.eae_compress_dev(codec.handle, ctypes.byref(prm), d_img.data_ptr(), n, h, wd, d_cont.data_ptr(), bound,
                                          d_total.data_ptr(), d_stats.data_ptr(), codec.stream))
        native.check(lib.eae_decompress_dev(codec.handle, ctypes.byref(prm), d_cont.data_ptr(), bound, n, h, wd,
                                            d_rec.data_ptr(), codec.stream))
        native.check(lib.eae_stream_synchronize(codec.stream))
        launches.append(lib.eae_launch_count() - before)
        want = numpy.array(plain.compress(images[i], prms[j]), copy=True)
        size = int(d_total.cpu()[0])
        assert size == want.size and numpy.array_equal(d_cont.cpu().numpy()[:size], want), (k, i, j)
        assert numpy.array_equal(d_rec.cpu().numpy(), plain.decompress(want, prms[j])), (k, i, j)
    assert len(set(launches[1:5])) == 1, launches      # direct, captured and replayed steps count the same launches
    # host-side state after a replayed step: the indices of the last step can be fetched
    plain.compress(images[0], prms[0])
    assert numpy.array_equal(codec.last_indices(n, h, wd), plain.last_indices(n, h, wd))


def test_batches_larger_than_the_workspace_chunk(native, golden, monkeypatch):
    """BASELINE config 4 shape at test scale: a batch that does not fit the activation workspace is transformed in chunks
    (csrc/codec.cu, chunk_images) while the coder sees all its streams at once; the container must not depend on the
    chunking. EAE_CHUNK_IMAGES shrinks the chunk so that 5 images take three passes (2 + 2 + 1)."""
    rng = numpy.random.default_rng(9)
    w = visible_weights(3, False)
    lum = util.synthetic_luma(rng, 5, 64, 96)
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1'))
    whole = native_codec.Codec(w, False, math='mixed')
    blob = numpy.array(whole.compress(lum, params), copy=True)
    rec = numpy.array(whole.decompress(blob, params), copy=True)
    monkeypatch.setenv('EAE_CHUNK_IMAGES', '2')
    chunked = native_codec.Codec(w, False, math='mixed')
    assert numpy.array_equal(chunked.compress(lum, params), blob)
    assert numpy.array_equal(chunked.decompress(blob, params), rec)
    assert numpy.array_equal(chunked.encode(lum[..., None]), whole.encode(lum[..., None]))


def test_container_errors(native, golden):
    w = wts.random_init(0, True)
    codec = native_codec.Codec(w, True)
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1'))
    lum = util.synthetic_luma(numpy.random.default_rng(5), 1, 64, 64)
    blob = codec.compress(lum, params).copy()
    with pytest.raises(RuntimeError):       # truncated payload -> resource error (code 2)
        codec.decompress(blob[:-40], params)
    bad = blob.copy()
    bad[0] ^= 0xFF
    with pytest.raises(ValueError):
        codec.decompress(bad, params)
    with pytest.raises(ValueError):         # bin width <= 0 (tools.py:924-925)
        codec.compress(lum, native_codec.CodingParams(numpy.zeros(128, dtype=numpy.float32), golden.table('1_10000', '1')))
    nan_table = golden.table('1_10000', '1').copy()
    nan_table[:, 0] = numpy.nan
    with pytest.raises(RuntimeError, match='Error of type 4'):
        codec.compress(lum, native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), nan_table))
    with pytest.raises(ValueError):         # EntropyAutoencoder.py:77-80
        codec.compress(numpy.zeros((1, 40, 64), dtype=numpy.uint8), params)


def test_device_entry_points_report_through_poll_status(native, golden):
    """The asynchronous entry points cannot return what happens on the device; eae_codec_poll_status does: a clean step,
    a container that does not fit its capacity, a garbled stream table and a truncated container (validated on the device:
    nothing outside the bytes the caller vouches for is read), an index outside int16."""
    import ctypes
    rng = numpy.random.default_rng(10)
    w = visible_weights(4, False)
    (n, h, wd) = (2, 64, 96)
    lum = util.synthetic_luma(rng, n, h, wd)
    prm_obj = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1'))
    prm = prm_obj.native()
    lib = native.lib()
    codec = native_codec.Codec(w, False, math='mixed', own_stream=True)
    bound = int(lib.eae_container_bound(n, h, wd, 10))
    dev = torch.device('cuda', 0)
    d_img = torch.from_numpy(lum).to(dev)
    d_cont = torch.zeros(bound, dtype=torch.uint8, device=dev)
    d_rec = torch.zeros((n, h, wd), dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)

    def compress(cap):
        native.check(lib.eae_compress_dev(codec.handle, ctypes.byref(prm), d_img.data_ptr(), n, h, wd, d_cont.data_ptr(), cap,
                                          d_total.data_ptr(), None, codec.stream))

    def decompress(buf, nbytes):
        native.check(lib.eae_decompress_dev(codec.handle, ctypes.byref(prm), buf.data_ptr(), nbytes, n, h, wd,
                                            d_rec.data_ptr(), codec.stream))

    compress(bound)
    decompress(d_cont, bound)
    status = codec.poll_status()
    assert status['code'] == 0 and not any(status[k] for k in status if k != 'code'), status
    size = int(d_total.cpu()[0])
    good = d_cont[:size].clone()
    want_rec = d_rec.cpu().numpy().copy()
    # the exact size is enough
    decompress(good, size)
    assert codec.poll_status()['code'] == 0 and numpy.array_equal(d_rec.cpu().numpy(), want_rec)
    # (1) capacity too small for the payload: the needed size is still reported, the status says so
    compress(32 + 8*128*n + 64)
    status = codec.poll_status(raise_on_error=False)
    assert status['container_overflow'] == 1 and status['code'] == native.ERR_ARGUMENT and int(d_total.cpu()[0]) == size
    assert codec.poll_status()['code'] == 0           # the record is cleared by a poll
    # (2) truncated container: the streams that would run past the end are read as empty (never past the bytes the
    #     caller vouches for) and the status reports the resource error (code 2) the host entry point returns
    decompress(good, size - 40)
    status = codec.poll_status(raise_on_error=False)
    assert status['container_invalid'] == 1 and status['code'] == 2, status
    # (3) a bit count no encoder can have written
    bad = good.clone()
    table = bad[32:32 + 8].view(torch.int32)
    table[0] = 1 << 30
    decompress(bad, size)
    status = codec.poll_status(raise_on_error=False)
    assert status['container_invalid'] == 1 and status['code'] == 1, status
    with pytest.raises(RuntimeError):
        decompress(bad, size)
        codec.poll_status()
    # (4) an index outside int16 (tools.py:126-133): bin widths of 1e-6
    tiny = native_codec.CodingParams(1e-6*numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1'))
    prm_tiny = tiny.native()
    native.check(lib.eae_compress_dev(codec.handle, ctypes.byref(prm_tiny), d_img.data_ptr(), n, h, wd, d_cont.data_ptr(), bound,
                                      d_total.data_ptr(), None, codec.stream))
    status = codec.poll_status(raise_on_error=False)
    assert status['int16_overflow'] == 1 and status['code'] == native.ERR_INT16_RANGE, status
    # and the codec still works afterwards
    compress(bound)
    decompress(d_cont, bound)
    assert codec.poll_status()['code'] == 0 and numpy.array_equal(d_rec.cpu().numpy(), want_rec)


def test_every_packing_of_the_coder_kernels_gives_the_same_container(native, golden):
    """GPU threads per coded stream (eae_codec_set_coder_lanes): 1 = 32 streams per warp (branch-free formulation, the
    bench's throughput setting), 2 / 4 = partial packing, 32 = one stream per warp (the scalar, branching formulation of
    csrc/coder_core.cuh: the latency setting), 0 = automatic. Same bytes from all of them, every stream equal to the
    reference coder on the same indices, exact round trip - on peaked and on wide symbol distributions (bin widths 1 and
    1/8: long runs of truncated-unary bins and Exp-Golomb suffixes)."""
    rng = numpy.random.default_rng(15)
    w = visible_weights(6, False)
    (n, h, wd) = (3, 128, 192)
    lum = util.synthetic_luma(rng, n, h, wd)
    for delta in (1., 0.125):
        params = native_codec.CodingParams(delta*numpy.ones(128, dtype=numpy.float32), golden.table('1_10000', '1'),
                                           0.05*golden.map_mean('1_10000'))
        blobs = []
        for lanes in (0, 1, 2, 4, 32):
            codec = native_codec.Codec(w, False, math='mixed')
            codec.set_coder_lanes(lanes)
            blob = numpy.array(codec.compress(lum, params), copy=True)
            idx = codec.last_indices(n, h, wd)
            rec = numpy.array(codec.decompress(blob, params), copy=True)
            assert numpy.array_equal(codec.last_indices(n, h, wd), idx), lanes       # the decoder inverts the encoder
            blobs.append((blob, rec))
        for (blob, rec) in blobs[1:]:
            assert numpy.array_equal(blob, blobs[0][0]) and numpy.array_equal(rec, blobs[0][1])
        (_, streams) = native_codec.parse_container(blobs[0][0])
        flat = idx.reshape(n*128, -1)
        for s in range(n*128):
            want = oracle_coder.encode_map(flat[s], params.table[s % 128], WHICH)
            assert want[0] == 0 and (streams[s][0], streams[s][1]) == (want[2], want[4]), (delta, s)
            assert numpy.array_equal(streams[s][2], want[1]) and numpy.array_equal(streams[s][3], want[3]), (delta, s)

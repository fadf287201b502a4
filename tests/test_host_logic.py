"""Host-side logic of the Python mirror of the reference interface: argument validation raises the
reference's exceptions before anything reaches the GPU; sharding helpers; container parsing."""
import numpy
import pytest

from autoencoder_based_image_compression_b200 import codec as native_codec
from autoencoder_based_image_compression_b200 import parallel
from autoencoder_based_image_compression_b200 import weights as wts
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae import batching
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph import constants as csts
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.EntropyAutoencoder import EntropyAutoencoder
from autoencoder_based_image_compression_b200.kodak_tensorflow.eae.graph.IsolatedDecoder import IsolatedDecoder
from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import compression, interface_cython
from autoencoder_based_image_compression_b200.kodak_tensorflow.tools import tools as tls


def test_constants():
    assert (csts.NB_MAPS_1, csts.NB_MAPS_2, csts.NB_MAPS_3) == (128, 128, 128)
    assert (csts.WIDTH_KERNEL_1, csts.WIDTH_KERNEL_2, csts.WIDTH_KERNEL_3) == (9, 5, 5)
    assert csts.STRIDE_PROD == 16


def test_weights_random_init_and_npz(tmp_path):
    for learned in (False, True):
        w = wts.random_init(0, learned, bin_width_init=0.5)
        wts.validate(w, learned)
        assert w['encoder/weights_1'].shape == (9, 9, 1, 128) and w['decoder/weights_6'].shape == (9, 9, 1, 128)
        assert numpy.allclose(w['encoder/gamma_1'], w['encoder/gamma_1'].T) and w['encoder/gamma_1'].min() >= 2e-5
        assert numpy.all(w[wts.BIN_WIDTHS_KEY] == numpy.float32(0.5))
        assert ('decoder/gamma_4' in w) == (not learned)
    path = str(tmp_path/'model.npz')
    wts.save(path, w)
    back = wts.load(path)
    assert sorted(back) == sorted(w) and all(numpy.array_equal(back[k], w[k]) for k in w)
    del w['decoder/weights_6']
    with pytest.raises(KeyError):
        wts.validate(w, True)


def test_model_classes_validate_like_the_reference():
    with pytest.raises(ValueError):     # EntropyAutoencoder.py:77-80
        EntropyAutoencoder(4, 500, 768, 1., 10000., '', False)
    with pytest.raises(ValueError):     # IsolatedDecoder.py:50-53
        IsolatedDecoder(4, 512, 770, False)
    ae = EntropyAutoencoder(4, 512, 768, 1., 10000., '', True)
    with pytest.raises(RuntimeError):
        ae.get_bin_widths()
    ae.initialization(None, '')
    assert ae.get_bin_widths().dtype == numpy.float32 and ae.get_bin_widths().shape == (128,)


def test_batching_argument_checks():
    ae = EntropyAutoencoder(4, 32, 48, 1., 10000., '', False)
    ae.initialization(None, '')
    with pytest.raises(TypeError):      # batching.py:86-87
        batching.encode_mini_batches(numpy.zeros((4, 32, 48, 1), dtype=numpy.float32), None, ae, 4)
    with pytest.raises(ValueError):     # tools.py:1130-1131
        batching.encode_mini_batches(numpy.zeros((6, 32, 48, 1), dtype=numpy.uint8), None, ae, 4)
    with pytest.raises(ValueError):     # unpacking of a 3D array (batching.py:91)
        batching.encode_mini_batches(numpy.zeros((4, 32, 48), dtype=numpy.uint8), None, ae, 4)
    dec = IsolatedDecoder(4, 32, 48, False)
    dec.initialization(None, '')
    with pytest.raises(ValueError):
        batching.decode_mini_batches(numpy.zeros((3, 2, 3, 128), dtype=numpy.float32), None, dec, 2)


def test_tools_argument_checks():
    with pytest.raises(ValueError):     # tools.py:917-918
        tls.quantize_per_map(numpy.zeros((1, 2, 2, 4), dtype=numpy.float32), numpy.ones((4, 1), dtype=numpy.float32))
    with pytest.raises(ValueError):     # :922-923
        tls.quantize_per_map(numpy.zeros((1, 2, 2, 4), dtype=numpy.float32), numpy.ones(3, dtype=numpy.float32))
    with pytest.raises(ValueError):     # :924-925
        tls.quantize_per_map(numpy.zeros((1, 2, 2, 4), dtype=numpy.float32), numpy.array([1, 1, 0, 1], dtype=numpy.float32))
    with pytest.raises(TypeError):      # :91-92
        tls.cast_bt601(numpy.zeros(4, dtype=numpy.int32))
    with pytest.raises(TypeError):      # :124-125
        tls.cast_float_to_int16(numpy.zeros(4, dtype=numpy.uint8))
    with pytest.raises(TypeError):      # :866-867
        tls.psnr_2d(numpy.zeros((2, 2), dtype=numpy.float32), numpy.zeros((2, 2), dtype=numpy.uint8))
    with pytest.raises(ValueError):     # :870-873
        tls.psnr_2d(numpy.zeros((2, 2, 1), dtype=numpy.uint8), numpy.zeros((2, 2, 1), dtype=numpy.uint8))
    with pytest.raises(ValueError):     # :313-314
        tls.count_nb_deads(numpy.zeros((2, 2, 2), dtype=numpy.float32))
    with pytest.raises(ValueError):
        tls.subdivide_set(10, 4)
    assert tls.subdivide_set(24, 4) == 6


def test_lossless_argument_checks(tmp_path):
    with pytest.raises(TypeError):      # compression.py:52-53
        compression.compress_lossless_maps(numpy.zeros((2, 2, 3), dtype=numpy.int32), 'unused.npy')
    path = str(tmp_path/'table.npy')
    numpy.save(path, numpy.full((4, 10), 0.5))
    with pytest.raises(ValueError):     # :63-64
        compression.compress_lossless_maps(numpy.zeros((2, 2, 3), dtype=numpy.int16), path)
    # a table rewritten in the same process is re-read, as the reference re-reads it for every image (:60): the cache is
    # keyed on the file's identity, not only on its path (no clearing needed)
    import os
    numpy.save(path, numpy.full(10, 0.5))
    os.utime(path, ns=(1, 1))
    with pytest.raises(ValueError):     # :61-62
        compression.compress_lossless_maps(numpy.zeros((2, 2, 3), dtype=numpy.int16), path)
    with pytest.raises(ValueError):     # :129-130
        compression.rescale_compress_lossless_maps(numpy.zeros((2, 2, 3), dtype=numpy.float32),
                                                   numpy.ones((3, 1), dtype=numpy.float32), path)
    with pytest.raises(ValueError):     # :135-136
        compression.rescale_compress_lossless_maps(numpy.zeros((2, 2, 3), dtype=numpy.float32),
                                                   numpy.ones(4, dtype=numpy.float32), path)
    with pytest.raises(ValueError):     # Cython buffer dtype mismatch
        interface_cython.compress_lossless_flattened_map(numpy.zeros(4, dtype=numpy.int32), numpy.full(3, 0.5))
    with pytest.raises(ValueError):
        interface_cython.compress_lossless_flattened_map(numpy.zeros((2, 2), dtype=numpy.int16), numpy.full(3, 0.5))
    with pytest.raises(OverflowError):  # interface_cython.pyx:50-52
        interface_cython.compress_lossless_flattened_map(numpy.zeros(4, dtype=numpy.int16), numpy.full(256, 0.5))


def test_coding_params_checks():
    with pytest.raises(ValueError):
        native_codec.CodingParams(numpy.ones(128), numpy.full(10, 0.5))
    with pytest.raises(ValueError):
        native_codec.CodingParams(numpy.ones(128), numpy.full((3, 10), 0.5))
    p = native_codec.CodingParams(numpy.ones(128), numpy.full((128, 10), 0.5), numpy.zeros(128))
    assert p.truncated_unary_length == 10 and p.native().truncated_unary_length == 10


def test_container_parser():
    n_streams = 128
    table = numpy.zeros((n_streams, 2), dtype=numpy.uint32)
    table[:, 0] = 9
    table[:, 1] = numpy.arange(n_streams) % 17
    payload = bytearray()
    for s in range(n_streams):
        payload += bytes([s % 251])*2 + bytes([7])*((int(table[s, 1]) + 7)//8)
    hdr = numpy.array([0x42454145, 1, 1, 16, 16, 128, 10, 0], dtype=numpy.uint32)
    blob = numpy.frombuffer(hdr.tobytes() + table.tobytes() + bytes(payload), dtype=numpy.uint8)
    (info, streams) = native_codec.parse_container(blob)
    assert info == {'n': 1, 'h': 16, 'w': 16, 'nb_maps': 128, 'L': 10, 'bytes': blob.size}
    assert streams[5][0] == 9 and streams[5][2].tolist() == [5, 5] and len(streams[5][3]) == 1
    with pytest.raises(ValueError):
        native_codec.parse_container(numpy.zeros(64, dtype=numpy.uint8))


def test_shard_range_covers_everything_once():
    for (n, world) in ((4096, 8), (24, 8), (5, 8), (256, 3), (0, 2)):
        seen = []
        for r in range(world):
            (a, b) = parallel.shard_range(n, r, world)
            seen += list(range(a, b))
        assert seen == list(range(n))
    sizes = [parallel.shard_range(24, r, 5)[1] - parallel.shard_range(24, r, 5)[0] for r in range(5)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_stats_pack_and_summary():
    # two images with PSNRs 30 and 40 dB: the reference averages the per-image PSNRs (reconstructing_eae_kodak.py:810-815)
    vec = parallel.pack_stats(numpy.arange(128), 8128, 3, 1000., 512*768*2, 2, 30. + 40.)
    s = parallel.summarize(parallel.unpack_stats(vec))
    assert s['total_bits'] == 8128 and s['nb_dead_maps'] == 3 and s['nb_images'] == 2
    assert abs(s['rate_bpp'] - 8128/(512*768*2)) < 1e-15
    assert s['psnr_db'] == 35.
    assert abs(s['psnr_db_pooled'] - 10*numpy.log10(255**2/(1000./(512*768*2)))) < 1e-12


def test_drop_in_module_names_resolve(tmp_path):
    """with the mirror directory on sys.path the reference's own import lines work."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, {root!r}); "
            "sys.path.insert(0, {root!r} + '/autoencoder_based_image_compression_b200/kodak_tensorflow'); "
            "import eae.batching, lossless.compression, lossless.interface_cython, lossless.stats, tools.tools as tls; "
            "import tensorflow as tf; assert callable(tf.reset_default_graph) and hasattr(tf.Session(), '__enter__'); "
            "assert callable(tls.visualize_rotated_luminance); "
            "assert tls.float_to_str(0.5) == '0dot5' and tls.float_to_str(10000.) == '10000' and tls.float_to_str(-1.5) == 'minus1dot5'; "
            "from eae.graph.EntropyAutoencoder import EntropyAutoencoder; "
            "from eae.graph.IsolatedDecoder import IsolatedDecoder; "
            "import eae.graph.constants as csts; "
            "assert csts.STRIDE_PROD == 16 and callable(eae.batching.encode_mini_batches) "
            "and callable(lossless.compression.rescale_compress_lossless_maps) and callable(tls.quantize_per_map) "
            "and callable(lossless.stats.save_statistics) and callable(lossless.stats.find_index_map_exception)"
            ).format(root=root)
    subprocess.check_call([sys.executable, '-c', code], cwd=str(tmp_path))


def test_statistics_argument_checks(tmp_path):
    """lossless/stats.py checks that do not need the device."""
    from autoencoder_based_image_compression_b200.kodak_tensorflow.lossless import stats
    y = numpy.zeros((1, 2, 2, 128), dtype=numpy.float32)
    with pytest.raises(TypeError):
        stats.compute_binary_probabilities(y.astype(numpy.float64), numpy.ones(128), numpy.zeros(128), 10)
    with pytest.raises(ValueError):
        stats.compute_binary_probabilities(y, numpy.zeros(128), numpy.zeros(128), 10)      # bin width 0
    with pytest.raises(ValueError):
        stats.count_binary_decisions(numpy.array([0.5, -0.5], dtype=numpy.float32), 0.5, 4)
    with pytest.raises(ValueError):      # stats.py:296-297
        stats.save_statistics(None, None, None, 1, numpy.ones(2), 10, str(tmp_path/'a.npy'), str(tmp_path/'b.pkl'), ['x.npy'])
    with pytest.raises(ValueError):      # tools.py:652-653
        stats.jensen_shannon_divergence(numpy.array([0., 1.]), numpy.array([0.5, 0.5]))
    assert abs(stats.jensen_shannon_divergence(numpy.array([0.5, 0.5]), numpy.array([0.5, 0.5]))) < 1e-15


def test_last_layer_tile_gather_covers_every_pixel_once_in_col2im_order():
    """Index arithmetic of the fused last layer (csrc/conv_umma.cu, kernel versions 6 / 7), restated in numpy: a tile of
    8 x 16 positions starting two positions before its first pixel block completes the 6 x 14 pixel blocks whose
    contributing positions lie inside it; the per-pixel sum runs in col2im_k9s4_kernel's order (da, db ascending), so
    the two paths must agree bit for bit - for sizes that are not multiples of the tile as well."""
    rng = numpy.random.default_rng(7)
    for (H1, W1) in ((4, 4), (8, 12), (13, 31)):
        (H, W) = (4*H1, 4*W1)
        P = rng.standard_normal((H1, W1, 81)).astype(numpy.float32)

        def tap(a, b, ky, kx):
            return P[a, b, ky*9 + kx] if 0 <= a < H1 and 0 <= b < W1 else numpy.float32(0.)

        want = numpy.zeros((H, W), dtype=numpy.float32)        # col2im_k9s4_kernel (transforms_simt.cu)
        for oy in range(H):
            for ox in range(W):
                acc = numpy.float32(0.)
                (a_hi, b_hi) = ((oy + 2) >> 2, (ox + 2) >> 2)
                for da in range(3):
                    (a, ky) = (a_hi - da, oy + 2 - 4*(a_hi - da))
                    if a < 0 or a >= H1 or ky > 8:
                        continue
                    for db in range(3):
                        (b, kx) = (b_hi - db, ox + 2 - 4*(b_hi - db))
                        if b < 0 or b >= W1 or kx > 8:
                            continue
                        acc = numpy.float32(acc + P[a, b, ky*9 + kx])
                want[oy, ox] = acc
        got = numpy.full((H, W), numpy.nan, dtype=numpy.float32)
        written = numpy.zeros((H, W), dtype=numpy.int32)
        (tiles_y, tiles_x) = ((H1 + 1 + 5)//6, (W1 + 1 + 13)//14)
        for ty in range(tiles_y):
            for tx in range(tiles_x):
                (q0, p0) = (ty*6, tx*14)
                (a0, b0) = (q0 - 2, p0 - 2)
                col = numpy.zeros((8, 16, 81), dtype=numpy.float32)      # TMA zero-fills outside the input
                for la in range(8):
                    for lb in range(16):
                        if 0 <= a0 + la < H1 and 0 <= b0 + lb < W1:
                            col[la, lb] = P[a0 + la, b0 + lb]
                for item in range(4*6*2*14):
                    (ly, pr) = divmod(item, 28)
                    (qa, rr) = (ly >> 2, ly & 3)
                    (pb, s0) = (pr >> 1, (pr & 1)*2)
                    (oy, ox) = (4*(q0 + qa) + rr - 2, 4*(p0 + pb) + s0 - 2)
                    if oy < 0 or oy >= H or ox < 0 or ox >= W:
                        continue
                    acc = [numpy.float32(0.), numpy.float32(0.)]
                    for da in range(3):
                        ky = rr + 4*da
                        if ky > 8:
                            continue
                        for db in range(3):
                            for e in range(2):
                                kx = s0 + 4*db + e
                                if kx <= 8:
                                    acc[e] = numpy.float32(acc[e] + col[qa + 2 - da, pb + 2 - db, ky*9 + kx])
                    got[oy, ox:ox + 2] = acc
                    written[oy, ox:ox + 2] += 1
        assert (written == 1).all()
        assert numpy.array_equal(got, want)


def test_the_reference_driver_imports_unmodified_on_the_mirror_packages():
    """CPU half of tests/test_gpu_reference_driver.py: the reference's own script, read from where it lies, imports with
    the mirror directory on sys.path (its tensorflow / eae / lossless / tools imports resolve to this package)."""
    from tests import refdrivers
    rek = refdrivers.load('reconstructing_eae_kodak.py')
    if rek is None:
        pytest.skip('reference driver not available')
    assert callable(rek.fix_gamma) and callable(rek.vary_gamma_fix_bin_widths)
    assert rek.tf.__file__.startswith(refdrivers.MIRROR)
    assert rek.eae.batching.__file__.startswith(refdrivers.MIRROR)
    assert rek.lossless.compression.__file__.startswith(refdrivers.MIRROR)
    assert rek.tls.__file__.startswith(refdrivers.MIRROR)
    stats_driver = refdrivers.load('collecting_stats_eae_extra.py')
    assert stats_driver is not None and stats_driver.lossless.stats.__file__.startswith(refdrivers.MIRROR)


def test_restore_path_resolution(tmp_path):
    from autoencoder_based_image_compression_b200 import weights as wts
    w = wts.random_init(0, True)
    wts.save(str(tmp_path / 'model_10.npz'), w)
    assert wts.resolve_path(str(tmp_path / 'model_10.ckpt')) == str(tmp_path / 'model_10.npz')
    assert wts.resolve_path(str(tmp_path / 'model_10')) == str(tmp_path / 'model_10.npz')
    assert sorted(wts.load(wts.resolve_path(str(tmp_path / 'model_10.ckpt')))) == sorted(w)
    with pytest.raises(IOError):
        wts.resolve_path(str(tmp_path / 'model_11.ckpt'))

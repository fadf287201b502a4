"""Headline benchmark: images/s of the full EAE codec hot path (encode -> quantize -> lossless code ->
bitstream -> decode) on synthetic luminance images.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4]

Under torchrun (N > 1) every rank drives one GPU with its own batches (no data-path collective; the per-map rate
statistics accumulate on the device and are reduced over the ranks ONCE, at the end of the timed region, as the reference
reduces them once at the end: reconstructing_eae_kodak.py:810-815). Rank 0 prints ONE JSON line.

A "step" is one pass of the hot path over one batch. --config selects the workload of BASELINE.json's `configs`:
  0  one 512 x 768 image per step (latency case: one warp per coded stream), bin width 1
  1  24 images of 512 x 768 per GPU and step, bin width 1                       (default: the headline)
  2  the same batch, bin widths delta x {1, 2, 4, 8} in turn (one EAE), rate / PSNR per delta against the oracle
  3  4096 images of 512 x 768 in all, sharded over the GPUs (up to 32 per step); the step count follows from the workload
  4  256 frames of 2160 x 3840 in all, sharded over the GPUs (whole-frame semantics, up to 16 frames per step)

  value  images/s with the batch already resident in HBM (eae_compress_dev + eae_decompress_dev)
  e2e    the same through the public host API (Codec.compress / Codec.decompress) from pinned host
         memory, host<->device copies inside the timed region
  parity one untimed batch against the CPU oracle (index mismatches, distance to the bin edges, bitstream identity
         with the reference coder, PSNR and rate differences)
--impl reference times the CPU implementation (oracle restatement of the transforms on torch-CPU +
the reference's own C++ coder from oracle/_ref when present, else the C port) on all host cores.
"""
import argparse
import ctypes
import json
import multiprocessing
import os
import subprocess
import sys
import tempfile
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'images/sec (512x768 luma, encode->bitstream->decode)'
# algorithmic GFLOP per 512 x 768 image (2 * MACs), fixed-delta variant (6 GDN/IGDN)
GFLOP_PER_IMAGE = {'gemm_conv': 5.0332 + 1.2583, 'gemm_tconv': 1.2583 + 5.0332,
                   'gemm_gdn': 2*(0.8053 + 0.2013 + 0.0503), 'gemm_thin': 2*0.5096}
L2_BYTES = 126 << 20

# batch per GPU and step, frame size, bin-width multipliers, GPU threads per coded stream (0 = a warp), pipeline slots,
# images of the whole job (None: K steps of `batch` images per GPU, weak scaling)
CONFIGS = {
    0: {'batch': 1, 'height': 512, 'width': 768, 'deltas': (1,), 'coder_lanes': 0, 'depth': 8, 'total': None,
        'name': 'configs[0]: one synthetic 512x768 luma image per step (latency case, one warp per coded stream)'},
    1: {'batch': 24, 'height': 512, 'width': 768, 'deltas': (1,), 'coder_lanes': 1, 'depth': 12, 'total': None,
        'name': 'configs[1]: batch of 24 synthetic 512x768 luma images per GPU'},
    2: {'batch': 24, 'height': 512, 'width': 768, 'deltas': (1, 2, 4, 8), 'coder_lanes': 1, 'depth': 12, 'total': None,
        'name': 'configs[2]: quantization sweep, one EAE, bin widths delta x {1, 2, 4, 8} on consecutive steps of 24 '
                'synthetic 512x768 luma images per GPU'},
    3: {'batch': 32, 'height': 512, 'width': 768, 'deltas': (1,), 'coder_lanes': 1, 'depth': 12, 'total': 4096,
        'name': 'configs[3]: 4096 synthetic 512x768 luma images in all, sharded over the GPUs, up to 32 per step'},
    4: {'batch': 16, 'height': 2160, 'width': 3840, 'deltas': (1,), 'coder_lanes': 1, 'depth': 8, 'total': 256,
        'name': 'configs[4]: 256 synthetic 2160x3840 luma frames in all (whole-frame semantics, latent 135x240), sharded over '
                'the GPUs, up to 16 per step'},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=96)
    ap.add_argument('--warmup', type=int, default=16)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=1, choices=sorted(CONFIGS), help='workload: index into BASELINE.json configs')
    ap.add_argument('--batch', type=int, default=None, help='images per GPU per step (default: the config\'s)')
    ap.add_argument('--height', type=int, default=None)
    ap.add_argument('--width', type=int, default=None)
    ap.add_argument('--math', default=os.environ.get('EAE_MATH', 'mixed'), choices=['fp32', 'tf32x3', 'tf32', 'mixed'])
    ap.add_argument('--cpu-sample', type=int, default=0, help='images in the CPU sample (0 = one per core)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true', help='skip the untimed parity batch against the CPU oracle')
    ap.add_argument('--blocking-sync', type=int, default=1,
                    help='1: host threads sleep while they wait for the GPU (one thread per pipeline slot and one process '
                         'per GPU would otherwise spin on more threads than the box has cores); 0: driver default')
    ap.add_argument('--coder-lanes', type=int, default=None,
                    help='GPU threads per coded stream (0 = one warp per stream: lowest latency)')
    ap.add_argument('--bin-widths', type=str, default=None,
                    help='comma-separated bin-width multipliers out of 1, 2, 4, 8 (default: the config\'s)')
    ap.add_argument('--depth', type=int, default=None,
                    help='pipeline slots (CUDA streams) that consecutive steps rotate over')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    batch_given = args.batch is not None
    for key in ('batch', 'height', 'width', 'coder_lanes', 'depth'):
        if getattr(args, key) is None:
            setattr(args, key, cfg[key])
    if os.environ.get('EAE_PIPELINE_DEPTH'):
        args.depth = int(os.environ['EAE_PIPELINE_DEPTH'])
    args.deltas = tuple(int(x) for x in args.bin_widths.split(',')) if args.bin_widths else cfg['deltas']
    args.total = cfg['total']
    if args.total and not batch_given:
        # a fixed job sharded over many ranks: keep at least 8 steps per rank, so that the pipeline slots have batches to
        # overlap (the coder of a batch is a latency chain: 256 4K frames over 8 GPUs as 2 steps of 16 run at half the rate)
        world = int(os.environ.get('WORLD_SIZE', '1'))
        args.batch = max(1, min(args.batch, args.total//world//8))
    return args


def load_tables(mult=1):
    with numpy.load(os.path.join(ROOT, 'tests', 'golden', 'tables.npz')) as data:
        return (numpy.ascontiguousarray(data['1_10000__binary_probabilities_{}'.format(mult)]),
                numpy.ascontiguousarray(data['1_10000__map_mean']))


def workload_name(args):
    return ('{}; {}x{}, {} per GPU and step; fixed-delta Kodak EAE (6 GDN/IGDN, random-init weights seed 0), bin width(s) {}, '
            'shipped tables 1_10000/binary_probabilities_<delta>, map_mean 1_10000').format(
                CONFIGS[args.config]['name'], args.height, args.width, args.batch, ' / '.join('%g' % d for d in args.deltas))


def base_config(args):
    """The keys both arms report (the driver compares them)."""
    return {'workload': workload_name(args), 'config': args.config, 'batch_per_gpu': args.batch,
            'height': args.height, 'width': args.width, 'bin_width_multipliers': list(args.deltas)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle / reference on the host cores

def _cpu_code_one(job):
    from oracle import coder
    (planar, table, which) = job
    (err, out, bits) = coder.compress_maps_planar(planar, table, which=which)
    assert err == 0 and numpy.array_equal(out, planar)
    return int(bits.sum())


def cpu_pipeline(images, weights, table, map_mean, delta, cores, pool, which):
    """CPU restatement of one batch: encoder -> centre/quantize -> coder (encode + decode per map, as the
    reference's compress_lossless does) -> decoder -> cast. The transforms run in mini-batches of 4 on all cores
    (torch threads), the coder of ALL images then runs one image per worker process on all cores.
    Returns (seconds, bits, reconstruction)."""
    import torch
    from oracle import glue, transforms
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    rec = numpy.zeros(images.shape, dtype=numpy.uint8)
    mean = map_mean.reshape((1, 1, 1, -1))
    widths = (delta*numpy.ones(128)).astype(numpy.float32)
    (planar, quantized) = ([], [])
    for i0 in range(0, images.shape[0], 4):          # reconstructing_eae_kodak.py:624 batch_size = 4
        x = images[i0:i0 + 4, :, :, None].astype(numpy.float32)
        y = transforms.encoder(x, weights, False)
        cq = glue.quantize_per_map(y - mean, widths)
        idx = glue.cast_float_to_int16(cq/widths.reshape((1, 1, 1, -1)))
        planar += [numpy.ascontiguousarray(idx[j].reshape(-1, 128).T) for j in range(idx.shape[0])]
        quantized.append(cq)
    bits = sum(pool.map(_cpu_code_one, [(p, table, which) for p in planar]))
    for (k, cq) in enumerate(quantized):
        rec[4*k:4*k + 4] = glue.cast_bt601(transforms.decoder(cq + mean, weights, False))[..., 0]
    return (time.perf_counter() - t0, bits, rec)


def run_cpu_arm(args, steps, warmup, sample_per_core=1):
    from autoencoder_based_image_compression_b200 import synthetic
    from autoencoder_based_image_compression_b200 import weights as wts
    from oracle import coder
    cores = os.cpu_count() or 1
    # images per step: a multiple of the core count (one image per worker process); the cpu_baseline leg of the GPU
    # arm runs ONE step of 16 images per core (about 10 s of CPU work), the reference arm K steps of one image per core
    scale = (args.height*args.width)/(512.*768.)
    sample = args.cpu_sample or max(4, int(min(cores, 32)*sample_per_core/max(1., scale)))
    which = 'ref' if coder.has_ref() else 'port'
    coder.build()
    # (one call in this process as well: the workers that do the coding are forked and torn down, and whoever lists the
    #  native libraries of the arm should see the coder's)
    assert coder.compress_lossless(numpy.array([0, 1, -2, 2, 1, 0, 0, 0], dtype=numpy.int16), numpy.full(3, 0.5), which)[2] == 20
    weights = wts.random_init(0, False)
    images = synthetic.synthetic_luma(numpy.random.default_rng(1), sample, args.height, args.width)
    tables = [load_tables(d) for d in args.deltas]
    with multiprocessing.get_context('fork').Pool(cores) as pool:
        for _ in range(warmup):
            cpu_pipeline(images[:4], weights, tables[0][0], tables[0][1], float(args.deltas[0]), cores, pool, which)
        seconds = 0.
        for k in range(steps):
            (table, map_mean) = tables[k % len(tables)]
            seconds += cpu_pipeline(images, weights, table, map_mean, float(args.deltas[k % len(tables)]), cores, pool, which)[0]
    value = sample*steps/seconds
    return {'value': value, 'unit': 'images/s', 'cores': cores,
            # "reference": the native part of the path (the C++ lossless coder, the dominant CPU cost) is the reference's
            # own code compiled from /root/reference into oracle/_ref; the transforms are a torch-CPU restatement because
            # TensorFlow cannot be installed offline (see `parts`). "port": the C restatement of the coder as well.
            'kind': 'reference' if which == 'ref' else 'port',
            'parts': {'transforms': 'torch-CPU fp32 restatement of the TF graph (oracle/transforms.py; TensorFlow unavailable offline)',
                      'coder': 'reference C++ from oracle/_ref' if which == 'ref' else 'C port (oracle/coder_oracle.c)',
                      'glue': 'numpy restatement of tools.py (oracle/glue.py)'},
            'sample': ('{} images of {}x{} per step x {} steps; coder = encode + in-call decode (compression.cpp:29-63), one '
                       'image per worker process').format(sample, args.height, args.width, steps),
            'ms_per_image': 1e3*seconds/(sample*steps)}


# ------------------------------------------------------------------------------------------------
# clocks

class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.gpu_index = gpu_index
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                                          '-i', str(gpu_index), '-lms', '20'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks = []
        reasons = set()
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(',')]
                    if len(parts) < 9:
                        continue
                    try:
                        clocks.append(float(parts[1]))
                        out['sm_max_mhz'] = float(parts[2])
                    except ValueError:
                        continue
                    for (name, val) in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                           parts[5:9]):
                        if val.lower().startswith('active'):
                            reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if clocks:
            # under load = the upper half of the samples (the sampler also sees idle gaps)
            clocks.sort()
            out['sm_mhz'] = float(numpy.median(clocks[len(clocks)//2:]))
            out['samples'] = len(clocks)
        out['reasons'] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# parity: one untimed batch against the CPU oracle (the oracle is the CHECKER here, never the thing measured)

def parity_block(args, device):
    """North star clauses on one batch of this workload, GPU path in the bench's arithmetic against the oracle:
    quantization indices (vs the float32 oracle pipeline, and both vs the float64 oracle: no float32 evaluation order can
    beat the float32 oracle's own count), distance of the mismatching coefficients to their bin edge, byte identity of
    every coded stream with the reference coder on the GPU's indices, PSNR and rate differences. Weights:
    weights.visible_init (reconstructions inside the BT.601 range; plain random-init weights decode to the clipping
    floor 16, which would make a PSNR comparison vacuous)."""
    import torch
    from autoencoder_based_image_compression_b200 import codec as native_codec
    from autoencoder_based_image_compression_b200 import synthetic
    from autoencoder_based_image_compression_b200 import weights as wts
    from oracle import coder as oracle_coder
    from oracle import glue as oracle_glue
    from oracle import transforms as T
    t0 = time.perf_counter()
    (h, w) = (args.height, args.width)
    n = max(1, min(args.batch, int(24*512*768/(h*w))))
    which = 'ref' if oracle_coder.has_ref() else 'port'
    weights = wts.visible_init(0, False)
    lum = synthetic.synthetic_luma(numpy.random.default_rng(1), n, h, w)
    codec = native_codec.Codec(weights, False, device=device, math=args.math)
    codec.set_coder_lanes(args.coder_lanes)
    y32 = numpy.concatenate([T.encoder(lum[i:i + 4, :, :, None].astype(numpy.float32), weights, False) for i in range(0, n, 4)])
    y64 = numpy.concatenate([T.encoder(lum[i:i + 4, :, :, None].astype(numpy.float64), weights, False, dtype=torch.float64)
                             for i in range(0, n, 4)])
    per_delta = []
    for mult in args.deltas:
        (table, map_mean) = load_tables(mult)
        delta = (mult*numpy.ones(128)).astype(numpy.float32)
        params = native_codec.CodingParams(delta, table, map_mean)
        (blob, stats) = codec.compress(lum, params, return_stats=True)
        idx = codec.last_indices(n, h, w).reshape(n*128, -1)
        rec = numpy.array(codec.decompress(blob, params), copy=True)
        (_, streams) = native_codec.parse_container(blob)
        identical = True
        for s in range(n*128):
            want = oracle_coder.encode_map(idx[s], table[s % 128], which)
            identical = identical and want[0] == 0 and (streams[s][0], streams[s][1]) == (want[2], want[4]) and \
                numpy.array_equal(streams[s][2], want[1]) and numpy.array_equal(streams[s][3], want[3])
        # the oracle pipeline (reconstructing_eae_kodak.py:144-224): its own indices, coder and decoder
        mean4 = map_mean.reshape((1, 1, 1, -1)).astype(numpy.float32)
        cq = oracle_glue.quantize_per_map(y32 - mean4, delta)
        k32 = oracle_glue.cast_float_to_int16(cq/delta.reshape((1, 1, 1, -1)))
        rec32 = numpy.concatenate([oracle_glue.cast_bt601(T.decoder(cq[i:i + 4] + mean4, weights, False))[..., 0]
                                   for i in range(0, n, 4)])
        bits32 = sum(oracle_coder.compress_lossless(k32[i, :, :, m].flatten(), table[m], which)[2]
                     for i in range(n) for m in range(128))
        t64 = (y64 - map_mean.astype(numpy.float64).reshape((1, 1, 1, -1)))/float(mult)
        k64 = numpy.rint(t64)
        edge = numpy.abs(numpy.abs(t64 - numpy.floor(t64)) - 0.5)
        k_gpu = idx.reshape(n, 128, -1).transpose(0, 2, 1).reshape(k64.shape)
        (bad_gpu, bad_32) = (k_gpu != k64, k32 != k64)
        psnr_gpu = [oracle_glue.psnr_2d(lum[i], rec[i]) for i in range(n)]
        psnr_ref = [oracle_glue.psnr_2d(lum[i], rec32[i]) for i in range(n)]
        per_delta.append({
            'bin_width': float(mult), 'coefficients': int(k64.size),
            'index_mismatches': int((k_gpu != k32).sum()),
            'index_mismatches_vs_fp64': int(bad_gpu.sum()), 'fp32_oracle_mismatches_vs_fp64': int(bad_32.sum()),
            'worst_edge_distance': float(edge[bad_gpu].max()) if bad_gpu.any() else 0.,
            'fp32_oracle_worst_edge_distance': float(edge[bad_32].max()) if bad_32.any() else 0.,
            'bitstream_identical_to_ref': bool(identical),
            'psnr_db': float(numpy.mean(psnr_gpu)), 'psnr_db_oracle': float(numpy.mean(psnr_ref)),
            'psnr_delta_db': float(numpy.max(numpy.abs(numpy.array(psnr_gpu) - numpy.array(psnr_ref)))),
            'rate_bpp': stats['total_bits']/float(n*h*w), 'rate_bpp_oracle': bits32/float(n*h*w),
            'rate_delta_rel': abs(stats['total_bits'] - bits32)/float(bits32),
            'pixels_off_by_one': float((rec != rec32).mean())})
    out = dict(per_delta[0])
    out.update({'images': n, 'math': args.math, 'weights': 'weights.visible_init(seed 0)', 'checker_coder':
                'oracle/_ref (reference C++)' if which == 'ref' else 'C port', 'seconds': time.perf_counter() - t0,
                'meets_north_star': bool(all(d['bitstream_identical_to_ref'] and d['index_mismatches'] <= 1e-4*d['coefficients'] and
                                             d['psnr_delta_db'] < 0.01 and d['rate_delta_rel'] < 1e-3 for d in per_delta))})
    if len(per_delta) > 1:
        out['per_bin_width'] = per_delta
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm

class Slot(object):
    """One pipeline slot: its own codec (weights + workspaces), CUDA stream and device / pinned buffers.
    Consecutive steps go to consecutive slots, so that the latency-bound lossless coder of one batch
    overlaps with the tensor-bound transforms of the next ones (classic multi-stream pipelining)."""

    def __init__(self, lib, native, native_codec, weights, math, device, n, h, w, bound, coder_lanes, d_acc):
        self.codec = native_codec.Codec(weights, False, device=device, math=math, own_stream=True)
        self.codec.set_coder_lanes(coder_lanes)
        native.check(lib.eae_codec_set_stats_accumulator(self.codec.handle, d_acc))
        self.stream = self.codec.stream
        self.d_container = lib.eae_device_alloc(bound)
        self.d_recon = lib.eae_device_alloc(n*h*w)
        self.d_total = lib.eae_device_alloc(8)
        self.d_stats = lib.eae_device_alloc(ctypes.sizeof(native.BatchStats))
        self.host_container = native.pinned_empty((bound,), numpy.uint8)
        self.host_recon = native.pinned_empty((n, h, w), numpy.uint8)
        self.ev_end = ctypes.c_void_p()
        native.check(lib.eae_event_create(ctypes.byref(self.ev_end)))


def measure_hist(lib, native, idx_planar, n, hw):
    """The warp-level histogram kernels (eae_histogram_streams_dev: extrema pass + counting pass, per image and map as
    rate_3d needs them) on the batch's own indices, CUDA events, median of 5 after 2 warm-up calls. Returns (ms, bins)."""
    nbytes = idx_planar.size*2
    d_idx = lib.eae_device_alloc(nbytes)
    native.check(lib.eae_memcpy_h2d(d_idx, native.ptr(idx_planar), nbytes, None))
    n_hist = n*128
    (d_mn, d_mx) = (lib.eae_device_alloc(4*n_hist), lib.eae_device_alloc(4*n_hist))
    cap = int(idx_planar.max()) - int(idx_planar.min()) + 1
    d_hist = lib.eae_device_alloc(8*n_hist*cap)
    (e0, e1) = (ctypes.c_void_p(), ctypes.c_void_p())
    native.check(lib.eae_event_create(ctypes.byref(e0)))
    native.check(lib.eae_event_create(ctypes.byref(e1)))
    times = []
    ms = ctypes.c_float(0.)
    for k in range(7):
        native.check(lib.eae_event_record(e0, None))
        native.check(lib.eae_histogram_streams_dev(d_idx, n, hw, 128, 1, d_mn, d_mx, None, None, 0, None))
        native.check(lib.eae_histogram_streams_dev(d_idx, n, hw, 128, 1, d_mn, None, None, d_hist, cap, None))
        native.check(lib.eae_event_record(e1, None))
        native.check(lib.eae_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        if k >= 2:
            times.append(ms.value)
    for p in (d_idx, d_mn, d_mx, d_hist):
        lib.eae_device_free(p)
    return (float(numpy.median(times)), cap)


def run_gpu_arm(args):
    import threading

    from autoencoder_based_image_compression_b200 import _native
    from autoencoder_based_image_compression_b200 import codec as native_codec
    from autoencoder_based_image_compression_b200 import parallel, synthetic
    from autoencoder_based_image_compression_b200 import weights as wts

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    lib = _native.lib()
    _native.require_gpu()
    torch = None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    _native.check(lib.eae_set_device(local_rank))
    if args.blocking_sync:
        _native.check(lib.eae_set_blocking_sync(1))

    (n, h, w) = (args.batch, args.height, args.width)
    steps = args.steps
    scaling = 'weak'
    if args.total:
        # a fixed job sharded over the ranks (parallel.shard_range): the step count follows from the workload
        (lo, hi) = parallel.shard_range(args.total, rank, world)
        steps = max(1, (hi - lo)//n)
        scaling = 'strong'
    weights = wts.random_init(0, False)
    all_params = []
    for mult in args.deltas:
        (table, map_mean) = load_tables(mult)
        all_params.append(native_codec.CodingParams((mult*numpy.ones(128)).astype(numpy.float32), table, map_mean))
    native_params = [p.native() for p in all_params]
    bound = int(lib.eae_container_bound(n, h, w, all_params[0].truncated_unary_length))
    depth = max(1, min(args.depth, steps))
    # running totals of the rate statistics of every step of every slot: ONE device accumulator per rank
    if world > 1:
        acc_t = torch.zeros(130, dtype=torch.int64, device='cuda')
        d_acc = acc_t.data_ptr()
    else:
        acc_t = None
        d_acc = lib.eae_device_alloc(ctypes.sizeof(_native.BatchStats))
    slots = [Slot(lib, _native, native_codec, weights, args.math, local_rank, n, h, w, bound, args.coder_lanes, d_acc)
             for _ in range(depth)]

    # Inputs rotate over enough distinct batches to exceed the L2 (each step also streams ~29 MB of fp32 activations
    # per 512 x 768 image through HBM), so no step finds its input cached from the previous one.
    batch_bytes = n*h*w
    rotate = max(2, -(-(L2_BYTES + batch_bytes)//batch_bytes))
    rotate = min(rotate, max(2, (2 << 30)//batch_bytes))          # (cap the pinned buffer at 2 GB)
    rng = numpy.random.default_rng(1000 + rank)
    host_images = _native.pinned_empty((rotate, n, h, w), numpy.uint8)
    base = synthetic.synthetic_luma(rng, n, h, w)
    for r in range(rotate):
        host_images[r] = numpy.roll(base, shift=(r, 7*r, 13*r), axis=(0, 1, 2))
    d_images = lib.eae_device_alloc(rotate*batch_bytes)
    _native.check(lib.eae_memcpy_h2d(d_images, _native.ptr(host_images), rotate*batch_bytes, None))
    _native.check(lib.eae_stream_synchronize(None))

    def step_dev(i, slot):
        img = d_images + (i % rotate)*batch_bytes
        prm = native_params[i % len(native_params)]
        _native.check(lib.eae_compress_dev(slot.codec.handle, ctypes.byref(prm), img, n, h, w,
                                           slot.d_container, bound, slot.d_total, slot.d_stats, slot.stream))
        _native.check(lib.eae_decompress_dev(slot.codec.handle, ctypes.byref(prm), slot.d_container, bound, n, h, w,
                                             slot.d_recon, slot.stream))

    def zero_acc():
        if world > 1:
            acc_t.zero_()
            torch.cuda.synchronize()
        else:
            zeros = numpy.zeros(ctypes.sizeof(_native.BatchStats), dtype=numpy.uint8)
            _native.check(lib.eae_memcpy_h2d(d_acc, _native.ptr(zeros), zeros.size, None))
            _native.check(lib.eae_stream_synchronize(None))

    def barrier():
        for slot in slots:
            _native.check(lib.eae_stream_synchronize(slot.stream))
        _native.check(lib.eae_stream_synchronize(None))
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    ev0 = ctypes.c_void_p()
    _native.check(lib.eae_event_create(ctypes.byref(ev0)))

    def timed_dev(active, nb_steps, first_index, reduce_stats):
        """`nb_steps` steps round-robin over the `active` slots; device time from the start event to the last slot's
        end event. With reduce_stats (N > 1) the rate statistics of the region are reduced over the ranks once, inside
        the timed region, after the last step of every slot."""
        zero_acc()
        barrier()
        _native.check(lib.eae_event_record(ev0, active[0].stream))
        for i in range(nb_steps):
            step_dev(first_index + i, active[i % len(active)])
        for slot in active:
            _native.check(lib.eae_event_record(slot.ev_end, slot.stream))
        worst = 0.
        ms = ctypes.c_float(0.)
        for slot in active:
            _native.check(lib.eae_event_elapsed_ms(ev0, slot.ev_end, ctypes.byref(ms)))      # (waits for the slot)
            worst = max(worst, ms.value)
        if reduce_stats and world > 1:
            # every slot of this rank has finished (the elapsed-time reads above waited for them): one NCCL all-reduce
            # of int64[130] over NVLink, timed on the device and added to the region
            (r0, r1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            r0.record()
            dist.all_reduce(acc_t)
            r1.record()
            r1.synchronize()
            worst += r0.elapsed_time(r1)
        barrier()
        return parallel.max_over_ranks(worst)

    # ---- device-resident timing ----
    # Warm-up: at least W steps and at least two per pipeline slot - a slot's first step launches kernel by kernel, its
    # second one captures and instantiates the step graphs (csrc/codec.cu, run_as_step_graph), from the third on it
    # replays them: the timed region must not contain the 2 x depth graph instantiations.
    warmup_run = max(args.warmup, 2*depth)
    # (the clock sampler starts before the warm-up: nvidia-smi needs ~0.1 s to come up and a 20-step region lasts 33 ms;
    #  its samples under load are those of the warm-up and of the timed region, which run back to back)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # (one step on one slot alone first: whatever the first launch of each kernel costs - module loading, attribute calls,
    #  the error-flag allocation - happens before twelve slots run beside each other)
    step_dev(0, slots[0])
    barrier()
    for i in range(warmup_run):
        step_dev(i, slots[i % depth])
    barrier()
    launches0 = lib.eae_launch_count()
    dev_ms = timed_dev(slots, steps, args.warmup, True)
    launches = lib.eae_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        job_bits = int(acc_t[128].item())          # (already summed over the ranks)
    else:
        acc = _native.BatchStats()
        _native.check(lib.eae_memcpy_d2h(ctypes.addressof(acc), d_acc, ctypes.sizeof(acc), None))
        _native.check(lib.eae_stream_synchronize(None))
        job_bits = int(acc.total_bits)
    # what the device-resident entry points recorded (they cannot return device-side failures themselves)
    for slot in slots:
        slot.codec.poll_status()

    # ---- serial pass of the same steps (one slot, per-kernel-class CUDA events): stage times, roofline ----
    serial_steps = min(steps, 96)
    lib.eae_profile_reset()
    lib.eae_profile_enable(1)
    serial_ms = timed_dev(slots[:1], serial_steps, args.warmup, False)
    lib.eae_profile_enable(0)
    profile = {}
    cls = 0
    while lib.eae_profile_name(cls):
        cnt = ctypes.c_uint64(0)
        tot = ctypes.c_double(0.)
        lib.eae_profile_read(cls, ctypes.byref(cnt), ctypes.byref(tot))
        profile[lib.eae_profile_name(cls).decode()] = (int(cnt.value), float(tot.value))
        cls += 1
    slots[0].codec.poll_status()

    # sanity of the timed work: the last step's container size and reconstruction
    last_slot = slots[0]
    total = numpy.zeros(1, dtype=numpy.uint64)
    _native.check(lib.eae_memcpy_d2h(_native.ptr(total), last_slot.d_total, 8, last_slot.stream))
    recon = numpy.empty((n, h, w), dtype=numpy.uint8)
    _native.check(lib.eae_memcpy_d2h(_native.ptr(recon), last_slot.d_recon, batch_bytes, last_slot.stream))
    _native.check(lib.eae_stream_synchronize(last_slot.stream))
    if total[0] <= 32 + 8*128*n or total[0] > bound or recon.min() < 16 or recon.max() > 235:
        raise RuntimeError('bench: the timed pipeline produced an implausible result')
    idx_last = slots[0].codec.last_indices(n, h, w)
    (hist_ms, hist_cap) = measure_hist(lib, _native, idx_last, n, (h//16)*(w//16))

    # ---- end-to-end through the public host API (pinned host buffers), one host thread per slot ----
    blob_bytes = [0]*depth

    def e2e_worker(k, first_index, nb_steps):
        slot = slots[k]
        for i in range(k, nb_steps, depth):
            params = all_params[(first_index + i) % len(all_params)]
            blob = slot.codec.compress(host_images[(first_index + i) % rotate], params, container=slot.host_container)
            slot.codec.decompress(blob, params, out=slot.host_recon)
            blob_bytes[k] += blob.size

    def timed_e2e(nb_steps, first_index):
        barrier()
        t0 = time.perf_counter()
        threads = [threading.Thread(target=e2e_worker, args=(k, first_index, nb_steps)) for k in range(depth)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        barrier()
        return parallel.max_over_ranks(1e3*(time.perf_counter() - t0))

    timed_e2e(max(args.warmup, depth), 0)
    # Three timed repetitions of the same K steps, the median reported (all three listed): the region is host-paced - 12
    # Python threads over 35-150 ms of wall clock - and about one run in fifteen lost 2-3x to a single hiccup of the box.
    e2e_runs = []
    for rep in range(3):
        blob_bytes = [0]*depth
        e2e_runs.append(timed_e2e(steps, args.warmup))
    e2e_ms = sorted(e2e_runs)[1]
    per_step_blob = sum(blob_bytes)//max(1, steps)
    h2d = batch_bytes + per_step_blob
    d2h = per_step_blob + batch_bytes + 8 + 8 + ctypes.sizeof(_native.BatchStats)

    images_total = n*world*steps
    config = base_config(args)
    config.update({'math': args.math, 'pipeline_depth': depth, 'coder_lanes': args.coder_lanes,
                   'pipelining': 'consecutive steps run on {} CUDA streams (one codec + workspace each), so the '
                                 'latency-bound coder of one batch overlaps the transforms of the next'.format(depth),
                   'l2': 'inputs rotate over {} distinct batches ({} MB; L2 = 126 MB); every step also streams about '
                         '{} MB of fp32 activations'.format(rotate, rotate*batch_bytes >> 20, int(29*n*h*w/(512.*768.))),
                   'parallelism': 'images sharded over {} GPU(s), no data-path collective; the int64[130] rate statistics '
                                  'accumulate on the device and are reduced over the ranks once per timed region (NCCL '
                                  'all-reduce over NVLink, inside the timed region)'.format(world),
                   'warmup_steps_run': warmup_run,
                   'steps_note': ('the step count follows from the fixed workload: {} images / ({} ranks x {} per step)'.format(
                       args.total, world, n)) if args.total else None})
    line = {
        'metric': METRIC,
        'value': images_total/(dev_ms/1e3),
        'unit': 'images/s',
        'n_gpus': world,
        'steps': steps,
        'warmup': args.warmup,
        'ms_per_step': dev_ms/steps,
        'higher_is_better': True,
        'scaling': scaling,
        'vs_baseline': None,
        'dtype': 'f32' if args.math == 'fp32' else args.math,
        'data': 'synthetic',
        'config': config,
        'mpixel_per_s': images_total*h*w/1e6/(dev_ms/1e3),
        'gpu_launches': int(launches),
        'e2e': {'value': images_total/(e2e_ms/1e3), 'unit': 'images/s',
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step_wall': e2e_ms/steps, 'host_threads': depth,
                'repetitions_images_per_s': [images_total/(t/1e3) for t in e2e_runs],
                'note': 'median of three timed repetitions of the same {} steps (wall clock, max over ranks)'.format(steps)},
        'serial': {'value': n*serial_steps/(serial_ms/1e3), 'ms_per_step': serial_ms/serial_steps,
                   'note': 'same steps on one stream, one batch at a time (batch latency)'},
        'latency_ms': serial_ms/serial_steps,
        'rate_bpp': float(total[0] - 32 - 8*128*n)*8./(n*h*w),
        'rate_bpp_job': job_bits/float(images_total*h*w),
        # (classes that never launched in the serial pass are left out: the histogram kernels are measured on their own below)
        'stage_ms_per_step': {k: v[1]/serial_steps for (k, v) in profile.items() if v[0] or k.startswith('gemm') or k.startswith('coder')},
    }
    if clocks is not None:
        line['clocks'] = {'sm_mhz': clocks['sm_mhz'], 'sm_max_mhz': clocks['sm_max_mhz'], 'reasons': clocks['reasons']}

    # ---- roofline of the dominant kernel: the tap-list GEMM (all four gemm_* classes are one kernel family) ----
    scale = (h*w)/(512.*768.)
    gemm_ms = sum(profile[k][1] for k in GFLOP_PER_IMAGE)
    gemm_launches = sum(profile[k][0] for k in GFLOP_PER_IMAGE)
    gflop = sum(GFLOP_PER_IMAGE.values())*scale*n*serial_steps
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    (bf16_peak, peak_src) = (1590., 'fallback 1.59 PFLOP/s bf16 (B200_PROFILING.md)')
    if os.path.isfile(peaks_path):
        with open(peaks_path) as f:
            bf16_peak = float(json.load(f)['bf16_tflops'])
        peak_src = 'MEASURED_PEAKS.json bf16_tflops (burst)'
    achieved = gflop/gemm_ms if gemm_ms > 0 else 0.     # GFLOP / ms = TFLOP/s
    # MMA work actually issued: tf32x3 = 3 MMAs per product (2 for layer 1, whose pixel operand is exact in TF32);
    # GDN / IGDN always use the 3-way split; fp32 = CUDA cores (no MMAs).
    passes = {'tf32': {'gemm_conv': 1., 'gemm_tconv': 1., 'gemm_gdn': 3., 'gemm_thin': 1.},
              'tf32x3': {'gemm_conv': 3., 'gemm_tconv': 3., 'gemm_gdn': 3., 'gemm_thin': 2.5},
              'mixed': {'gemm_conv': 3., 'gemm_tconv': 1., 'gemm_gdn': 2.05, 'gemm_thin': 1.5},   # fused IGDN5 / IGDN6 norms: one pass
              'fp32': {'gemm_conv': 0., 'gemm_tconv': 0., 'gemm_gdn': 0., 'gemm_thin': 0.}}[args.math]
    executed = sum(GFLOP_PER_IMAGE[k]*passes[k] for k in GFLOP_PER_IMAGE)*scale*n*serial_steps/gemm_ms if gemm_ms > 0 else 0.
    # DRAM bytes per launch of this kernel family: from the committed ncu capture of this workload (written next to the
    # per-layer table by scripts/ncu_gemm_layers.py), scaled to this batch
    traffic = None
    traffic_src = None
    for name in sorted(os.listdir(os.path.join(ROOT, 'profiles')), reverse=True):
        if name.endswith('_gemm_traffic.json'):
            with open(os.path.join(ROOT, 'profiles', name)) as f:
                rec = json.load(f)
            if rec.get('math') == args.math:
                traffic = rec['dram_bytes_per_step']/rec['launches_per_step']*(n/float(rec['images']))*scale
                traffic_src = 'profiles/' + name
                break
    # cuBLAS TF32 throughput measured on this pool's B200 (scripts/measure_tf32_peak.py), for information: the fraction is
    # taken against the (higher) bf16 / 2 figure derived from MEASURED_PEAKS.json.
    tf32_cublas = None
    tf32_path = os.path.join(ROOT, 'profiles', 'r01_tf32_peak.json')
    if os.path.isfile(tf32_path):
        with open(tf32_path) as f:
            tf32_cublas = json.load(f)
    line['roofline'] = {
        'bound': 'tensor', 'kernel': 'tap-list implicit GEMM (convs, transposed convs, GDN/IGDN), math=' + args.math,
        'achieved': achieved, 'peak': bf16_peak/2., 'unit': 'TFLOP/s',
        'frac': achieved/(bf16_peak/2.) if bf16_peak else None, 'traffic': traffic, 'traffic_source': traffic_src,
        'peak_source': peak_src + ' / 2: tcgen05 kind::tf32 runs at half the bf16 rate; fp32 data, so the TF32 '
                                  'tensor peak is the bound the north star names',
        'measured_in': 'the serial pass of the same steps inside this run (CUDA events around every launch)',
        'algorithmic_gflop_per_step': gflop/serial_steps, 'launches_per_step': gemm_launches/float(serial_steps),
        'avg_launch_ms': gemm_ms/gemm_launches if gemm_launches else None,
        'share_of_step': gemm_ms/serial_ms if serial_ms else None,
        'executed_mma_tflops': executed, 'frac_executed': executed/(bf16_peak/2.) if bf16_peak else None,
        'tf32_cublas_measured_tflops': {'burst': tf32_cublas['tf32_tflops'], 'sustained': tf32_cublas['tf32_tflops_sustained']}
                                       if tf32_cublas else None,
        'note': 'achieved / frac count ALGORITHMIC flops (SURVEY 8d). tf32x3 issues 3 TF32 MMAs per product (frac <= 1/3 '
                'by construction); mixed = 3 per product on the analysis side, which decides the indices, 1 on the synthesis '
                'side, whose bar is the PSNR (frac <= 1/2); frac_executed is the tensor-pipe view of the same time; traffic = '
                'dram read + write bytes per launch from the ncu capture under profiles/',
    }
    # ---- the HBM-bound kernels (north star: achieved GB/s of the quantize, histogram and coder kernels against the
    # measured peak) ----
    hbm_peak = 6552.
    if os.path.isfile(peaks_path):
        with open(peaks_path) as f:
            hbm_peak = float(json.load(f).get('hbm_gbs', hbm_peak))
    idx_bytes = n*(h//16)*(w//16)*128*2               # int16 indices of the batch
    stream_bytes = line['rate_bpp']*n*h*w/8.           # coded bytes of the batch
    per_step = {k: v[1]/serial_steps for (k, v) in profile.items()}
    per_step['hist'] = hist_ms
    per_step['coder_encode_arith'] = per_step['coder_encode'] - per_step.get('binarize', 0.)
    algorithmic = {'quantize': 3*idx_bytes,            # un-fused path only: fp32 latent read, int16 planar indices written
                   'dequantize': 3*idx_bytes,
                   'binarize': idx_bytes + stream_bytes,        # indices read; truncated-unary string + bypass stream written
                   'hist': idx_bytes,                           # indices read (second pass from L2), counters in shared memory
                   'coder_encode_arith': 2*stream_bytes,
                   'coder_decode': idx_bytes + stream_bytes,
                   'pack': 2*stream_bytes}
    line['hbm_kernels'] = {
        k: {'algorithmic_bytes_per_step': float(b), 'ms_per_step': per_step[k],
            'achieved_gbs': float(b)/per_step[k]/1e6 if per_step.get(k, 0.) > 0 else None,
            'frac_of_hbm_peak': float(b)/per_step[k]/1e6/hbm_peak if per_step.get(k, 0.) > 0 else None}
        for (k, b) in algorithmic.items() if k in per_step}
    line['hbm_kernels']['note'] = ('serial pass, CUDA events per kernel class (hist: its own measurement on the last batch\'s '
                                   'indices, {} bins per histogram); peak {:.0f} GB/s (MEASURED_PEAKS.json). quantize / dequantize '
                                   'are 0 where they are fused into the transforms (the default). The arithmetic coder kernels are '
                                   'latency-bound (one thread per coded stream, a dependent chain per bin), not bandwidth-bound: '
                                   'their fraction is reported, not optimised for').format(hist_cap, hbm_peak)
    if rank == 0 and world == 1 and not args.no_parity:
        for slot in slots[1:]:            # (free the workspaces of the pipeline before the parity batch)
            slot.codec.close()
        line['parity'] = parity_block(args, local_rank)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = run_cpu_arm(args, steps=1, warmup=1, sample_per_core=16)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    if args.impl == 'reference':
        if rank != 0:
            return
        base = run_cpu_arm(args, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        config = base_config(args)
        # (the same keys as the GPU arm's line, with this arm's values)
        config.update({'math': 'fp32 on the host cores', 'pipeline_depth': 0, 'coder_lanes': None,
                       'pipelining': 'none: transforms in mini-batches of 4 on all cores, then the coder one image per worker process',
                       'l2': 'not applicable (host arm)', 'parallelism': '{} host cores of one box'.format(base['cores']),
                       'warmup_steps_run': min(args.warmup, 1),
                       'steps_note': 'every step is a bounded sample of the workload: ' + base['sample']})
        line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'images/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': base['ms_per_image'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config,
                'cpu_baseline': base, 'gpu_launches': 0,
                'e2e': {'value': base['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return
    run_gpu_arm(args)


if __name__ == '__main__':
    main()

"""Headline benchmark: images/s of the full EAE codec hot path (encode -> quantize -> lossless code ->
bitstream -> decode) on synthetic 512 x 768 luminance images.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Under torchrun (N > 1) every rank drives one GPU with its own batch (weak scaling, no data-path
collective; one NCCL all-reduce of the rate statistics per step). Rank 0 prints ONE JSON line.

A "step" is one pass of the hot path over one batch (BASELINE.json configs[1]: 24 images per GPU):
  value  images/s with the batch already resident in HBM (eae_compress_dev + eae_decompress_dev)
  e2e    the same through the public host API (Codec.compress / Codec.decompress) from pinned host
         memory, host<->device copies inside the timed region
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=96)
    ap.add_argument('--warmup', type=int, default=16)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=24, help='images per GPU per step')
    ap.add_argument('--height', type=int, default=512)
    ap.add_argument('--width', type=int, default=768)
    ap.add_argument('--math', default=os.environ.get('EAE_MATH', 'mixed'), choices=['fp32', 'tf32x3', 'tf32', 'mixed'])
    ap.add_argument('--cpu-sample', type=int, default=0, help='images in the CPU sample (0 = one per core)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--blocking-sync', type=int, default=1,
                    help='1: host threads sleep while they wait for the GPU (one thread per pipeline slot and one process '
                         'per GPU would otherwise spin on more threads than the box has cores); 0: driver default')
    ap.add_argument('--coder-lanes', type=int, default=1,
                    help='GPU threads per coded stream (0 = one warp per stream: lowest latency)')
    ap.add_argument('--depth', type=int, default=int(os.environ.get('EAE_PIPELINE_DEPTH', '12')),
                    help='pipeline slots (CUDA streams) that consecutive steps rotate over')
    return ap.parse_args()


def load_tables():
    with numpy.load(os.path.join(ROOT, 'tests', 'golden', 'tables.npz')) as data:
        return (numpy.ascontiguousarray(data['1_10000__binary_probabilities_1']),
                numpy.ascontiguousarray(data['1_10000__map_mean']))


def workload_name(args):
    return ('configs[1]: batch of {} synthetic {}x{} luma images per GPU, fixed-delta Kodak EAE (6 GDN/IGDN, '
            'random-init weights seed 0), bin width 1.0, shipped table 1_10000/binary_probabilities_1, '
            'map_mean 1_10000').format(args.batch, args.height, args.width)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle / reference on the host cores

def _cpu_code_one(job):
    from oracle import coder
    (planar, table, which) = job
    (err, out, bits) = coder.compress_maps_planar(planar, table, which=which)
    assert err == 0 and numpy.array_equal(out, planar)
    return int(bits.sum())


def cpu_pipeline(images, weights, table, map_mean, cores, pool, which):
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
    (planar, quantized) = ([], [])
    for i0 in range(0, images.shape[0], 4):          # reconstructing_eae_kodak.py:624 batch_size = 4
        x = images[i0:i0 + 4, :, :, None].astype(numpy.float32)
        y = transforms.encoder(x, weights, False)
        cq = glue.quantize_per_map(y - mean, numpy.ones(128, dtype=numpy.float32))
        idx = glue.cast_float_to_int16(cq)
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
    sample = args.cpu_sample or max(4, min(cores, 32)*sample_per_core)
    which = 'ref' if coder.has_ref() else 'port'
    coder.build()
    (table, map_mean) = load_tables()
    weights = wts.random_init(0, False)
    images = synthetic.synthetic_luma(numpy.random.default_rng(1), sample, args.height, args.width)
    with multiprocessing.get_context('fork').Pool(cores) as pool:
        for _ in range(warmup):
            cpu_pipeline(images[:4], weights, table, map_mean, cores, pool, which)
        seconds = 0.
        for _ in range(steps):
            seconds += cpu_pipeline(images, weights, table, map_mean, cores, pool, which)[0]
    value = sample*steps/seconds
    return {'value': value, 'unit': 'images/s', 'cores': cores,
            'kind': 'port',
            'sample': ('{} images of {}x{} per step x {} steps; transforms = torch-CPU fp32 restatement of the TF '
                       'graph (TensorFlow unavailable offline), coder = {} (encode + in-call decode), one image per '
                       'worker process').format(sample, args.height, args.width, steps,
                                                "reference C++ from oracle/_ref" if which == 'ref' else 'C port'),
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
                                          '-i', str(gpu_index), '-lms', '100'],
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
# GPU arm

class Slot(object):
    """One pipeline slot: its own codec (weights + workspaces), CUDA stream and device / pinned buffers.
    Consecutive steps go to consecutive slots, so that the latency-bound lossless coder of one batch
    overlaps with the tensor-bound transforms of the next ones (classic multi-stream pipelining)."""

    def __init__(self, lib, native, native_codec, weights, math, device, n, h, w, bound, world, coder_lanes):
        self.codec = native_codec.Codec(weights, False, device=device, math=math, own_stream=True)
        self.codec.set_coder_lanes(coder_lanes)
        self.stream = self.codec.stream
        self.d_container = lib.eae_device_alloc(bound)
        self.d_recon = lib.eae_device_alloc(n*h*w)
        self.d_total = lib.eae_device_alloc(8)
        self.stats_t = None
        if world > 1:
            import torch
            self.stats_t = torch.zeros(130, dtype=torch.int64, device='cuda')
            self.torch_stream = torch.cuda.ExternalStream(self.stream.value)
            self.d_stats = self.stats_t.data_ptr()
        else:
            self.d_stats = lib.eae_device_alloc(ctypes.sizeof(native.BatchStats))
        self.host_container = native.pinned_empty((bound,), numpy.uint8)
        self.host_recon = native.pinned_empty((n, h, w), numpy.uint8)
        self.ev_end = ctypes.c_void_p()
        native.check(lib.eae_event_create(ctypes.byref(self.ev_end)))


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
    (table, map_mean) = load_tables()
    weights = wts.random_init(0, False)
    params = native_codec.CodingParams(numpy.ones(128, dtype=numpy.float32), table, map_mean)
    native_params = params.native()
    bound = int(lib.eae_container_bound(n, h, w, params.truncated_unary_length))
    depth = max(1, args.depth)
    slots = [Slot(lib, _native, native_codec, weights, args.math, local_rank, n, h, w, bound, world, args.coder_lanes)
             for _ in range(depth)]

    # Inputs rotate over enough distinct batches to exceed the L2 (each step also streams ~0.7 GB of
    # fp32 activations through HBM), so no step finds its input cached from the previous one.
    batch_bytes = n*h*w
    rotate = max(2, -(-(L2_BYTES + batch_bytes)//batch_bytes))
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
        _native.check(lib.eae_compress_dev(slot.codec.handle, ctypes.byref(native_params), img, n, h, w,
                                           slot.d_container, bound, slot.d_total, slot.d_stats, slot.stream))
        if world > 1:
            with torch.cuda.stream(slot.torch_stream):
                dist.all_reduce(slot.stats_t)     # per-map bit totals over all ranks (NCCL over NVLink)
        _native.check(lib.eae_decompress_dev(slot.codec.handle, ctypes.byref(native_params), slot.d_container, bound, n, h, w,
                                             slot.d_recon, slot.stream))

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

    def timed_dev(active, steps, first_index):
        """`steps` steps round-robin over the `active` slots; device time from the start event to the last
        slot's end event."""
        barrier()
        _native.check(lib.eae_event_record(ev0, active[0].stream))
        for i in range(steps):
            step_dev(first_index + i, active[i % len(active)])
        for slot in active:
            _native.check(lib.eae_event_record(slot.ev_end, slot.stream))
        worst = 0.
        ms = ctypes.c_float(0.)
        for slot in active:
            _native.check(lib.eae_event_elapsed_ms(ev0, slot.ev_end, ctypes.byref(ms)))
            worst = max(worst, ms.value)
        barrier()
        return parallel.max_over_ranks(worst)

    # ---- device-resident timing ----
    for i in range(max(args.warmup, depth)):
        step_dev(i, slots[i % depth])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.eae_launch_count()
    dev_ms = timed_dev(slots, args.steps, args.warmup)
    launches = lib.eae_launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    # ---- serial pass of the same steps (one slot, per-kernel-class CUDA events): stage times, roofline ----
    lib.eae_profile_reset()
    lib.eae_profile_enable(1)
    serial_ms = timed_dev(slots[:1], args.steps, args.warmup)
    lib.eae_profile_enable(0)
    profile = {}
    for cls in range(11):
        cnt = ctypes.c_uint64(0)
        tot = ctypes.c_double(0.)
        lib.eae_profile_read(cls, ctypes.byref(cnt), ctypes.byref(tot))
        profile[lib.eae_profile_name(cls).decode()] = (int(cnt.value), float(tot.value))

    # sanity of the timed work: the last step's container size and reconstruction
    last_slot = slots[0]
    total = numpy.zeros(1, dtype=numpy.uint64)
    _native.check(lib.eae_memcpy_d2h(_native.ptr(total), last_slot.d_total, 8, last_slot.stream))
    recon = numpy.empty((n, h, w), dtype=numpy.uint8)
    _native.check(lib.eae_memcpy_d2h(_native.ptr(recon), last_slot.d_recon, batch_bytes, last_slot.stream))
    _native.check(lib.eae_stream_synchronize(last_slot.stream))
    last = host_images[(args.warmup + args.steps - 1) % rotate]
    sse = float(((recon.astype(numpy.int64) - last.astype(numpy.int64))**2).sum())
    if total[0] <= 32 + 8*128*n or recon.min() < 16 or recon.max() > 235:
        raise RuntimeError('bench: the timed pipeline produced an implausible result')

    # ---- end-to-end through the public host API (pinned host buffers), one host thread per slot ----
    blob_bytes = [0]*depth

    def e2e_worker(k, first_index, steps):
        slot = slots[k]
        for i in range(k, steps, depth):
            blob = slot.codec.compress(host_images[(first_index + i) % rotate], params, container=slot.host_container)
            slot.codec.decompress(blob, params, out=slot.host_recon)
            blob_bytes[k] += blob.size

    def timed_e2e(steps, first_index):
        barrier()
        t0 = time.perf_counter()
        threads = [threading.Thread(target=e2e_worker, args=(k, first_index, steps)) for k in range(depth)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        barrier()
        return parallel.max_over_ranks(1e3*(time.perf_counter() - t0))

    timed_e2e(max(args.warmup, depth), 0)
    blob_bytes = [0]*depth
    e2e_ms = timed_e2e(args.steps, args.warmup)
    per_step_blob = sum(blob_bytes)//max(1, args.steps)
    h2d = batch_bytes + per_step_blob
    d2h = per_step_blob + batch_bytes + 8 + 8 + ctypes.sizeof(_native.BatchStats)

    images_total = n*world*args.steps
    line = {
        'metric': METRIC,
        'value': images_total/(dev_ms/1e3),
        'unit': 'images/s',
        'n_gpus': world,
        'steps': args.steps,
        'warmup': args.warmup,
        'ms_per_step': dev_ms/args.steps,
        'higher_is_better': True,
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f32' if args.math == 'fp32' else args.math,
        'data': 'synthetic',
        'config': {'workload': workload_name(args), 'batch_per_gpu': n, 'math': args.math,
                   'pipeline_depth': depth, 'coder_lanes': args.coder_lanes,
                   'pipelining': 'consecutive steps run on {} CUDA streams (one codec + workspace each), so the '
                                 'latency-bound coder of one batch overlaps the transforms of the next'.format(depth),
                   'l2': 'inputs rotate over {} distinct batches ({} MB > 126 MB L2); every step also streams about '
                         '{} MB of fp32 activations'.format(rotate, rotate*batch_bytes >> 20, 29*n),
                   'parallelism': 'images sharded over {} GPU(s), no data-path collective, one NCCL all-reduce of '
                                  'int64[130] rate statistics per step'.format(world)},
        'mpixel_per_s': images_total*h*w/1e6/(dev_ms/1e3),
        'gpu_launches': int(launches),
        'e2e': {'value': images_total/(e2e_ms/1e3), 'unit': 'images/s',
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step_wall': e2e_ms/args.steps, 'host_threads': depth},
        'serial': {'value': images_total/(serial_ms/1e3), 'ms_per_step': serial_ms/args.steps,
                   'note': 'same steps on one stream, one batch at a time (batch latency)'},
        'rate_bpp': float(total[0] - 32 - 8*128*n)*8./(n*h*w),
        'psnr_db_random_weights': float(10.*numpy.log10(255.**2/(sse/(n*h*w)))) if sse > 0 else None,
        'stage_ms_per_step': {k: v[1]/args.steps for (k, v) in profile.items()},
    }
    if clocks is not None:
        line['clocks'] = {'sm_mhz': clocks['sm_mhz'], 'sm_max_mhz': clocks['sm_max_mhz'], 'reasons': clocks['reasons']}

    # ---- roofline of the dominant kernel: the tap-list GEMM (all four gemm_* classes are one kernel) ----
    scale = (h*w)/(512.*768.)
    gemm_ms = sum(profile[k][1] for k in GFLOP_PER_IMAGE)
    gemm_launches = sum(profile[k][0] for k in GFLOP_PER_IMAGE)
    gflop = sum(GFLOP_PER_IMAGE.values())*scale*n*args.steps
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
    executed = sum(GFLOP_PER_IMAGE[k]*passes[k] for k in GFLOP_PER_IMAGE)*scale*n*args.steps/gemm_ms if gemm_ms > 0 else 0.
    # DRAM bytes per launch of this kernel from the committed ncu captures of this workload
    # (profiles/r01_ncu_full_gemm_layers_v13.md: 13 launches per step, 1573 MB per 24-image step in the mixed mode;
    #  1593 MB with 3xTF32 everywhere).
    traffic = {'mixed': 1573.3e6, 'fp32': None}.get(args.math, 1593.3e6)
    traffic = traffic/13.*(n/24.)*scale if traffic else None
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
        'frac': achieved/(bf16_peak/2.) if bf16_peak else None, 'traffic': traffic,
        'peak_source': peak_src + ' / 2: tcgen05 kind::tf32 runs at half the bf16 rate; fp32 data, so the TF32 '
                                  'tensor peak is the bound the north star names',
        'measured_in': 'the serial pass of the same steps inside this run (CUDA events around every launch)',
        'algorithmic_gflop_per_step': gflop/args.steps, 'launches_per_step': gemm_launches/args.steps,
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
    # ---- the HBM-bound kernels (north star: achieved GB/s of the quantize and coder kernels against the measured peak) ----
    hbm_peak = 6552.
    if os.path.isfile(peaks_path):
        with open(peaks_path) as f:
            hbm_peak = float(json.load(f).get('hbm_gbs', hbm_peak))
    idx_bytes = n*(h//16)*(w//16)*128*2               # int16 indices of the batch
    stream_bytes = line['rate_bpp']*n*h*w/8.           # coded bytes of the batch
    algorithmic = {'quantize': 3*idx_bytes,            # fp32 latent read, int16 planar indices written
                   'dequantize': 3*idx_bytes,
                   'coder_encode': idx_bytes + stream_bytes,
                   'coder_decode': idx_bytes + stream_bytes,
                   'pack': 2*stream_bytes}
    line['hbm_kernels'] = {
        k: {'algorithmic_bytes_per_step': float(b), 'ms_per_step': profile[k][1]/args.steps,
            'achieved_gbs': float(b)/(profile[k][1]/args.steps)/1e6 if profile[k][1] > 0 else None,
            'frac_of_hbm_peak': float(b)/(profile[k][1]/args.steps)/1e6/hbm_peak if profile[k][1] > 0 else None}
        for (k, b) in algorithmic.items()}
    line['hbm_kernels']['note'] = ('serial pass, CUDA events per kernel class; peak {:.0f} GB/s (MEASURED_PEAKS.json). The coder '
                                   'kernels are latency-bound (one thread per coded stream, a dependent chain per bin), not '
                                   'bandwidth-bound: their fraction is reported, not optimised for').format(hbm_peak)
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
        line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'images/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': base['ms_per_image'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': workload_name(args), 'batch_per_gpu': args.batch},
                'cpu_baseline': base, 'gpu_launches': 0,
                'e2e': {'value': base['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return
    run_gpu_arm(args)


if __name__ == '__main__':
    main()

"""Multi-GPU: one process per GPU, images sharded across ranks, one small reduction of rate statistics.

Every image is coded independently of every other (kodak_tensorflow/lossless/compression.py:67-81,
reconstructing_eae_kodak.py:212-224), so the data path has NO collective: rank r compresses and
reconstructs the contiguous slice ``shard_range(n, r, world)`` of the batch. The only exchange is the
sum over ranks of the per-map bit totals, the per-image PSNR sum and the image count (about 1 KB), which is what
``numpy.mean(rate, axis=1)`` / ``numpy.mean(psnr, axis=1)`` reduce in the reference
(reconstructing_eae_kodak.py:810-815): the MEAN OF THE PER-IMAGE PSNRs, not the PSNR of the pooled error. Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy


def shard_range(nb_items, rank, world_size):
    """Contiguous, balanced slice [start, stop) of ``nb_items`` for ``rank`` (first ranks get the remainder)."""
    if world_size < 1 or rank < 0 or rank >= world_size:
        raise ValueError('bad rank / world_size')
    base = nb_items//world_size
    extra = nb_items % world_size
    start = rank*base + min(rank, extra)
    return (start, start + base + (1 if rank < extra else 0))


def pack_stats(bits_per_map, total_bits, nb_dead_maps, sum_squared_error, nb_pixels, nb_images, sum_psnr=0.):
    """Statistics of one rank as a float64 vector of 128 + 6 entries (the counts are exactly representable: below
    2**53). ``sum_psnr``: sum over this rank's images of their PSNR (tools.psnr_2d, tools.py:831-881)."""
    vec = numpy.zeros(128 + 6, dtype=numpy.float64)
    vec[:128] = numpy.asarray(bits_per_map, dtype=numpy.float64)
    vec[128:] = (total_bits, nb_dead_maps, sum_squared_error, nb_pixels, nb_images, sum_psnr)
    return vec


def unpack_stats(vec):
    vec = numpy.asarray(vec, dtype=numpy.float64)
    return {'bits_per_map': vec[:128].copy(), 'total_bits': vec[128], 'nb_dead_maps': vec[129],
            'sum_squared_error': vec[130], 'nb_pixels': vec[131], 'nb_images': vec[132], 'sum_psnr': vec[133]}


def summarize(stats):
    """Mean rate (bpp) and mean PSNR over all images of all ranks, as the reference reports them
    (reconstructing_eae_kodak.py:810-815: numpy.mean over the images of the per-image values). ``psnr_db_pooled`` is
    the PSNR of the mean squared error pooled over all pixels - a different number (Jensen), kept under its own name."""
    out = dict(stats)
    out['rate_bpp'] = stats['total_bits']/stats['nb_pixels'] if stats['nb_pixels'] else float('nan')
    out['psnr_db'] = stats['sum_psnr']/stats['nb_images'] if stats['nb_images'] else float('nan')
    mse = stats['sum_squared_error']/stats['nb_pixels'] if stats['nb_pixels'] else float('nan')
    out['psnr_db_pooled'] = 10.*numpy.log10(255.**2/mse) if mse and mse > 0. else float('inf')
    return out


def all_reduce_stats(vec, device=None):
    """Sum of ``vec`` over all ranks of the default process group (no-op without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return numpy.asarray(vec, dtype=numpy.float64)
    t = torch.from_numpy(numpy.ascontiguousarray(vec, dtype=numpy.float64))
    if dist.get_backend() == 'nccl':
        t = t.to(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(value, device=None):
    """Max of a scalar over ranks (device-timed durations are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == 'nccl':
        t = t.to(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

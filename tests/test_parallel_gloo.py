"""world_size-2 gloo run of the multi-rank host logic: shard the batch, reduce the statistics."""
import os
import socket

import numpy
import torch.distributed as dist
import torch.multiprocessing as mp

from autoencoder_based_image_compression_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nb_images, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    (a, b) = parallel.shard_range(nb_images, rank, world)
    # Per-image "bits" that every rank can recompute: image i contributes (i + 1) * (map + 1).
    bits = numpy.zeros(128)
    sse = 0.
    for i in range(a, b):
        bits += (i + 1)*numpy.arange(1, 129)
        sse += 10.*(i + 1)
    vec = parallel.pack_stats(bits, bits.sum(), rank, sse, (b - a)*64, b - a, sum(30. + i for i in range(a, b)))
    total = parallel.all_reduce_stats(vec)
    slowest = parallel.max_over_ranks(1.5 + rank)
    numpy.save(os.path.join(out_dir, 'rank{}.npy'.format(rank)), numpy.append(total, slowest))
    dist.destroy_process_group()


def test_two_rank_shard_and_reduce(tmp_path):
    (world, nb_images) = (2, 7)
    mp.spawn(_worker, args=(world, _free_port(), nb_images, str(tmp_path)), nprocs=world, join=True)
    results = [numpy.load(str(tmp_path/'rank{}.npy'.format(r))) for r in range(world)]
    assert numpy.array_equal(results[0], results[1])
    stats = parallel.unpack_stats(results[0][:-1])
    tri = nb_images*(nb_images + 1)/2
    assert numpy.array_equal(stats['bits_per_map'], tri*numpy.arange(1, 129))
    assert stats['nb_images'] == nb_images and stats['nb_pixels'] == nb_images*64
    assert stats['sum_squared_error'] == 10.*tri and stats['nb_dead_maps'] == 1
    # mean of the per-image PSNRs over both ranks, as numpy.mean(psnr, axis=1) gives it in the reference
    assert parallel.summarize(stats)['psnr_db'] == sum(30. + i for i in range(nb_images))/nb_images
    assert results[0][-1] == 2.5

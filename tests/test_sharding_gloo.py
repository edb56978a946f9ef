"""N > 1 host logic on CPU: two gloo ranks shard a job by image and reduce their statistics.  (The encode
itself needs a B200; here each rank 'encodes' its shard with the CPU oracle so that the sharded job can be
checked against the unsharded one byte for byte.)"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "jpeg-xl-lossy-image-compression-thesis_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_images, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG)
    import oracle_lib
    ora = oracle_lib.load(rebuild=False)
    local = pkg.ShardStats()
    sizes = {}
    for i in pkg.shard_indices(num_images, rank, world):
        img = pkg.synth_image(96, 64, i)
        cs = ora.encode(img, pkg.distance_for_image(i), 7, pkg.PROPOSAL_COMBINED, 0).dump("codestream")
        sizes[i] = int(cs.size)
        local.images += 1; local.pixels += 96 * 64; local.codestream_bytes += int(cs.size)
        local.kernel_launches += 30
    local.device_ms = 10.0 * (rank + 1)
    job = pkg.gather_stats(local)
    out[rank] = (job.images, job.pixels, job.codestream_bytes, job.max_ms, job.kernel_launches, sizes)
    dist.destroy_process_group()


def test_shard_indices_round_robin(pkg):
    assert pkg.shard_indices(10, 0, 4) == [0, 4, 8] and pkg.shard_indices(10, 3, 4) == [3, 7]
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in pkg.shard_indices(37, r, world))
        assert got == list(range(37))                         # every image exactly once
    assert pkg.shard_indices(3, 7, 8) == []                   # more ranks than images: empty shard
    with pytest.raises(ValueError):
        pkg.shard_indices(4, 2, 2)
    assert [pkg.distance_for_image(i) for i in range(8)] == [0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 0.5, 1.0]


def test_gather_stats_without_group(pkg):
    j = pkg.gather_stats(pkg.ShardStats(2, 1000, 300, 5.0, 60))
    assert (j.images, j.pixels, j.codestream_bytes, j.max_ms, j.kernel_launches) == (2, 1000, 300, 5.0, 60)
    assert j.bpp == pytest.approx(2.4) and j.mp_per_s == pytest.approx(0.2)


def test_two_gloo_ranks_match_single_rank(pkg, oracle):
    num_images, world = 5, 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, num_images, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][:5] == res[1][:5]                            # every rank sees the same job totals
    images, pixels, total_bytes, max_ms, launches, _ = res[0]
    assert images == num_images and pixels == num_images * 96 * 64 and max_ms == 20.0 and launches == 30 * num_images
    sizes = {}
    for r in range(world):
        sizes.update(res[r][5])
    assert sorted(sizes) == list(range(num_images))
    # the sharded job produces exactly the bytes of the unsharded one
    for i in range(num_images):
        cs = oracle.encode(pkg.synth_image(96, 64, i), pkg.distance_for_image(i), 7, pkg.PROPOSAL_COMBINED, 0).dump("codestream")
        assert cs.size == sizes[i]
    assert total_bytes == sum(sizes.values())


def test_numa_binding_is_best_effort():
    """bind_to_gpu_numa_node never raises: without a GPU / sysfs entry it reports that nothing was bound."""
    import importlib
    pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
    info = pkg.bind_to_gpu_numa_node(0)
    assert set(info) >= {"numa_node", "cpus"}

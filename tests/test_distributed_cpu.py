"""Host-side logic of the multi-GPU path (misaki_render_b200/distributed.py) on CPU: sample-range sharding and
the film reduction over a world_size-2 gloo group.  Each rank renders ITS sample range with the CPU oracle (the
checker stands in for the GPU here -- there is none in this container) and the reduced film must equal the
oracle's single-process render of the whole job."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from misaki_render_b200 import capi, distributed as msk_dist  # noqa: E402


def test_shard_samples_tiles_the_range():
    for spp in (0, 1, 5, 16, 64, 4096):
        for world in (1, 2, 3, 4, 8):
            ranges = [msk_dist.shard_samples(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
                assert e0 == b1 and b0 <= e0
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        msk_dist.shard_samples(16, 2, 2)


def test_shard_desc_keeps_the_job():
    rd = capi.render_desc(spp=64, max_depth=7, rr_depth=3, base_seed=11, sample_begin=8, sample_end=40)
    parts = [msk_dist.shard_desc(rd, r, 4) for r in range(4)]
    assert [(p.sample_begin, p.sample_end) for p in parts] == [(8, 16), (16, 24), (24, 32), (32, 40)]
    for p in parts:
        assert (p.spp, p.max_depth, p.rr_depth, p.base_seed, p.clear_film) == (64, 7, 3, 11, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import pyoracle
    from workloads import scenes
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd = scenes.cbox(32, 32)
    rd = capi.render_desc(spp=6, max_depth=4)
    osc = pyoracle.OracleScene(sd)
    mine = msk_dist.shard_desc(rd, rank, world)
    film, st = osc.render(mine, nthreads=1)
    t = torch.from_numpy(film)
    msk_dist.reduce_film(t, 0)
    paths = torch.tensor([int(st.paths)], dtype=torch.int64)
    dist.all_reduce(paths)
    if rank == 0:
        whole, st_all = osc.render(rd, nthreads=1)
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
        np.save(os.path.join(out_dir, "whole.npy"), whole)
        np.save(os.path.join(out_dir, "paths.npy"), np.array([int(paths.item()), int(st_all.paths)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_film_reduce_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced, whole = np.load(tmp_path / "reduced.npy"), np.load(tmp_path / "whole.npy")
    paths = np.load(tmp_path / "paths.npy")
    assert paths[0] == paths[1] == 32 * 32 * 6
    # same samples, same seeds; only the order of the float additions into the film differs
    np.testing.assert_allclose(reduced, whole, rtol=2e-5, atol=1e-6)
    assert np.abs(whole[..., 4]).sum() > 0

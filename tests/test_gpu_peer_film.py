"""msk_gpu_reduce_film (csrc/msk_peer.cu): the multi-GPU film reduction over CUDA-IPC peer memory, SURVEY 8e.

Two worker processes, one per rank, as bench.py --gpus N runs them.  On a box with one GPU both ranks use device 0
(CUDA IPC and the device-side flag protocol work the same between two processes on one device; the kernels of the
two contexts are time-sliced), with two or more GPUs they use devices 0 and 1 and the pull goes through NVLink."""
import multiprocessing as mp
import os
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W = H = 48
SPP = 8


def _worker(rank, world, ndev, to_root, from_root, result):
    try:
        os.environ["MSK_PEER_TIMEOUT_S"] = "20"
        import torch
        from misaki_render_b200 import capi, distributed as msk_dist
        from workloads import scenes

        dev_index = rank % ndev
        torch.cuda.set_device(dev_index)
        dev = torch.device("cuda", dev_index)

        def exchange(handle):  # star through the parent-created pipes: rank r -> root, root -> everyone
            if rank == 0:
                handles = [handle] + [to_root[r].recv() for r in range(1, world)]
                for r in range(1, world):
                    from_root[r].send(handles)
                return handles
            to_root[rank].send(handle)
            return from_root[rank].recv()

        out = {}
        with capi.Context(dev_index) as ctx:
            peer = msk_dist.PeerFilm(ctx, (H, W, 5), rank, world, exchange=exchange)
            film = peer.tensor(dev)
            ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
            # (1) protocol: three epochs of rank-specific constants; the root must see the exact sum every time
            sums = []
            for epoch in range(1, 4):
                with torch.cuda.stream(ext):
                    film.fill_(float(rank + 1) * epoch)
                    peer.reduce()
                    if rank == 0:
                        sums.append(film.clone())
                peer.check()
            if rank == 0:
                want = [sum(float(r + 1) * e for r in range(world)) for e in range(1, 4)]
                out["sums_ok"] = all(bool((s == w).all().item()) for s, w in zip(sums, want))
            # (2) the real thing: each rank renders its sample range, the root's reduced film equals the whole job
            sd = scenes.cbox(W, H)
            rd = capi.render_desc(spp=SPP, max_depth=4)
            with capi.Scene(ctx, sd) as sc:
                with torch.cuda.stream(ext):
                    sc.render_dev(msk_dist.shard_desc(rd, rank, world), peer.ptr)
                    peer.reduce()
                peer.check()
                if rank == 0:
                    out["reduced"] = film.cpu().numpy().copy()
                    whole, _ = sc.render(rd)
                    out["whole"] = whole
            film = None
            peer.close()
        result.put((rank, out, None))
    except Exception:  # noqa: BLE001 -- reported to the parent
        result.put((rank, None, traceback.format_exc()))


def test_peer_film_reduction_two_ranks():
    import torch
    ndev = torch.cuda.device_count()
    world = 2
    mpc = mp.get_context("spawn")
    to_root = [mpc.Pipe(duplex=False) for _ in range(world)]
    from_root = [mpc.Pipe(duplex=False) for _ in range(world)]
    result = mpc.Queue()
    procs = []
    for r in range(world):
        # rank r sends on to_root[r][1] and receives on from_root[r][0]; the root holds the other ends
        recv_ends = {q: to_root[q][0] for q in range(world)} if r == 0 else {}
        send_ends = {q: from_root[q][1] for q in range(world)} if r == 0 else {}
        tr = recv_ends if r == 0 else {r: to_root[r][1]}
        fr = send_ends if r == 0 else {r: from_root[r][0]}
        p = mpc.Process(target=_worker, args=(r, world, ndev, tr, fr, result))
        p.start()
        procs.append(p)
    got = {}
    try:
        for _ in range(world):
            rank, out, err = result.get(timeout=240)
            assert err is None, f"rank {rank} failed:\n{err}"
            got[rank] = out
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    root = got[0]
    assert root["sums_ok"]
    assert np.isfinite(root["reduced"]).all()
    # same samples, same seeds; only the float summation order of the film differs (batches vs ranks)
    np.testing.assert_allclose(root["reduced"], root["whole"], rtol=2e-5, atol=1e-6)

"""Multi-GPU sharding of the render job: one process per GPU, the scene and BVH
replicated, the camera samples of every pixel partitioned by SAMPLE RANGE, and one sum-reduction of the
XYZAW film to rank 0 -- by default the library's own kernel that pulls the peers' films over NVLink peer memory
(PeerFilm below, csrc/msk_peer.cu: no NCCL call and no host barrier on the data path); torch.distributed's reduce
(NCCL, gloo in the CPU tests) is the alternative and the plumbing that exchanges the IPC handles.

The reference has a single shared-memory level of parallelism -- tbb::parallel_for over 32x32 tiles with a
mutex-guarded Film::put (src/librender/integrator.cpp:54-75, films/hdrfilm.cpp:43-46).  Because sample
(pixel p, index s) is seeded with Sampler::seed(p * spp + s) (samplers/independent.cpp:20-26), the union of
the ranks' sample ranges reproduces the single-GPU job up to float summation order, and the film channels
are plain sums (the division by W happens in HDRFilm::image, hdrfilm.cpp:71-76), so the only exchange step
is that reduction.  Tile partitioning would still need a sum because the 2-pixel filter border of
neighbouring tiles overlaps (imageblock.cpp:40-48).
"""
from __future__ import annotations

import copy

from . import capi


def shard_samples(spp: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced sample range [begin, end) of rank `rank`; ranges tile [0, spp) exactly.
    Ranks beyond `spp` get an empty range."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(int(spp), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_desc(rd: capi.MskRenderDesc, rank: int, world: int) -> capi.MskRenderDesc:
    """The MskRenderDesc of this rank: same job (spp, seeds, depths), its own sub-range of rd's sample range."""
    out = copy.copy(rd)
    n = rd.sample_end - rd.sample_begin
    b, e = shard_samples(n, rank, world)
    out.sample_begin, out.sample_end = rd.sample_begin + b, rd.sample_begin + e
    out.clear_film = 1
    return out


def reduce_film(film, dst: int = 0, group=None):
    """Sum the per-rank XYZAW films into rank `dst` (in place).  `film` is a torch tensor (CUDA for NCCL,
    CPU for gloo).  A no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def render_sharded(scene: capi.Scene, rd: capi.MskRenderDesc, film_dev, rank: int, world: int, group=None,
                   host_out=None, peer=None):
    """Render this rank's share of `rd` into the CUDA tensor `film_dev` (H x W x 5 float32) on the context's
    stream, reduce to rank 0, and (rank 0, if `host_out` -- a pinned CPU tensor -- is given) copy the final
    film to the host.  Returns this rank's MskStats."""
    import torch
    ext = torch.cuda.ExternalStream(scene.ctx.stream, device=film_dev.device)
    with torch.cuda.stream(ext):
        stats = scene.render_dev(shard_desc(rd, rank, world), film_dev.data_ptr())
        if peer is not None:  # film_dev is peer.tensor(): the library's NVLink peer-memory reduction
            peer.reduce()
            peer.check()  # drains the stream; raises if a device-side wait timed out (the film would be this rank's only)
        else:
            reduce_film(film_dev, 0, group)
        if host_out is not None and rank == 0:
            host_out.copy_(film_dev, non_blocking=True)
        ext.synchronize()
    return stats


class PeerFilm:
    """The rank's film in CUDA-IPC-exportable device memory, wired to its peers (one process per GPU).

    `exchange` is how the 64-byte IPC handles travel between the processes: by default torch.distributed's object
    all-gather on `group`; tests pass their own (a pipe).  The root opens every other rank's film in rank order;
    `reduce()` then runs msk_gpu_reduce_film on the context's stream: ONE kernel on the root that waits for the
    peers on the device and sums their films into its own through NVLink, in rank order (deterministic)."""

    def __init__(self, ctx: capi.Context, shape, rank: int, world: int, root: int = 0, group=None, exchange=None):
        self.rank, self.world, self.root = rank, world, root
        self.share = capi.FilmShare(ctx, shape)
        mine = self.share.export()
        if exchange is None:
            import torch.distributed as dist

            def exchange(handle):
                out = [None] * world
                dist.all_gather_object(out, handle, group=group)
                return out
        handles = exchange(mine) if world > 1 else [mine]
        if rank == root:
            self.share.open_peers([h for r, h in enumerate(handles) if r != root])
        self._step = 0

    @property
    def ptr(self) -> int:
        return self.share.ptr

    def tensor(self, device):
        import torch
        return torch.as_tensor(self.share, device=device)

    def reduce(self):
        """Asynchronous; call once per step on every rank, after the rank's render was enqueued."""
        if self.world > 1:
            self._step += 1
            self.share.reduce(self.rank == self.root, self._step)

    def check(self):
        self.share.check()

    def close(self):
        self.share.close()

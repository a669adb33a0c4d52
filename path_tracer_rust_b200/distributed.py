"""Multi-GPU driver: one process per GPU, samples-per-pixel sharded across ranks, one fp32 sum-reduce.

The reference is single-process (rayon over pixels, src/render/mod.rs:1021-1023).  Every (pixel, sample) is
independent and the reference clamps only after averaging (mod.rs:849-856), so rank g renders the global sample
indices [g*spp/G, (g+1)*spp/G) of every pixel into an UNCLAMPED fp32 sum framebuffer; the framebuffers are added
with one NCCL reduce over NVLink (the path's only collective) and rank 0 resolves (sum/spp, clamp).  Global sample
indices keep the 2x2 sub-pixel pattern (mod.rs:814-815) and the RNG streams partition-invariant, so the image does
not depend on G except for the fp32 summation order of the G partial sums.
"""
from __future__ import annotations

from typing import Optional, Protocol, Tuple


def shard_samples(spp_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, begin+count) of rank `rank`; contiguous, disjoint, covers [0, spp_total)."""
    if world_size < 1 or not (0 <= rank < world_size) or spp_total < 0:
        raise ValueError("bad shard arguments")
    begin = rank * spp_total // world_size
    end = (rank + 1) * spp_total // world_size
    return begin, end - begin


class ShardRenderer(Protocol):
    def render_sum(self, spp_begin: int, spp_count: int): ...      # -> torch.Tensor [W*H*3] fp32 on the rank's device
    def resolve(self, sum_fb, spp_total: int): ...                 # -> torch.Tensor [W*H*3] fp32 (mean, clamped)


class CudaShardRenderer:
    """The product implementation: ptb_render_device into a torch-owned device buffer, ptb_resolve_device."""

    def __init__(self, backend, width: int, height: int, seed: int, device):
        import torch
        self.torch = torch
        self.be, self.W, self.H, self.seed, self.device = backend, width, height, seed, device
        self.fb = torch.zeros(width * height * 3, dtype=torch.float32, device=device)

    def render_sum(self, spp_begin: int, spp_count: int):
        torch = self.torch
        self.fb.zero_()
        stream = torch.cuda.current_stream(self.device)
        if spp_count > 0:
            self.be.render_device(self.W, self.H, spp_count, self.fb.data_ptr(), spp_begin=spp_begin, seed=self.seed,
                                  stream=stream.cuda_stream)
        return self.fb

    def resolve(self, sum_fb, spp_total: int):
        stream = self.torch.cuda.current_stream(self.device)
        self.be.resolve_device(sum_fb.data_ptr(), sum_fb.numel(), spp_total, sum_fb.data_ptr(), stream=stream.cuda_stream)
        return sum_fb


def render_sharded(renderer: ShardRenderer, spp_total: int, rank: int = 0, world_size: int = 1, group=None, dst: int = 0):
    """Runs one sharded frame.  Returns the resolved image tensor on rank `dst`, None elsewhere."""
    begin, count = shard_samples(spp_total, world_size, rank)
    fb = renderer.render_sum(begin, count)
    if world_size > 1:
        import torch.distributed as dist
        dist.reduce(fb, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    return renderer.resolve(fb, spp_total)


def slice_floats(n_floats: int, world_size: int, rank: int) -> Tuple[int, int]:
    """16-byte aligned [first, first+count) slice of the framebuffer that `rank` reduces and resolves."""
    quads = (n_floats + 3) // 4
    q0, q1 = rank * quads // world_size, (rank + 1) * quads // world_size
    first, end = min(q0 * 4, n_floats), min(q1 * 4, n_floats)
    return first, end - first


class PeerMemoryFrame:
    """Multi-GPU frame driver that keeps the reduction on NVLink peer memory instead of a library collective.

    Every rank owns a sum framebuffer allocated by its ptb_ctx and exported with CUDA IPC; rank `dst` additionally owns the
    output image.  A frame is: render own sample shard -> barrier -> ONE kernel per rank that loads its slice of every rank's
    framebuffer over NVLink, adds them in rank order (deterministic), divides by spp, clamps (mod.rs:849-856) and stores the
    result straight into rank dst's output buffer -> barrier -> rank dst copies the image to the host.
    torch.distributed is used for the handle exchange and the barriers only.
    """

    def __init__(self, backend, width: int, height: int, seed: int, rank: int, world_size: int, group=None, dst: int = 0):
        """Collective over `group`: either every rank ends up with a working frame or every rank raises BackendError."""
        import torch.distributed as dist
        from .api import BackendError
        self.be, self.W, self.H, self.seed = backend, width, height, seed
        self.rank, self.world, self.group, self.dst, self.dist = rank, world_size, group, dst, dist
        self.n_floats = width * height * 3
        self.fb = self.out = 0
        self._opened, self.peer_fb, self.dst_out = [], [], 0

        def agree(value):  # every rank learns every rank's value (None = that rank failed)
            if world_size == 1:
                return [value]
            got = [None] * world_size
            dist.all_gather_object(got, value, group=group)
            return got

        mine, err = None, None
        try:
            self.fb = backend.device_alloc(self.n_floats * 4)
            self.out = backend.device_alloc(self.n_floats * 4) if rank == dst else 0
            mine = {"fb": backend.ipc_export(self.fb), "out": backend.ipc_export(self.out) if rank == dst else None}
        except Exception as exc:  # noqa: BLE001
            err = exc
        handles = agree(mine)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                for g, h in enumerate(handles):
                    if g == rank:
                        self.peer_fb.append(self.fb)
                    else:
                        p = backend.ipc_open(h["fb"])
                        self._opened.append(p)
                        self.peer_fb.append(p)
                if rank == dst:
                    self.dst_out = self.out
                else:
                    self.dst_out = backend.ipc_open(handles[dst]["out"])
                    self._opened.append(self.dst_out)
            except Exception as exc:  # noqa: BLE001
                err = exc
        ok = all(agree(err is None))
        if not ok:
            self._release_local()
            raise BackendError(-2, f"peer-memory frame could not be set up on every rank (rank {rank}: {err})")

    def _release_local(self):
        for p in self._opened:
            try:
                self.be.ipc_close(p)
            except Exception:  # noqa: BLE001
                pass
        self._opened = []
        for p in (self.fb, self.out):
            if p:
                try:
                    self.be.device_free(p)
                except Exception:  # noqa: BLE001
                    pass
        self.fb = self.out = 0

    def _barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def render(self, spp_total: int, host_out=None, to_host: bool = True):
        """One frame.  Returns the host image (numpy [W*H,3]) on rank dst (None elsewhere, or when to_host is False and the
        resolved image is left in rank dst's device buffer `self.out`)."""
        import numpy as np
        begin, count = shard_samples(spp_total, self.world, self.rank)
        self.be.device_memset(self.fb, 0, self.n_floats * 4)
        if count > 0:
            self.be.render_device(self.W, self.H, count, self.fb, spp_begin=begin, seed=self.seed, stream=0)
        self.be.device_sync()
        self._barrier()                                    # every partial sum is complete and visible
        first, n = slice_floats(self.n_floats, self.world, self.rank)
        self.be.peer_reduce_resolve(self.peer_fb, first, n, spp_total, self.dst_out)
        self.be.device_sync()
        self._barrier()                                    # every slice has landed in rank dst's buffer
        if self.rank != self.dst or not to_host:
            return None
        img = host_out if host_out is not None else np.empty((self.W * self.H, 3), np.float32)
        self.be.device_to_host(img, self.out)
        return img

    def close(self):
        self._barrier()
        for p in self._opened:
            self.be.ipc_close(p)
        self._opened = []
        self._barrier()       # nobody has a peer's buffer mapped any more: the owners may free
        self._release_local()

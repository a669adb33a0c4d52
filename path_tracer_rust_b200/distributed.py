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

#!/usr/bin/env python
"""N-rank check of the peer-memory frame driver against the NCCL path and the oracle (run under torchrun on N GPUs)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import path_tracer_rust_b200 as P
from path_tracer_rust_b200.distributed import CudaShardRenderer, PeerMemoryFrame, render_sharded, shard_samples

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
W, H, spp = 96, 64, 13
be = P.Backend(local)
be.upload_scene(P.Scene.load("cornell"))
frame = PeerMemoryFrame(be, W, H, seed=5, rank=rank, world_size=world)
img = frame.render(spp)
shard = CudaShardRenderer(be, W, H, seed=5, device=dev)
ref = render_sharded(shard, spp, rank, world)
if rank == 0:
    import oracle_lib as O
    osc = O.OracleScene(os.path.join(ROOT, "scenes", "cornell.json"))
    acc = None
    for g in range(world):                      # partial sums added in rank order, like the kernel does
        b, c = shard_samples(spp, world, g)
        part = osc.render_sum(W, H, c, spp_begin=b, seed=5)[0]
        acc = part if acc is None else (acc + part).astype(np.float32)
    want = O.resolve(acc, spp)
    assert np.array_equal(img.view(np.uint32), want.view(np.uint32)), "peer-memory frame differs from the oracle"
    np.testing.assert_allclose(img.reshape(-1), ref.cpu().numpy(), rtol=1e-5, atol=1e-6)
    print(f"PEER_REDUCE_OK world={world} bit-exact vs oracle (rank-ordered partial sums), matches the NCCL path", flush=True)
frame.close()
if os.environ.get("PTB_PEER_CHECK_QUICK"):      # the pytest wrapper (tests/test_gpu_fullsize.py) only wants the parity check
    be.close()
    dist.destroy_process_group()
    sys.exit(0)
# timing at 4K: peer-memory reduce+resolve vs NCCL reduce + resolve
W, H = 3840, 2160
frame = PeerMemoryFrame(be, W, H, seed=5, rank=rank, world_size=world)
shard = CudaShardRenderer(be, W, H, seed=5, device=dev)
for name in ("peer", "nccl"):
    ts = []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "peer":
            frame.render(world)                 # 1 spp per rank: the frame time is dominated by the reduction
        else:
            out = render_sharded(shard, world, rank, world)
            if rank == 0:
                out.cpu()
        torch.cuda.synchronize(); dist.barrier()
        ts.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        print(f"4K frame at {world} spp total, {name}: {min(ts):.2f} ms (best of 4)", flush=True)
frame.close()
be.close()
dist.destroy_process_group()

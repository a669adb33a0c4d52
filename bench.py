#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing backend (see BASELINE.json / SURVEY.md 8d).

Metric: Mpaths/s (pixel-samples per second) on scenes/cornell.json at 3840x2160 x 4096 spp, the samples sharded across
the N GPUs of one box (strong scaling: total work fixed) and summed with one NCCL reduce; Mray-segments/s is reported
beside it.  A "step" is one full frame (all 4096 spp of every pixel).  `value` is timed with the scene resident on the
device; `e2e` goes through the public host API with the scene uploaded from host memory and the image read back to
pinned host memory inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--spp S]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

`--impl reference` times the CPU path (the strict-fp32 C restatement in oracle/; the Rust reference cannot be built in this
image) on a bounded sample of the same workload, on rank 0 only.

Besides the headline line's own keys the JSON line carries
  "extra_workloads": the other BASELINE.json configs measured in the same run (C1 cornell 450x300x100 with an un-extrapolated
                     CPU time of the whole frame, C2 sphere scenes, C3 mesh.json 1080p x 1024, C5 the synthetic 1.31 M-triangle scene
                     4K x 1024), each with value, e2e, roofline and clocks; at N > 1 they are sharded over the ranks like the headline;
  "cabi_multi":      (N > 1) the headline workload once more through ptb_create_multi -- ONE process, one context, N GPUs, no torch on
                     the data path -- driven by rank 0 while the other ranks wait.
`--extras none` (or a comma list of workload names) limits the former, `--no-cabi-multi` skips the latter.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (scene id, W, H, spp)   -- BASELINE.json configs
    "cornell4k": ("cornell", 3840, 2160, 4096),      # configs[3], the multi-GPU headline
    "cornell_default": ("cornell", 450, 300, 100),   # configs[0]
    "single_sphere_1080p": ("single-sphere", 1920, 1080, 256),
    "three_spheres_1080p": ("three-spheres", 1920, 1080, 256),
    "mesh_1080p": ("mesh", 1920, 1080, 1024),        # configs[2]
    "synthetic4k": ("synthetic", 3840, 2160, 1024),  # configs[4]: 1.31 M-triangle displaced icosphere + 10 k spheres
}
SYNTHETIC = dict(level=8, n_spheres=10000, scale=10.0)   # S = 10 as BASELINE.md C5 / SURVEY.md 8d state


def resolve_scene(scene_id: str, rank: int = 0):
    """(json path, base dir) of a workload scene; the synthetic scene is generated procedurally (seeded) on first use."""
    if scene_id != "synthetic":
        return os.path.join(ROOT, "scenes", f"{scene_id}.json"), ROOT
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tools", "make_synthetic_scene.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    out = os.path.join(os.environ.get("PTB_SCRATCH", "/tmp"), f"ptb_synthetic_rank{rank}")
    return mk.make_synthetic(out, **SYNTHETIC), out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        while not self._stop.wait(0.02):   # 20 ms: the small configs finish in tens of milliseconds
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break

    def start(self):
        if self._nvml is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_run(scene_id: str, W: int, H: int, target_s: float, threads: int | None = None):
    """Times the CPU path (oracle in the reference's own mode: sequential RNG, libm sin/cos, recursive radiance,
    shuffled pixels, all host threads) on a bounded sample: the same scene and camera at 1/16 of the resolution per axis,
    spp chosen by a short calibration so the run takes about `target_s` seconds."""
    import oracle_lib as O
    threads = threads or os.cpu_count() or 1
    path, base = resolve_scene(scene_id)
    osc = O.OracleScene(path, base)
    w, h = max(W // 16, 16), max(H // 16, 16)
    if scene_id == "synthetic":   # the reference algorithm is O(1.3 M triangle tests) per ray: only a tiny crop is affordable
        w, h = 48, 27
    ref = dict(rng=O.RNG_SEQ, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE, threads=threads, shuffle=1)
    cal_spp = 1 if scene_id == "synthetic" else 4
    t0 = time.perf_counter()
    osc.render_sum(w, h, cal_spp, seed=1, **ref)
    cal = max(time.perf_counter() - t0, 1e-4)
    spp = int(max(1, min(4096, cal_spp * target_s / cal)))
    t0 = time.perf_counter()
    _, st = osc.render_sum(w, h, spp, seed=2, **ref)
    dt = time.perf_counter() - t0
    samples = w * h * spp
    return {"seconds": dt, "samples": samples, "segments": int(st[0]), "mpaths_s": samples / dt * 1e-6,
            "mseg_s": int(st[0]) / dt * 1e-6, "threads": threads,
            "tests_per_segment": {"sphere": int(st[1]) / max(int(st[0]), 1), "gate": int(st[2]) / max(int(st[0]), 1),
                                  "triangle": int(st[3]) / max(int(st[0]), 1)},
            "sample": f"{scene_id}.json same camera at {w}x{h} ({w * h / (W * H):.5f} of the {W}x{H} pixels) x {spp} spp, "
                      f"{threads} threads, pixel loop only"}


def cpu_full_frame(scene_id: str, W: int, H: int, spp: int, threads: int | None = None):
    """The WHOLE frame on the CPU, nothing extrapolated: the oracle in the reference's own mode (sequential
    RNG, libm sin/cos, recursive radiance, shuffled pixels) on all host threads.  Only sensible for the reference's default
    450x300 x 100 spp frame (13.5 M samples, seconds)."""
    import oracle_lib as O
    threads = threads or os.cpu_count() or 1
    path, base = resolve_scene(scene_id)
    osc = O.OracleScene(path, base)
    t0 = time.perf_counter()
    _, st = osc.render_sum(W, H, spp, seed=2, rng=O.RNG_SEQ, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE, threads=threads, shuffle=1)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "samples": W * H * spp, "segments": int(st[0]), "mpaths_s": W * H * spp / dt * 1e-6,
            "mseg_s": int(st[0]) / dt * 1e-6, "threads": threads,
            "tests_per_segment": {"sphere": int(st[1]) / max(int(st[0]), 1), "gate": int(st[2]) / max(int(st[0]), 1),
                                  "triangle": int(st[3]) / max(int(st[0]), 1)},
            "sample": f"the whole frame: {scene_id}.json {W}x{H} x {spp} spp in {dt:.2f} s on {threads} threads (C port of the rayon path, "
                      "pixel loop only, nothing extrapolated)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    scene_id, W, H, spp = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    per_step = max(2.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cpu_reference_run(scene_id, W, H, per_step)
    runs = [cpu_reference_run(scene_id, W, H, per_step) for _ in range(args.steps)]
    tot_t = sum(r["seconds"] for r in runs)
    val = sum(r["samples"] for r in runs) / tot_t * 1e-6
    seg = sum(r["segments"] for r in runs) / tot_t * 1e-6
    line = {"impl": "reference", "metric": "Mpaths/s", "value": val, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_t / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: scenes/{scene_id}.json {W}x{H} x {spp} spp", "scene": scene_id, "width": W,
                       "height": H, "spp": spp},
            "mray_segments_per_s": seg,
            "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": runs[-1]["threads"], "kind": "port", "sample": runs[-1]["sample"]},
            "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


EXTRAS = {
    # name: (warm-up frames, warm-up spp (0 = full), timed frames, e2e frames)
    "cornell_default": (3, 0, 20, 5),
    "single_sphere_1080p": (3, 0, 20, 5),
    "three_spheres_1080p": (3, 0, 20, 5),
    "mesh_1080p": (1, 0, 3, 1),
    "synthetic4k": (1, 32, 2, 1),        # 45 s per 1024-spp frame on one GPU: the warm-up frame is a short one
}


# rough wall-clock cost of an extra workload on ONE GPU in seconds (scene generation / loading + its frames), used by --time-budget
EXTRA_COST_S = {"cornell_default": 5, "single_sphere_1080p": 3, "three_spheres_1080p": 3, "mesh_1080p": 15, "synthetic4k": 60 + 130}


def main():
    t_start = time.perf_counter()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cornell4k", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (development only; recorded in config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default=None, help="'all', 'none' or a comma list of workloads measured beside the headline "
                    "(default: all for the default headline workload at full spp, none otherwise)")
    ap.add_argument("--no-cabi-multi", action="store_true", help="N > 1: skip the single-process ptb_create_multi measurement")
    ap.add_argument("--time-budget", type=float, default=780.0,
                    help="seconds of wall clock this run may use: an extra workload that would not fit is skipped (and listed as skipped), "
                         "the headline line is always printed")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="backend option (ptb_set_option), development only; recorded in config")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"],
                    help="N>1: sum the ranks' framebuffers with the fused peer-memory kernel (default) or with an NCCL reduce")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import path_tracer_rust_b200 as P
    from path_tracer_rust_b200.distributed import CudaShardRenderer, PeerMemoryFrame, render_sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the backend has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # host-side waits: an NCCL barrier would spin a kernel on the waiting ranks' GPUs
    else:
        dist = None
        cpu_group = None

    # `python bench.py --gpus N` WITHOUT torchrun (one process): all N GPUs through one multi-GPU context (ptb_create_multi)
    single_process_multi = world == 1 and args.gpus > 1
    if single_process_multi and torch.cuda.device_count() < args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but only {torch.cuda.device_count()} CUDA device(s) are visible")
    n_gpus = args.gpus if single_process_multi else world
    be = P.Backend(list(range(args.gpus))) if single_process_multi else P.Backend(local_rank)
    for kv in args.opt:
        name, _, value = kv.partition("=")
        be.set_option(name, float(value))
    l2_flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    peaks = load_peaks()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sms * 128 * peaks["sm_max_mhz"] * 1e6 * 1e-12          # T lane-ops/s, un-fused (FMA is barred by parity)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(workload: str, spp_override: int, warmup: int, warmup_spp: int, steps: int, e2e_steps: int, cpu_mode: str):
        """One workload on all ranks: `warmup` untimed frames, `steps` frames timed with CUDA events (scene resident, L2 flushed
        between frames, max over ranks), `e2e_steps` frames through the host API (scene upload + render + reduce + resolve + D2H).
        cpu_mode: 'sample' = bounded CPU sample (cpu_reference_run), 'frame' = the whole frame on the CPU, 'none'."""
        scene_id, W, H, spp = WORKLOADS[workload]
        reduced = bool(spp_override)
        if spp_override:
            spp = spp_override
        scene_path, scene_base = resolve_scene(scene_id, rank)
        scene = P.Scene.load(scene_path, base_dir=scene_base)
        be.upload_scene(scene)
        use_peer = world > 1 and args.reduce == "peer"
        frame = None
        if use_peer:
            # CUDA IPC needs every rank to see every peer; if any rank cannot set it up, all ranks fall back to the NCCL reduce
            ok = 1
            try:
                frame = PeerMemoryFrame(be, W, H, seed=2026, rank=rank, world_size=world)
            except Exception as exc:  # noqa: BLE001
                print(f"[rank {rank}] peer-memory frame unavailable ({exc}); falling back to NCCL", file=sys.stderr, flush=True)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                frame = None
                use_peer = False
        shard = CudaShardRenderer(be, W, H, seed=2026, device=dev) if not (use_peer or single_process_multi) else None
        nfl = W * H * 3
        host_img = torch.empty(nfl, dtype=torch.float32).pin_memory() if rank == 0 else None
        host_np = host_img.numpy().reshape(-1, 3) if rank == 0 else None

        def step_resident(n_spp):
            if single_process_multi:   # the multi-GPU context renders, reduces over peer memory and copies to the host in one call
                be.render(W, H, n_spp, seed=2026, out=host_np)
                return None
            if use_peer:
                return frame.render(n_spp, to_host=False)
            return render_sharded(shard, n_spp, rank, world)

        def step_e2e():
            be.upload_scene(scene)                       # H2D of the scene (flatten + upload + device BVH build)
            if world == 1:   # (also the single-process multi-GPU context)
                # the reference-facing call itself: ptb_render() with a HOST output buffer (here pinned), blocking like render()
                be.render(W, H, spp, seed=2026, out=host_np)
                return
            if use_peer:
                frame.render(spp, host_out=host_np)      # reduce + resolve over peer memory, then D2H on rank 0
                return
            img = render_sharded(shard, spp, rank, world)
            if rank == 0:
                host_img.copy_(img, non_blocking=True)   # D2H of the resolved image
                torch.cuda.current_stream().synchronize()

        for _ in range(max(warmup, 0)):
            step_resident(warmup_spp or spp)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        seg_total, launches, times = 0, 0, []
        for _ in range(steps):
            l2_flush.fill_(1.0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_resident(spp)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            st = be.stats()
            seg_total += st["segments"]
            launches += st["kernel_launches"] + (1 if ((rank == 0 or use_peer) and not single_process_multi) else 0)
        barrier()
        clocks = sampler.stop()
        st_last = be.stats()
        kernel_ms_last = st_last["render_ms"]
        e2e_times = []
        for _ in range(e2e_steps):
            barrier()
            t0 = time.perf_counter()
            step_e2e()
            torch.cuda.synchronize()
            e2e_times.append((time.perf_counter() - t0) * 1e3)
        barrier()
        t_sum = torch.tensor([sum(times), sum(e2e_times), float(seg_total), float(launches)], dtype=torch.float64, device=dev)
        if dist is not None:
            t_max = t_sum.clone()
            dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
            total_ms, e2e_ms = float(t_max[0]), float(t_max[1])
            seg_total, launches = float(t_sum[2]), int(t_sum[3])
        else:
            total_ms, e2e_ms = float(t_sum[0]), float(t_sum[1])
        if frame is not None:
            frame.close()
        if rank != 0:
            return None

        samples_per_step = W * H * spp
        value = samples_per_step * steps / (total_ms * 1e-3) * 1e-6
        seg_rate = seg_total / (total_ms * 1e-3) * 1e-6
        e2e_value = samples_per_step * e2e_steps / (e2e_ms * 1e-3) * 1e-6 if e2e_steps else None
        sd = scene._desc.contents
        scene_bytes = int(sd.n_objects) * 80 + int(sd.n_triangles) * 36 + 36
        cpu = None
        if cpu_mode == "sample" and n_gpus == 1:
            cpu = cpu_reference_run(scene_id, W, H, 12.0)
        elif cpu_mode == "frame" and n_gpus == 1:
            cpu = cpu_full_frame(scene_id, W, H, spp)
        # roofline of the dominant kernel: FP32 issue slots.  Algorithmic flops per segment follow SURVEY.md 8d:
        # 17 per sphere/gate test + 45 per triangle test + 30 per four-wide BVH node + 120 shading, with the reference algorithm's
        # own test counts for the shared-memory list and device-counted BVH work.
        if st_last["n_bvh_nodes"] > 0 and st_last["segments"] > 0:
            seg_last = float(st_last["segments"])
            tps = {"sphere": 0.0, "gate": float(st_last["n_loose_objects"]),
                   "triangle": float(st_last["n_loose_triangles"]) + st_last["bvh_prims_tested"] / seg_last,
                   "bvh_nodes": st_last["bvh_nodes_visited"] / seg_last}
            dominant = "k_wf_trace (+ k_wf_shade, k_wf_generate, k_wf_accumulate; whole wavefront pipeline timed)"
        else:
            tps = cpu["tests_per_segment"] if cpu and "tests_per_segment" in cpu else None
            if tps is None:
                if scene_id == "cornell":    # the reference algorithm's own counts (oracle statistics): 4 spheres, 7 gates, 11.04 triangles
                    tps = {"sphere": 4.0, "gate": 7.0, "triangle": 11.0}
                elif st_last["n_loose_triangles"] == 0:   # sphere-only scene: intersect_scene tests every sphere for every segment
                    tps = {"sphere": float(st_last["n_loose_objects"]), "gate": 0.0, "triangle": 0.0}
                else:
                    tps = {"sphere": 0.0, "gate": 0.0, "triangle": 0.0}
            tps = dict(tps, bvh_nodes=0.0)
            dominant = "k_render" if not (W * H < 2 * sms * 24 * 32) else "k_wf_shade (+ k_wf_generate, k_wf_accumulate: wavefront integrator, small frame)"
        flops_per_seg = 17.0 * (tps["sphere"] + tps["gate"]) + 45.0 * tps["triangle"] + 30.0 * tps["bvh_nodes"] + 120.0
        # secondary (HBM) view for BVH scenes: queue entries written by generate / shade (76 B), read by trace (44 B of the
        # segments that reach the BVH: counted as all) and by shade (72 B), hit written back by trace (8 B); BVH nodes and
        # primitives are L2-resident and not counted
        hbm = None
        if tps["bvh_nodes"] > 0:
            bytes_per_seg = 76.0 + 44.0 + 72.0 + 8.0
            gbs = seg_total / max(n_gpus, 1) / steps * bytes_per_seg / (kernel_ms_last * 1e-3) * 1e-9 if kernel_ms_last > 0 else None
            hbm = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"] if gbs else None,
                   "bytes_per_segment": bytes_per_seg,
                   "note": "queue traffic only (algorithmic bytes per ray segment); BVH nodes / primitives are served from L2"}
        seg_per_gpu = seg_total / max(n_gpus, 1) / steps
        kern_s = kernel_ms_last * 1e-3
        achieved = seg_per_gpu * flops_per_seg / kern_s * 1e-12 if kern_s > 0 else None
        traffic, ncu_note = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and not reduced:
            tj = json.load(open(tpath)).get(workload)
            if tj:
                traffic, ncu_note = tj["dram_bytes_per_launch"], {k: tj[k] for k in ("kernel", "launch", "algorithmic_bytes_per_launch", "ncu", "source")}
        return {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / max(steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{workload}: scenes/{scene_id}.json {W}x{H} x {spp} spp, spp sharded over {n_gpus} GPU(s), "
                                   + ("one process, ptb_create_multi: fused peer-memory reduce+resolve kernel over NVLink, image copied to the host "
                                      "inside every timed step" if single_process_multi else
                                      ("fused peer-memory reduce+resolve kernel over NVLink (CUDA IPC)" if use_peer
                                       else ("NCCL fp32 sum-reduce" if world > 1 else "no reduce step"))), "scene": scene_id, "width": W, "height": H, "spp": spp,
                       "spp_reduced_for_development": reduced, "l2": "flushed between steps (256 MiB device write)",
                       "parallelism": f"spp-shard x{n_gpus}", "backend_options": list(args.opt),
                       "warmup_spp": warmup_spp or spp},
            "mray_segments_per_s": seg_rate, "segments_per_sample": seg_total / (samples_per_step * steps),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": nfl * 4,
                    "ms_per_step": e2e_ms / e2e_steps if e2e_steps else None, "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": fp32_peak, "unit": "Tlane-op/s",
                         "frac": (achieved / fp32_peak) if achieved else None, "traffic": traffic, "traffic_detail": ncu_note,
                         "kernel": dominant, "flops_per_segment": flops_per_seg, "tests_per_segment": tps, "kernel_ms": kernel_ms_last,
                         "peak_source": f"SMs({sms}) x 128 lanes x sm_max_mhz({peaks['sm_max_mhz']}) from MEASURED_PEAKS.json ({peaks['source']}); "
                                        "no tensor or HBM bound applies: scene and path state live in shared memory / registers"},
            "roofline_hbm": hbm,
            "cpu_baseline": ({"value": cpu["mpaths_s"], "unit": "Mpaths/s", "cores": cpu["threads"], "kind": "port",
                              "sample": cpu["sample"], "mray_segments_per_s": cpu["mseg_s"]} if cpu else None),
        }

    # ---- headline -----------------------------------------------------------------------------------------------------
    e2e_steps = max(1, min(args.steps, 2))      # the resident steps already warmed the device up
    line = measure(args.workload, args.spp, args.warmup, 0, args.steps, e2e_steps, "none" if args.no_cpu_baseline else "sample")

    # ---- the other BASELINE configs, same run ---------------------------------------------------------------------------
    if args.extras is None:
        extras = [w for w in EXTRAS if w != args.workload] if (args.workload == "cornell4k" and not args.spp) else []
    elif args.extras in ("none", ""):
        extras = []
    elif args.extras == "all":
        extras = [w for w in EXTRAS if w != args.workload]
    else:
        extras = [w for w in args.extras.split(",") if w in EXTRAS]
    extra_lines = {}
    for w in extras:
        wu, wu_spp, k, ke = EXTRAS[w]
        # every rank takes the same decision: rank 0's clock, broadcast
        est = (60 if w == "synthetic4k" else 0) + (EXTRA_COST_S[w] - (60 if w == "synthetic4k" else 0)) / max(n_gpus, 1)
        fits = torch.tensor([1 if (time.perf_counter() - t_start) + est <= args.time_budget else 0], dtype=torch.int32, device=dev)
        if dist is not None:
            dist.broadcast(fits, src=0)
        if int(fits.item()) == 0:
            if rank == 0:
                extra_lines[w] = {"skipped": f"would not fit the time budget of {args.time_budget:.0f} s "
                                             f"({time.perf_counter() - t_start:.0f} s used, ~{est:.0f} s needed)"}
            continue
        t0 = time.perf_counter()
        try:
            r = measure(w, 0, wu, wu_spp, k, ke, "frame" if (w == "cornell_default" and not args.no_cpu_baseline) else "none")
        except Exception as exc:  # noqa: BLE001  (an extra must never take the headline down)
            r = {"error": f"{type(exc).__name__}: {exc}"} if rank == 0 else None
        if rank == 0:
            r["wall_s"] = time.perf_counter() - t0
            extra_lines[w] = r

    # ---- N > 1: the same frame through ONE process and ONE context (ptb_create_multi), rank 0 drives all GPUs -------------
    cabi = None
    if world > 1 and not args.no_cabi_multi:
        barrier()
        if rank == 0:
            try:
                scene_id, W, H, spp = WORKLOADS[args.workload]
                spp = args.spp or spp
                scene_path, scene_base = resolve_scene(scene_id, rank)
                scene = P.Scene.load(scene_path, base_dir=scene_base)
                mb = P.Backend(list(range(world)))
                for kv in args.opt:
                    name, _, value = kv.partition("=")
                    mb.set_option(name, float(value))
                host = torch.empty(W * H * 3, dtype=torch.float32).pin_memory().numpy().reshape(-1, 3)
                mb.upload_scene(scene)
                mb.render(W, H, max(spp // 8, world), seed=2026, out=host)           # warm-up: allocations, clocks
                ts = []
                for _ in range(2):
                    t0 = time.perf_counter()
                    mb.upload_scene(scene)
                    mb.render(W, H, spp, seed=2026, out=host)
                    ts.append(time.perf_counter() - t0)
                st = mb.stats()
                cabi = {"value": W * H * spp / min(ts) * 1e-6, "unit": "Mpaths/s", "ms_per_step": min(ts) * 1e3, "steps": 2, "n_gpus": world,
                        "device_ms_slowest_gpu": st["render_ms"], "gpu_launches": st["kernel_launches"],
                        "how": "ptb_create_multi + ptb_upload_scene + ptb_render from rank 0's process: one host thread per GPU, "
                               "spp split by global sample index, fused peer-memory reduce+resolve, D2H to pinned host memory; "
                               "wall clock around the calls (end to end), best of 2"}
                mb.close()
            except Exception as exc:  # noqa: BLE001
                cabi = {"error": f"{type(exc).__name__}: {exc}"}
        # the other ranks wait on the HOST (gloo): a pending NCCL barrier is a kernel spinning on their GPU, and two processes on one
        # GPU are time-sliced -- it would halve the speed of the devices rank 0 is driving (measured)
        dist.barrier(group=cpu_group)

    if rank == 0:
        line["extra_workloads"] = extra_lines
        if cabi is not None:
            line["cabi_multi"] = cabi
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    be.close()


if __name__ == "__main__":
    main()

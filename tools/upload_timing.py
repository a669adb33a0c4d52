#!/usr/bin/env python
"""Times ptb_upload_scene on the synthetic scene for the BVH variants (host flatten + device LBVH + host eight-wide collapse)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import path_tracer_rust_b200 as P
import bench
path, base = bench.resolve_scene("synthetic")
t0 = time.perf_counter(); sc = P.Scene.load(path, base_dir=base); print(f"scene load {time.perf_counter()-t0:.2f} s", flush=True)
for opts in ({"bvh_wide": 0}, {"bvh_wide": 1, "bvh_wide_sah": 0}, {"bvh_wide": 1, "bvh_wide_sah": 2}):
    be = P.Backend(0)
    for k, v in opts.items():
        be.set_option(k, v)
    for rep in range(3):
        t0 = time.perf_counter(); be.upload_scene(sc); dt = time.perf_counter() - t0
        st = be.stats()
        print(opts, f"upload_scene wall {dt*1e3:.0f} ms (stats: upload {st['upload_ms']:.0f} ms, bvh_build {st['bvh_build_ms']:.0f} ms)", flush=True)
    be.close()
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))

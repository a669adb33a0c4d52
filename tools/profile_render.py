#!/usr/bin/env python
"""Small fixed render for ncu captures: python tools/profile_render.py [scene] [W] [H] [spp] [repeats]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import path_tracer_rust_b200 as P
import path_tracer_rust_b200.api as A

scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 32
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
be = P.Backend(0)
for kv in os.environ.get('PTB_OPTS', '').split(','):
    if '=' in kv:
        be.set_option(kv.split('=')[0], float(kv.split('=')[1]))
if scene == "synthetic":   # the BASELINE C5 scene (1.31 M triangles + 10 k spheres), generated like bench.py does
    import bench
    path, base = bench.resolve_scene("synthetic")
    be.upload_scene(P.Scene.load(path, base_dir=base))
else:
    be.upload_scene(P.Scene.load(scene))
for i in range(reps):
    t0 = time.perf_counter()
    be.render(W, H, spp, seed=i, out_kind=A.PTB_OUT_SUM)
    dt = time.perf_counter() - t0
    st = be.stats()
    print(f"{scene} {W}x{H}x{spp}: kernel {st['render_ms']:.2f} ms, wall {dt*1e3:.1f} ms, {st['samples']/st['render_ms']*1e-3:.1f} Mpaths/s, "
          f"{st['segments']/st['render_ms']*1e-3:.1f} Mseg/s, launches {st['kernel_launches']}", flush=True)
be.close()

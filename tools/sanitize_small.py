#!/usr/bin/env python
"""Tiny renders of every code path for compute-sanitizer (memcheck): megakernel, wavefront, BVH build + traversal, intersect,
peer reduce.  python tools/sanitize_small.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import path_tracer_rust_b200 as P
import path_tracer_rust_b200.api as A

spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tools", "make_synthetic_scene.py"))
mk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mk)
syn = mk.make_synthetic("/tmp/ptb_sanitize", level=3, n_spheres=120, scale=4.0, seed=3)
be = P.Backend(0)
assert be.selftest() == 0
rng = np.random.default_rng(0)
for scene, base in (("cornell", None), ("mesh", None), (syn, "/tmp/ptb_sanitize")):
    sc = P.Scene.load(scene, base_dir=base)
    for integ in (1, 2):
        be.set_option("integrator", integ)
        be.set_option("wavefront_paths", 3000)
        be.upload_scene(sc)
        img = be.render(37, 23, 5, seed=1, out_kind=A.PTB_OUT_SUM)
        o = rng.uniform(-3, 3, (5000, 3)).astype(np.float32)
        d = rng.normal(size=(5000, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        be.intersect(np.concatenate([o, d], 1))
        be.primary_hits(33, 17)
        print(sc.id, integ, float(img.sum()), be.stats()["segments"], flush=True)
be.set_option("integrator", 0)
be.upload_scene(P.Scene.load("cornell"))
fr = P.PeerMemoryFrame(be, 33, 21, seed=2, rank=0, world_size=1)
print("peer frame", float(fr.render(3).sum()))
fr.close()
be.close()
print("SANITIZE_RUN_OK")

#!/usr/bin/env python
"""Worst-case rays for the BVH padding (needs a GPU): 2 M rays grazing mctri.off triangles at |det| ~ 1e-4, closest hits of a BVH
built with 0 %, 1 %, 10 % and 100 % of the analytic padding against the brute-force scan.  python tools/pad_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, ctypes as C
import os
os.environ.setdefault('PTB_LIBRARY', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'path_tracer_rust_b200', 'libptb_exp.so'))  # bvh_pad_scale_UNSAFE only exists in the experiments build
import path_tracer_rust_b200 as P
f32=np.float32
rng = np.random.default_rng(2024)
sc = P.Scene.load('mesh')
d = sc._desc.contents
o = d.objects[0]
tris = np.frombuffer(C.string_at(d.triangles, 36 * sc.n_triangles), f32).reshape(-1, 3, 3)[o.tri_begin:o.tri_begin + o.tri_count] + np.array(list(o.position), f32)
n = 2_000_000
pick = rng.integers(0, len(tris), n)
A, E1, E2 = tris[pick, 0], tris[pick, 1] - tris[pick, 0], tris[pick, 2] - tris[pick, 0]
N = np.cross(E1, E2); area2 = np.linalg.norm(N, axis=1, keepdims=True); nrm = N / area2
u = rng.choice([0.0, 1.0, 0.5], n) + rng.normal(scale=0.02, size=n)
v = rng.uniform(-0.02, 1.02, n) * (1 - np.clip(u, 0, 1))
Pnt = A + E1 * u[:, None] + E2 * v[:, None]
phi = rng.uniform(0, 2 * np.pi, n)
e1n = E1 / np.linalg.norm(E1, axis=1, keepdims=True)
inpl = e1n * np.cos(phi)[:, None] + np.cross(nrm, e1n) * np.sin(phi)[:, None]
k = rng.uniform(0.5, 20.0, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
dirs = inpl + nrm * (k * 1e-4 / area2); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
dist = rng.uniform(0.05, 12.0, (n, 1))
rays = np.concatenate([Pnt - dirs * dist, dirs], 1).astype(f32)
bf = P.Backend(0); bf.set_option("bvh_min_tris", 1e18); bf.upload_scene(sc); ref = bf.intersect(rays)
for ps in (1.0, 0.1, 0.01, 0.0):
    b = P.Backend(0); b.set_option("bvh_pad_scale_UNSAFE", ps); b.upload_scene(sc); got = b.intersect(rays)
    mism = int((got[0] != ref[0]).sum() + ((got[0]==ref[0]) & (got[2].view(np.uint32) != ref[2].view(np.uint32))).sum())
    print(f"pad_scale {ps}: mismatches vs brute force: {mism} of {n}", flush=True)
    b.close()

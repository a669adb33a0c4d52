#!/usr/bin/env python
"""Builds the full-size synthetic scene (level 8 icosphere + 10k spheres), uploads it, reports build/traversal numbers and
checks a crop of primary hits against the brute-force oracle.  python tools/synthetic_check.py [level] [spheres] [W] [H] [spp]"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import os
os.environ.setdefault('PTB_LIBRARY', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'path_tracer_rust_b200', 'libptb_exp.so'))  # bvh_pad_scale_UNSAFE only exists in the experiments build
import path_tracer_rust_b200 as P
import path_tracer_rust_b200.api as A

spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tools", "make_synthetic_scene.py"))
mk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mk)

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nsph = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
H = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
spp = int(sys.argv[5]) if len(sys.argv) > 5 else 8
out = os.environ.get("SYN_DIR", "/tmp/ptb_synthetic")
t0 = time.perf_counter()
path = mk.make_synthetic(out, level=level, n_spheres=nsph, scale=4.0)
t1 = time.perf_counter()
sc = P.Scene.load(path, base_dir=out)
t2 = time.perf_counter()
be = P.Backend(0)
if os.environ.get('PAD_SCALE'):
    be.set_option('bvh_pad_scale_UNSAFE', float(os.environ['PAD_SCALE']))
for kv in os.environ.get('PTB_OPTS', '').split(','):
    if '=' in kv:
        be.set_option(kv.split('=')[0], float(kv.split('=')[1]))
be.upload_scene(sc)
t3 = time.perf_counter()
st = be.stats()
print(f"generate {t1-t0:.2f}s, parse {t2-t1:.2f}s, upload+build {t3-t2:.2f}s (bvh build {st['bvh_build_ms']:.2f} ms), "
      f"objects {sc.n_objects}, triangles {sc.n_triangles}, bvh nodes {st['n_bvh_nodes']}, bvh tris {st['n_bvh_triangles']}, "
      f"bvh spheres {st['n_bvh_spheres']}, loose objs {st['n_loose_objects']}", flush=True)
for i in range(2):
    be.render(W, H, spp, seed=i, out_kind=A.PTB_OUT_SUM)
    s = be.stats()
    print(f"render {W}x{H}x{spp}: {s['render_ms']:.1f} ms, {s['samples']/s['render_ms']*1e-3:.1f} Mpaths/s, "
          f"{s['segments']/s['render_ms']*1e-3:.1f} Mseg/s, {s['segments']/s['samples']:.2f} seg/sample, "
          f"{s['bvh_nodes_visited']/max(s['segments'],1):.1f} nodes/seg, {s['bvh_prims_tested']/max(s['segments'],1):.1f} prims/seg", flush=True)
if os.environ.get("SYN_ORACLE", "1") == "1":
    import oracle_lib as O
    osc = O.OracleScene(path, out)
    w, h = 64, 40
    g = be.primary_hits(w, h)
    t = time.perf_counter()
    o = osc.primary_hits(w, h)
    print(f"oracle brute-force primary hits {w}x{h}: {time.perf_counter()-t:.1f}s")
    ok = all(np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b) for a, b in zip(g, o))
    print("primary hits bit-exact vs oracle:", ok, "mesh pixels:", int((g[1] >= 0).sum()), "sphere pixels:", int(((g[0] >= 0) & (g[1] < 0)).sum()))
    assert ok
be.close()

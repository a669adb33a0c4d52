#!/usr/bin/env python
"""Synthetic benchmark scene (BASELINE.json configs[4]): a displaced icosphere OFF mesh + N spheres inside the Cornell walls.

Writes the reference's own formats: `<out>/meshes/icosphere_l<level>.off` (OFF, triangles only, so load_off.rs accepts it)
and `<out>/scenes/<id>.json` (serde layout of src/render/mod.rs:85-90).  Everything is procedural and seeded:
  - mesh: icosahedron subdivided `level` times (20*4^level triangles; level 8 = 1 310 720), vertices pushed radially by a
    fixed sum of sinusoids ("value noise"), unit radius in the file, scaled by `scale` through MeshFile.scale; centred on
    its local origin so the reference's quirky gate sphere (centre = min + max/2, mod.rs:478-482) still contains it;
  - spheres: centres uniform in the box, radii U[0.02, 0.08]*S, 70 % diffuse / 15 % specular / 15 % refract;
  - walls + ceiling light + camera: those of scenes/cornell.json, all lengths multiplied by S (focal length, sensor size and
    aspect ratio unchanged).  S matters because the reference rejects triangles with |det| < 1e-4 in ABSOLUTE units
    (mod.rs:571): a million-triangle mesh is only visible if the world is large enough (hard parts).
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def icosphere(level: int):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e.sort(axis=1)
        key = e[:, 0] * (len(v) + 1) + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // (len(v) + 1), uniq % (len(v) + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(f)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return v, f


def displace(v: np.ndarray, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    r = np.ones(len(v))
    for octave in range(4):
        k = rng.normal(size=3) * (2.0 ** octave) * 2.5
        ph = rng.uniform(0, 2 * np.pi)
        r += (0.12 / (1.6 ** octave)) * np.sin(v @ k + ph)
    r /= r.max()
    return v * r[:, None]


def write_off(path: str, v: np.ndarray, f: np.ndarray):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("OFF\n#\n#  procedurally displaced icosphere (tools/make_synthetic_scene.py)\n#\n")
        fh.write(f"{len(v)} {len(f)} 0\n")
        np.savetxt(fh, v.astype(np.float32), fmt="%.7g")
        np.savetxt(fh, np.concatenate([np.full((len(f), 1), 3, np.int64), f], 1), fmt="%d")


def make_synthetic(out_dir: str, level: int = 8, n_spheres: int = 10000, scale: float = 10.0, mesh_radius: float | None = None,
                   seed: int = 1, scene_id: str | None = None) -> str:
    S = float(scale)
    scene_id = scene_id or f"synthetic_l{level}_s{n_spheres}_x{S:g}"
    off_rel = f"meshes/icosphere_l{level}.off"
    off_path = os.path.join(out_dir, off_rel)
    if not os.path.exists(off_path):
        v, f = icosphere(level)
        write_off(off_path, displace(v, 0xB200), f)
    base = json.load(open(os.path.join(ROOT, "scenes", "cornell.json")))
    objects = []
    R = float(mesh_radius) if mesh_radius is not None else 1.6 * S
    objects.append({"type_": {"MeshFile": {"path": off_rel, "scale": R}}, "position": [0.0, -2.0 * S + R * 1.02, -1.0 * S],
                    "material": {"color": [0.75, 0.6, 0.3], "emmission": [0.0, 0.0, 0.0], "reflect_type": "Diffuse"}})
    rng = np.random.default_rng(seed)
    lo = np.array([-2.6, -2.0, -8.8]) * S
    hi = np.array([2.6, 1.9, 7.0]) * S
    for _ in range(n_spheres):
        c = rng.uniform(lo, hi)
        rad = float(rng.uniform(0.02, 0.08) * S)
        u = rng.uniform()
        kind = "Diffuse" if u < 0.70 else ("Specular" if u < 0.85 else "Refract")
        col = rng.uniform(0.3, 0.95, 3) if kind == "Diffuse" else np.array([0.999, 0.999, 0.999])
        objects.append({"type_": {"Sphere": {"radius": rad}}, "position": [float(x) for x in c],
                        "material": {"color": [float(x) for x in col], "emmission": [0.0, 0.0, 0.0], "reflect_type": kind}})
    for o in base["objects"]:
        if "Mesh" not in o["type_"]:
            continue
        m = o["type_"]["Mesh"]

        def sc3(p):
            return [float(np.float32(x) * np.float32(S)) for x in p]

        tris = [{k: sc3(t[k]) for k in ("a", "b", "c")} for t in m["triangles"]]
        bb = [{k: sc3(t[k]) for k in ("a", "b", "c")} for t in m["bounding_box"]]
        bs = {"position": sc3(m["bounding_sphere"]["position"]), "radius": float(np.float32(m["bounding_sphere"]["radius"]) * np.float32(S))}
        objects.append({"type_": {"Mesh": {"triangles": tris, "bounding_sphere": bs, "bounding_box": bb}},
                        "position": sc3(o["position"]), "material": o["material"]})
    cam = dict(base["camera"])
    cam["position"] = [float(np.float32(x) * np.float32(S)) for x in cam["position"]]
    cam.pop("updating_direction", None)
    scene = {"id": scene_id, "objects": objects, "camera": cam}
    path = os.path.join(out_dir, "scenes", f"{scene_id}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(scene, fh)
    return path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--spheres", type=int, default=10000)
    ap.add_argument("--scale", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    print(make_synthetic(a.out_dir, a.level, a.spheres, a.scale, seed=a.seed))

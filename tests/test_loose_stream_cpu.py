"""CPU check of the flattened scene (no GPU): walk the shared-memory object stream that ptb_upload_scene builds
(ptb_flatten_loose, layout in path_tracer_rust_b200/csrc/pt_device.cuh: DScene) with a numpy fp32 restatement of the device's
lock-step scan (closest_hit_loose / finish_hit in pt_scene_dev.cuh) and compare object, triangle, t, point and normal with the
oracle bit for bit.  numpy evaluates every fp32 operation separately (no FMA), like the device code built with --fmad=false.

This pins the host half of the product path -- pre-translated triangles, pair interleaving, host-side unit normals, scan order,
skip distances, end marker, the vote-free flag of wall quads -- on machines without a GPU; the kernels themselves are covered by
the `-m gpu` tests.
"""
import os

import numpy as np
import pytest

import oracle_lib as O

f32 = np.float32
KIND_SPHERE, KIND_END = 0, 1
EPS = f32(1e-4)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def dot(ax, ay, az, bx, by, bz):
    return (ax * bx + ay * by) + az * bz          # glam / V3 dot: (x x' + y y') + z z'


def sphere_t(c, r2, o, d):
    """sphere_t of pt_device.cuh (mod.rs:412-438): t, or -1 for a miss."""
    opx, opy, opz = c[0] - o[:, 0], c[1] - o[:, 1], c[2] - o[:, 2]
    b = dot(opx, opy, opz, d[:, 0], d[:, 1], d[:, 2])
    det = b * b - dot(opx, opy, opz, opx, opy, opz) + r2
    with np.errstate(invalid="ignore"):
        s = np.sqrt(det)
        t0, t1 = b - s, b + s
        t = np.where(t0 >= EPS, t0, np.where(t1 >= EPS, t1, f32(-1.0)))
        return np.where(det < 0, f32(-1.0), t).astype(f32)


def gate_pass(c, r2, o, d):
    """intersect_sphere(..).is_some() (mod.rs:267-272), evaluated exactly (the device's shortcuts are tested separately)."""
    return sphere_t(c, r2, o, d) >= 0


def triangle_hit(a, e1, e2, o, d):
    """triangle_hit of pt_device.cuh (mod.rs:560-593): (hit mask, t)."""
    px = d[:, 1] * e2[2] - e2[1] * d[:, 2]
    py = d[:, 2] * e2[0] - e2[2] * d[:, 0]
    pz = d[:, 0] * e2[1] - e2[0] * d[:, 1]
    det = dot(e1[0], e1[1], e1[2], px, py, pz)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = f32(1.0) / det
        tx, ty, tz = o[:, 0] - a[0], o[:, 1] - a[1], o[:, 2] - a[2]
        u = dot(tx, ty, tz, px, py, pz) * inv
        qx = ty * e1[2] - e1[1] * tz
        qy = tz * e1[0] - e1[2] * tx
        qz = tx * e1[1] - e1[0] * ty
        v = dot(d[:, 0], d[:, 1], d[:, 2], qx, qy, qz) * inv
        t = dot(e2[0], e2[1], e2[2], qx, qy, qz) * inv
        hit = ~(np.abs(det) < EPS) & ~(u < 0) & ~(u > 1) & ~(v < 0) & ~((u + v) > 1) & ~(t <= 0)
    return hit, t.astype(f32)


def walk_stream(stream, tris, rays):
    """closest_hit_loose + finish_hit over all rays at once; returns obj, tri, t, point, normal like ptb_intersect."""
    si = stream.view(np.int32)
    ti = tris.view(np.int32)
    o, d = rays[:, :3].copy(), rays[:, 3:].copy()
    n = len(rays)
    best_t = np.full(n, np.inf, f32)
    best_sphere = np.full(n, -1, np.int64)   # stream offset of the winning sphere record, or -1
    best_tri = np.full(n, -1, np.int64)      # index of the winning triangle record, or -1
    rec, seen_end = 0, False
    while rec < len(stream):
        hdr, kind = stream[rec], si[rec + 1, 0]
        if kind == KIND_SPHERE:
            assert si[rec + 1, 2] == rec, "a sphere record must carry its own offset"
            t = sphere_t(hdr[:3], hdr[3], o, d)
            upd = (t >= 0) & (t < best_t)
            best_t = np.where(upd, t, best_t)
            best_sphere = np.where(upd, rec, best_sphere)
            best_tri = np.where(upd, -1, best_tri)
            rec += 2
            continue
        if kind == KIND_END:
            seen_end = True
            break
        k_begin, n_tri, skip = si[rec + 1, 1], si[rec + 1, 2], si[rec + 1, 3]
        n_pairs = 1 if n_tri == 0 else n_tri // 2          # n_tri == 0 marks a vote-free single pair
        assert skip == 2 + 5 * n_pairs
        passed = gate_pass(hdr[:3], hdr[3], o, d)
        for p in range(n_pairs):
            q = stream[rec + 2 + 5 * p: rec + 7 + 5 * p].reshape(-1)      # 20 floats: "|2" interleaved, see triangle_pair_hit
            for w in (0, 1):
                a, e1, e2 = q[0 + w:6 + w:2], q[6 + w:12 + w:2], q[12 + w:18 + w:2]
                hit, t = triangle_hit(a, e1, e2, o, d)
                upd = passed & hit & (t < best_t)
                best_t = np.where(upd, t, best_t)
                best_tri = np.where(upd, k_begin + 2 * p + w, best_tri)
                best_sphere = np.where(upd, -1, best_sphere)
        rec += skip
    assert seen_end, "the stream must end with an end marker"
    obj = np.full(n, -1, np.int32)
    tri = np.full(n, -1, np.int32)
    hit_any = (best_sphere >= 0) | (best_tri >= 0)
    t_out = np.where(hit_any, best_t, f32(0)).astype(f32)
    x = (o + d * best_t[:, None]).astype(f32)
    normal = np.zeros((n, 3), f32)
    sp = best_sphere >= 0
    if sp.any():
        r = best_sphere[sp]
        obj[sp] = si[r + 1, 3]
        v = x[sp] - stream[r, :3]
        length = np.sqrt(dot(v[:, 0], v[:, 1], v[:, 2], v[:, 0], v[:, 1], v[:, 2]))
        normal[sp] = v * (f32(1.0) / length)[:, None]                    # glam normalize: v * (1 / length)
    tr = best_tri >= 0
    if tr.any():
        k = best_tri[tr]
        obj[tr] = ti[2 * k, 3]
        tri[tr] = ti[2 * k + 1, 0]
        normal[tr] = tris[2 * k, :3]
    x[~hit_any] = 0
    return obj, tri, t_out, x, normal


def _rays(osc, rng, n):
    o = rng.uniform(-3, 3, (n, 3)).astype(f32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    rays = np.concatenate([o, d], 1)
    first = osc.intersect(rays)
    hit = first[0] >= 0
    d2 = rng.normal(size=(int(hit.sum()), 3))
    d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(f32)
    return np.concatenate([rays, np.concatenate([first[3][hit], d2], 1)], 0)   # + rays that start on surfaces


@pytest.mark.parametrize("scene,n", [("cornell", 60_000), ("three-spheres", 20_000), ("cartesian", 20_000), ("mesh", 6_000)])
@pytest.mark.parametrize("quad_min_ratio", [0.125, 0.0, 1e9])
def test_flattened_stream_matches_oracle(scene, n, quad_min_ratio):
    import path_tracer_rust_b200 as P
    sc = P.Scene.load(scene)
    osc = O.OracleScene(os.path.join(ROOT, "scenes", f"{scene}.json"))
    stream, tris = sc.flatten_loose(quad_min_ratio)
    rays = _rays(osc, np.random.default_rng(5), n)
    want = osc.intersect(rays)
    got = walk_stream(stream, tris, rays)
    names = ("obj", "tri", "t", "point", "normal")
    for name, a, b in zip(names, got, want):
        if a.dtype == f32:
            same = bits(a) == bits(b)
        else:
            same = a == b
        assert same.all(), (scene, name, int((~same).sum()), a[~same][:3], b[~same][:3])
    assert (want[0] >= 0).sum() > 500          # (open scenes: most random rays miss)


def test_empty_mesh_is_dropped_from_the_stream(tmp_path):
    """a mesh with zero triangles (inline `triangles: []`) used to be written with n_tri = 0, the marker of a vote-free
    wall quad, and desynchronised the walk.  The reference renders it as no hit (mod.rs:558 loops zero times)."""
    import json
    import path_tracer_rust_b200 as P
    base = json.load(open(os.path.join(ROOT, "scenes", "cornell.json")))
    empty = {"type_": {"Mesh": {"triangles": [], "bounding_sphere": {"position": [0.0, 0.0, 0.0], "radius": 50.0},
                                "bounding_box": base["objects"][-1]["type_"]["Mesh"]["bounding_box"]}},
             "position": [0.0, 0.0, 0.0], "material": base["objects"][0]["material"]}
    objs = list(base["objects"])
    objs.insert(2, empty)          # between spheres
    objs.insert(len(objs) - 2, dict(empty, position=[1.0, 0.0, 0.0]))   # between walls
    objs.append(dict(empty, position=[2.0, 0.0, 0.0]))                  # first in the reference's (reverse) scan order
    path = tmp_path / "empty_mesh.json"
    json.dump({"id": "empty_mesh", "objects": objs, "camera": base["camera"]}, open(path, "w"))
    sc = P.Scene.load(str(path), base_dir=ROOT)
    osc = O.OracleScene(str(path), ROOT)
    for ratio in (0.125, 0.0, 1e9):
        stream, tris = sc.flatten_loose(ratio)
        rays = _rays(osc, np.random.default_rng(11), 30_000)
        want = osc.intersect(rays)
        got = walk_stream(stream, tris, rays)
        for name, a, b in zip(("obj", "tri", "t", "point", "normal"), got, want):
            same = (bits(a) == bits(b)) if a.dtype == f32 else (a == b)
            assert same.all(), (ratio, name, int((~same).sum()))
        assert (want[0] >= 0).sum() > 500


def test_quad_flag_follows_the_ratio():
    import path_tracer_rust_b200 as P
    sc = P.Scene.load("cornell")

    def flags(ratio):
        stream, _ = sc.flatten_loose(ratio)
        si, rec, out = stream.view(np.int32), 0, []
        while si[rec + 1, 0] != KIND_END:
            if si[rec + 1, 0] == KIND_SPHERE:
                rec += 2
            else:
                out.append(int(si[rec + 1, 2]))
                rec += si[rec + 1, 3]
        return out

    assert flags(0.125) == [0] * 7          # the seven walls of the box: vote-free
    assert flags(1e9) == [2] * 7            # never
    assert flags(0.0) == [0] * 7


def test_flatten_rejects_bad_descriptions():
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    cam = {"position": [0, 0, 5], "direction": [0, 0, -1]}
    bad = P.Scene.from_arrays([{"kind": "mesh", "position": [0, 0, 0], "color": [1, 1, 1], "emission": [0, 0, 0], "tri_begin": 0,
                                "tri_count": 3, "bs_radius": 1.0}], np.zeros((1, 9), f32), cam)
    with pytest.raises(A.BackendError):
        bad.flatten_loose()

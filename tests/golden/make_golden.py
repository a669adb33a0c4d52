#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the CPU ORACLE (the Rust reference cannot run in this image; the oracle is pinned to
the reference's own known answers by tests/test_oracle_kat.py).  The fixture freezes the oracle's outputs so that
  - a later change of the oracle that alters any result is caught on the CPU (tests/test_golden.py::test_oracle_reproduces_golden),
  - the CUDA path can be checked against committed vectors without running the oracle (…::test_cuda_reproduces_golden, -m gpu).
Contents per scene: deterministic primary hits (48x32), and the lock-step sum framebuffer + segment count (24x16 x 4 spp, seed 7,
Philox / det sin-cos / forward accumulation).  Plus the Philox known answers, a sin/cos table and the gamma table.

    python tests/golden/make_golden.py        # rewrites golden_v1.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O

SCENES = ["cornell", "mesh", "single-sphere", "two-spheres", "three-spheres", "cartesian"]
PW, PH = 48, 32
W, H, SPP, SEED = 24, 16, 4, 7


def build():
    out = {}
    for sid in SCENES:
        sc = O.OracleScene(os.path.join(ROOT, "scenes", f"{sid}.json"))
        obj, tri, t = sc.primary_hits(PW, PH)
        fb, st = sc.render_sum(W, H, SPP, seed=SEED)
        key = sid.replace("-", "_")
        out[f"{key}__obj"], out[f"{key}__tri"], out[f"{key}__t_bits"] = obj, tri, t.view(np.uint32)
        out[f"{key}__fb_bits"], out[f"{key}__segments"] = fb.view(np.uint32), np.array([st[0]], np.uint64)
    out["philox_ctr"] = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [1, 2, 3, 4]], np.uint32)
    out["philox_key"] = np.array([[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0], [5, 6]], np.uint32)
    out["philox_out"] = np.stack([O.philox(c, k) for c, k in zip(out["philox_ctr"], out["philox_key"])])
    xs = np.linspace(0, 6.2831855, 257).astype(np.float32)
    sc_ = np.array([O.sincos_det(float(x)) for x in xs], np.float32)
    out["sincos_x_bits"], out["sincos_bits"] = xs.view(np.uint32), sc_.view(np.uint32)
    gx = np.linspace(-0.25, 1.25, 301).astype(np.float32)
    out["gamma_x_bits"], out["gamma_u8"] = gx.view(np.uint32), np.array([O.gamma_u8(float(x)) for x in gx], np.uint8)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **build())
    print("wrote", os.path.join(HERE, "golden_v1.npz"))

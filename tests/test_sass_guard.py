"""Build-time guard for the parity-critical code generation (CPU only: reads the SASS of the built libptb.so).

Parity with the reference needs every fp32 add and multiply of the intersection and shading code to round separately
(rustc never contracts to FMA).  ptb_create probes that on the GPU; this test catches a wrong build flag or a
compiler that starts fusing the packed operations already on the build machine: no packed FMA anywhere, the packed multiplies
where they are meant to be, and only the handful of scalar FMAs that the correctly rounded reciprocal / square root / division
sequences contain.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "path_tracer_rust_b200", "libptb.so")


def _sass_by_kernel():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            ops = line.split("*/", 1)[1].split()
            if ops and ops[0].startswith("@"):      # predicate prefix
                ops = ops[1:]
            if ops:
                kernels[name].append(ops[0].rstrip(";"))
    return kernels


def _count(ops, prefix):
    return sum(1 for o in ops if o == prefix or o.startswith(prefix + "."))


def test_no_fused_packed_arithmetic_and_few_scalar_fmas(ensure_built):
    kernels = _sass_by_kernel()
    ours = {k: v for k, v in kernels.items() if "ptb" in k and any(t in k for t in ("k_render", "k_intersect", "k_wf_", "k_gather_prims"))}
    assert len(ours) >= 8, sorted(kernels)
    for name, ops in ours.items():
        assert _count(ops, "FFMA2") == 0, name                      # a packed multiply-add would fuse two roundings
        assert _count(ops, "HFMA2.MMA") == 0, name
    render = next(v for k, v in ours.items() if "k_renderILb0" in k)
    inter = next(v for k, v in ours.items() if "k_intersectILb0" in k)
    for ops, what in ((render, "k_render<false>"), (inter, "k_intersect<false>")):
        # 27 packed multiplies per inlined triangle-pair test: the looped one and the vote-free wall-quad one
        assert _count(ops, "FMUL2") == 54, (what, _count(ops, "FMUL2"))
        assert _count(ops, "FADD2") == 6, (what, _count(ops, "FADD2"))        # tvec = o - a only (no product among its inputs)
    # scalar FMAs exist only inside the rn reciprocal / square-root / division refinement sequences and in Philox-free
    # address arithmetic; a build without --fmad=false has hundreds
    assert _count(inter, "FFMA") <= 60, _count(inter, "FFMA")      # 51 as built; 76 when built with --fmad=true
    assert _count(render, "FFMA") <= 100, _count(render, "FFMA")   # 76 as built; 152 when built with --fmad=true
    assert _count(inter, "FADD") >= 100 and _count(inter, "FMUL") >= 40

"""CPU check of the compressed eight-wide BVH (pt_bvh8.h): the host builder (pt_bvh8_build.cpp) and the SAME node test the CUDA trace
kernel uses (pt_bvh8.cuh is host + device) are compiled with g++ and walked against an exact double-precision box test: no primitive
whose box a ray segment touches may be culled, for dense / scattered / flat / huge boxes, axis-parallel rays and tiny sets."""
import os
import subprocess

from conftest import ROOT


def test_bvh8_collapse_and_traversal_cull_nothing(tmp_path):
    exe = str(tmp_path / "bvh8_check")
    src = os.path.join(ROOT, "tests", "cpp", "bvh8_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", exe, src,
                    os.path.join(ROOT, "path_tracer_rust_b200", "csrc", "pt_bvh8_build.cpp")], check=True)
    for n, rays in ((1, 50), (2, 200), (3, 200), (4, 300), (9, 500), (37, 2000), (5000, 4000), (20000, 6000), (60000, 300)):
        r = subprocess.run([exe, str(n), str(rays)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "missed 0 " in r.stdout, (n, r.stdout, r.stderr)
    # negative control: the same check must FIND lost hits when the builder leaves out the one-step margin that covers the
    # traversal's decode error (a fifth of the rays graze box corners and edges) -- i.e. the check has teeth and the margin is needed
    exe0 = str(tmp_path / "bvh8_check_nomargin")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-DPTB_BVH8_NO_MARGIN", "-o", exe0, src,
                    os.path.join(ROOT, "path_tracer_rust_b200", "csrc", "pt_bvh8_build.cpp")], check=True)
    r = subprocess.run([exe0, "20000", "6000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "missed 0 " not in r.stdout, r.stdout

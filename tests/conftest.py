import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCENES = ["cornell", "mesh", "single-sphere", "two-spheres", "three-spheres", "cartesian"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def scene_path(scene_id: str) -> str:
    return os.path.join(ROOT, "scenes", f"{scene_id}.json")


CAMERA = {"position": [0.0, 0.0, 0.0], "direction": [0.0, 0.0, -1.0], "focal_length": 0.035,
          "sensor_width": 0.036, "aspect_ratio": 1.5}


def sphere_obj(pos, radius=1.0, color=(1.0, 0.0, 0.0), emission=(0.0, 0.0, 0.0), refl="Diffuse"):
    return {"type_": {"Sphere": {"radius": radius}}, "position": list(pos),
            "material": {"color": list(color), "emmission": list(emission), "reflect_type": refl}}


def write_scene(path, objects, scene_id="kat", camera=None):
    with open(path, "w") as f:
        json.dump({"id": scene_id, "objects": objects, "camera": camera or CAMERA}, f)
    return str(path)


@pytest.fixture
def kat_scene(tmp_path):
    def make(objects, name="kat", camera=None):
        return write_scene(tmp_path / f"{name}.json", objects, name, camera)
    return make


def pytest_collection_modifyitems(config, items):
    """GPU tests skip (instead of erroring) on a machine without a CUDA device or without the built library."""
    n_dev = 0
    try:
        import path_tracer_rust_b200 as P
        n_dev = P.load_library().ptb_device_count()
    except Exception:
        n_dev = 0
    if n_dev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (or libptb.so not built)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def ensure_built():
    """libptb.so, the render CLI and the oracle are build products (git-ignored); build them if this checkout has none."""
    need = [os.path.join(ROOT, "path_tracer_rust_b200", "libptb.so"), os.path.join(ROOT, "path_tracer_rust_b200", "render"),
            os.path.join(ROOT, "oracle", "_build", "libpt_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()

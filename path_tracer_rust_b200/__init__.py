"""B200-native backend for the per-pixel Monte-Carlo radiance loop of filippo-orru/path-tracer-rust.

The product is libptb.so (hand-written CUDA for sm_100a behind the C ABI in include/ptb.h).  This package is the
host-side mirror of the reference's render interface (src/render/mod.rs) over that C ABI, used by the tests, the
bench and the multi-GPU driver.  There is no CPU fallback: importing works anywhere, computing needs the built
library and a B200.
"""
from .api import (Backend, BackendError, Image, RenderConfig, RenderDone, RenderUpdate, Resolution, Scene,
                  gamma_correction, library_path, load_library, render, to_int_with_gamma_correction)
from .distributed import PeerMemoryFrame, render_sharded

__all__ = ["Backend", "BackendError", "Image", "RenderConfig", "RenderDone", "RenderUpdate", "Resolution", "Scene",
           "gamma_correction", "library_path", "load_library", "render", "render_sharded", "PeerMemoryFrame", "to_int_with_gamma_correction"]

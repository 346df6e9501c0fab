"""sol_rs_b200 — B200 (sm_100a) ray-tracing hot path behind the sol-rs scene/ray API.

Host-side mirror of the reference's Rust API for this path (names follow /root/reference
src/scene, src/ray, examples/5-pathtrace.rs); all device work happens in libsolb.so (hand-written
CUDA, include/solb.h).  No CPU fallback: importing works anywhere, creating a Context needs a GPU.
"""
from . import _native  # noqa: F401
from ._native import SolbError  # noqa: F401
from .context import Context, Image2d  # noqa: F401
from . import scene, ray, util, multigpu, io  # noqa: F401

__all__ = ["Context", "Image2d", "SolbError", "scene", "ray", "util", "multigpu", "io"]

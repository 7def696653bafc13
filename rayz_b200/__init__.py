"""rayz_b200 — B200-native (sm_100a) path-tracing backend for jlucier/rayz's per-pixel sample loop.

The product is `lib/librayz_cuda.so` (hand-written CUDA behind the C ABI of include/rayz_cuda.h);
this package is the host-side mirror of the reference's Camera/MemPool/Tracer/Image API over it.
Importing the package does not load CUDA; the first backend call does, and fails loudly if the
library or a GPU is missing (there is no CPU fallback).
"""
from . import _abi as abi  # noqa: F401
from .host import (ASPECT_RATIO, Backend, Camera, Image, MemPool, Tracer, Xoshiro256, penultimate_scene,  # noqa: F401
                   random_bouncing, scene_struct)

__all__ = ["abi", "ASPECT_RATIO", "Backend", "Camera", "Image", "MemPool", "Tracer", "Xoshiro256", "penultimate_scene", "random_bouncing",
           "scene_struct"]

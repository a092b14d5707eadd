"""velvet_b200 -- B200-native (sm_100a) XPBD cloth solver: a headless drop-in for the hot path of
vitalight/Velvet (VtClothSolverGPU::Simulate + SpatialHashGPU::Hash + VtBuffer) behind a C ABI.

The package is a thin ctypes host over velvet_b200/lib/libvelvet_b200.so (include/velvet_b200.h).  There is
no CPU implementation here: loading fails loudly when the CUDA library has not been built.
"""
from ._capi import (BUFFER_IDS, COLLIDER_CUBE, COLLIDER_PLANE, COLLIDER_SPHERE, EXPORTED_SYMBOLS, LIB_PATH,
                    ITERATE_AUTO, ITERATE_GRID, ITERATE_TILES, MATH_EXACT, MATH_FAST, PIPELINE_FUSED, PIPELINE_SEAM, VelvetError, VtHashParams, VtSDFCollider, VtSimParams, load)
from .solver import (GenerateClothMesh, MakeCollider, SpatialHashGPU, TransformMatrix, VtClothObjectGPU,
                     VtClothSolverGPU, build_scene, default_params, sphere_plane_colliders)
from . import seam

__all__ = [
    "BUFFER_IDS", "COLLIDER_CUBE", "COLLIDER_PLANE", "COLLIDER_SPHERE", "EXPORTED_SYMBOLS", "LIB_PATH",
    "ITERATE_AUTO", "ITERATE_GRID", "ITERATE_TILES", "MATH_EXACT", "MATH_FAST", "PIPELINE_FUSED", "PIPELINE_SEAM", "VelvetError", "VtHashParams", "VtSDFCollider", "VtSimParams", "load",
    "GenerateClothMesh", "MakeCollider", "SpatialHashGPU", "TransformMatrix", "VtClothObjectGPU",
    "VtClothSolverGPU", "build_scene", "default_params", "sphere_plane_colliders", "seam",
]

"""ctypes binding of include/velvet_b200.h.  No CPU fallback: importing without the built CUDA library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvelvet_b200.so")
if os.environ.get("VELVET_VARIANT"):  # experiments only, see build.py
    LIB_PATH = os.path.join(_HERE, "lib", f"libvelvet_b200_{os.environ['VELVET_VARIANT']}.so")


class VtSimParams(C.Structure):
    """Common.hpp L19-47 (80 bytes)."""
    _fields_ = [
        ("numSubsteps", C.c_int32), ("numIterations", C.c_int32), ("maxNumNeighbors", C.c_int32),
        ("maxSpeed", C.c_float), ("gravity", C.c_float * 3), ("bendCompliance", C.c_float),
        ("damping", C.c_float), ("relaxationFactor", C.c_float), ("longRangeStretchiness", C.c_float),
        ("collisionMargin", C.c_float), ("friction", C.c_float), ("enableSelfCollision", C.c_uint8),
        ("_pad", C.c_uint8 * 3), ("interleavedHash", C.c_int32), ("numParticles", C.c_uint32),
        ("particleDiameter", C.c_float), ("deltaTime", C.c_float), ("particleDiameterScalar", C.c_float),
        ("hashCellSizeScalar", C.c_float),
    ]


class VtSDFCollider(C.Structure):
    """VtClothSolverGPU.cuh L8-18 (196 bytes)."""
    _fields_ = [
        ("type", C.c_int32), ("position", C.c_float * 3), ("scale", C.c_float * 3), ("deltaTime", C.c_float),
        ("curTransform", C.c_float * 9), ("invCurTransform", C.c_float * 16), ("lastTransform", C.c_float * 16),
    ]


class VtHashParams(C.Structure):
    """SpatialHashGPU.cuh L7-15 (24 bytes)."""
    _fields_ = [
        ("numObjects", C.c_uint32), ("maxNumNeighbors", C.c_uint32), ("cellSpacing", C.c_float),
        ("cellSpacing2", C.c_float), ("tableSize", C.c_int32), ("particleDiameter2", C.c_float),
    ]


assert C.sizeof(VtSimParams) == 80 and C.sizeof(VtSDFCollider) == 196 and C.sizeof(VtHashParams) == 24

COLLIDER_SPHERE, COLLIDER_PLANE, COLLIDER_CUBE = 0, 1, 2
PIPELINE_FUSED, PIPELINE_SEAM = 0, 1
MATH_EXACT, MATH_FAST = 0, 1
ITERATE_AUTO, ITERATE_TILES, ITERATE_GRID = 0, 1, 2

BUFFER_IDS = {name: i for i, name in enumerate([
    "positions", "normals", "indices", "velocities", "predicted", "deltas", "deltaCounts", "invMasses",
    "stretchIndices", "stretchLengths", "bendIndices", "bendAngles", "attachParticleIDs", "attachSlotIDs",
    "attachDistances", "attachSlotPositions", "neighbors", "initialPositions", "particleHash", "particleIndex",
    "cellStart", "cellEnd", "sdfColliders"])}

# every symbol include/velvet_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "velvet_last_error", "velvet_version", "velvet_default_params",
    "velvet_SetSimulationParams", "velvet_InitializePositions", "velvet_PredictPositions", "velvet_SolveStretch",
    "velvet_SolveBending", "velvet_SolveAttachment", "velvet_ApplyDeltas", "velvet_CollideSDF",
    "velvet_CollideParticles", "velvet_Finalize", "velvet_ComputeNormal", "velvet_HashObjects", "velvet_SortPairs",
    "velvet_seam_set_stream", "velvet_selftest_division", "velvet_selftest_constraints", "velvet_device_synchronize", "velvet_alloc", "velvet_free", "velvet_copy",
    "velvet_solver_create", "velvet_solver_destroy", "velvet_solver_params", "velvet_solver_set_pipeline",
    "velvet_solver_set_math_mode", "velvet_solver_set_iterate_mode", "velvet_solver_iterate_kernel", "velvet_solver_set_tile_size", "velvet_solver_add_cloth", "velvet_solver_add_stretch",
    "velvet_solver_add_attach_slot", "velvet_solver_add_attach", "velvet_solver_add_bend",
    "velvet_solver_update_colliders", "velvet_make_collider", "velvet_solver_simulate", "velvet_solver_simulate_dt",
    "velvet_solver_synchronize", "velvet_solver_hash", "velvet_solver_hash_fused", "velvet_solver_buffer", "velvet_solver_download",
    "velvet_solver_upload", "velvet_solver_set_render_targets", "velvet_solver_sync_render_targets", "velvet_solver_set_hash_host_readable", "velvet_solver_check_nan", "velvet_solver_grab", "velvet_solver_drag", "velvet_solver_release", "velvet_solver_readback_async", "velvet_solver_readback_pipelined", "velvet_solver_readback_wait", "velvet_solver_stream",
    "velvet_solver_last_launch_count", "velvet_solver_simulate_timed", "velvet_generate_cloth_mesh",
    "velvet_transform_matrix", "velvet_cloth_object_start", "velvet_solver_add_cloth_instances", "velvet_solver_dd_setup", "velvet_solver_dd_info",
    "velvet_solver_dd_offsets", "velvet_solver_dd_prepare_stepped", "velvet_solver_dd_step", "velvet_dd_plan_grid", "velvet_plan_grid_tiles", "velvet_plan_grid_smem_wavefronts", "velvet_plan_grid_digest", "velvet_grid_plan_check", "velvet_dd_peer_blob_bytes",
    "velvet_solver_dd_peer_export", "velvet_solver_dd_peer_import", "velvet_solver_dd_peer_close", "velvet_solver_dd_simulate", "velvet_hash_create", "velvet_hash_destroy",
    "velvet_hash_set_initial_positions", "velvet_hash_hash", "velvet_hash_buffer",
]

_lib = None


class VelvetError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"velvet_b200 error {status}: {message}")
        self.status = status


def load():
    """Loads libvelvet_b200.so (built by velvet_b200.build / __graft_entry__.build()).  Raises if it is missing:
    the product has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first (python -m velvet_b200.build). "
            "velvet_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    v, i, u, f = C.c_void_p, C.c_int, C.c_uint, C.c_float
    L.velvet_last_error.restype = C.c_char_p
    L.velvet_solver_params.restype = C.POINTER(VtSimParams)
    L.velvet_solver_params.argtypes = [v]
    L.velvet_solver_stream.restype = v
    L.velvet_solver_stream.argtypes = [v]
    sig = {
        "velvet_default_params": [C.POINTER(VtSimParams)],
        "velvet_SetSimulationParams": [C.POINTER(VtSimParams)],
        "velvet_InitializePositions": [v, i, i, v],
        "velvet_PredictPositions": [v, v, v, f],
        "velvet_SolveStretch": [v, v, v, v, v, v, u],
        "velvet_SolveBending": [v, v, v, v, v, v, u, f],
        "velvet_SolveAttachment": [v, v, v, v, v, v, v, v, i],
        "velvet_ApplyDeltas": [v, v, v],
        "velvet_CollideSDF": [v, v, v, u, f],
        "velvet_CollideParticles": [v, v, v, v, v, v],
        "velvet_Finalize": [v, v, v, f],
        "velvet_ComputeNormal": [v, v, v, u],
        "velvet_HashObjects": [v, v, v, v, v, v, v, VtHashParams],
        "velvet_SortPairs": [v, v, u, i],
        "velvet_seam_set_stream": [v],
        "velvet_device_synchronize": [],
        "velvet_alloc": [C.POINTER(v), C.c_size_t],
        "velvet_free": [v],
        "velvet_copy": [v, v, C.c_size_t],
        "velvet_solver_create": [C.POINTER(v), i, C.POINTER(VtSimParams)],
        "velvet_solver_destroy": [v],
        "velvet_solver_set_pipeline": [v, i],
        "velvet_solver_set_tile_size": [v, i],
        "velvet_solver_set_math_mode": [v, i],
        "velvet_solver_set_iterate_mode": [v, i],
        "velvet_solver_iterate_kernel": [v, C.POINTER(i)],
        "velvet_solver_add_cloth": [v, v, i, v, i, v, f, C.POINTER(i)],
        "velvet_solver_add_stretch": [v, i, i, f],
        "velvet_solver_add_attach_slot": [v, v],
        "velvet_solver_add_attach": [v, i, i, f],
        "velvet_solver_add_bend": [v, u, u, u, u, f],
        "velvet_solver_update_colliders": [v, v, i],
        "velvet_make_collider": [i, v, v, v, v, f, C.POINTER(VtSDFCollider)],
        "velvet_solver_simulate": [v, i],
        "velvet_solver_simulate_dt": [v, f, i],
        "velvet_solver_synchronize": [v],
        "velvet_solver_hash": [v],
        "velvet_solver_hash_fused": [v],
        "velvet_solver_buffer": [v, i, C.POINTER(v), C.POINTER(C.c_size_t)],
        "velvet_solver_download": [v, i, v, C.c_size_t],
        "velvet_solver_upload": [v, i, v, C.c_size_t],
        "velvet_solver_readback_async": [v, v, v],
        "velvet_solver_readback_pipelined": [v, v, v, C.POINTER(i)],
        "velvet_solver_readback_wait": [v, i],
        "velvet_solver_last_launch_count": [v],
        "velvet_solver_simulate_timed": [v, C.POINTER(C.c_char_p), C.POINTER(f), i],
        "velvet_generate_cloth_mesh": [i, v, v],
        "velvet_transform_matrix": [v, v, v, v],
        "velvet_cloth_object_start": [v, i, v, v, v, v, i, C.POINTER(i)],
        "velvet_solver_add_cloth_instances": [v, i, v, v, v, i, v, i],
        "velvet_hash_create": [C.POINTER(v), f, i, f, i],
        "velvet_hash_destroy": [v],
        "velvet_hash_set_initial_positions": [v, v, C.c_size_t],
        "velvet_hash_hash": [v, v, C.c_size_t],
        "velvet_hash_buffer": [v, i, C.POINTER(v), C.POINTER(C.c_size_t)],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = L
    return L


def check(status: int) -> int:
    if status < 0:
        raise VelvetError(status, load().velvet_last_error().decode("utf-8", "replace"))
    return status

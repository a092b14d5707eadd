"""CPU-only: the C-ABI library loads and exports every symbol include/velvet_b200.h declares; host-side
validation works without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import velvet_b200 as vb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "velvet_b200.h")).read()
    declared = set(re.findall(r"VELVET_API\s+[\w\s\*]+?\b(velvet_\w+)\s*\(", header))
    assert len(declared) >= 50
    assert declared == set(vb.EXPORTED_SYMBOLS), declared ^ set(vb.EXPORTED_SYMBOLS)
    L = vb.load()
    for sym in declared:
        assert hasattr(L, sym), sym


def test_error_convention_without_gpu_or_with_bad_arguments():
    L = vb.load()
    assert L.velvet_version() >= 100
    # NULL params -> negative status + message, never exit()
    assert L.velvet_SetSimulationParams(None) == -1
    assert b"NULL" in L.velvet_last_error()
    assert L.velvet_default_params(None) == -1
    assert L.velvet_generate_cloth_mesh(0, None, None) == -1


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle: no import, include, link or symbol use (comments that
    merely name the oracle as the thing the kernels are checked against are fine)."""
    patterns = [r"^\s*(from|import)\s+oracle\b", r"#\s*include\s*[\"<][^\">]*oracle", r"\bo[12]_[a-z_]+\s*\(", r"libo[12]_",
                r"refcuda", r"CDLL\([^)]*oracle"]
    for dirpath, _, files in os.walk(os.path.join(ROOT, "velvet_b200")):
        if os.path.basename(dirpath).startswith("build"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in patterns:
                    assert not re.search(pat, text, re.M), (f, pat)
    # and the shared library has no dependency on the oracle libraries
    import subprocess
    needed = subprocess.run(["readelf", "-d", vb.LIB_PATH], capture_output=True, text=True).stdout
    assert "libo1" not in needed and "libo2" not in needed and "refcuda" not in needed


def test_tile_plan_record_order_avoids_shared_memory_bank_conflicts():
    """tile_plan.cpp emits each tile's constraint records in an order whose quarter-warps (8 consecutive records) touch 8
    different 16-byte bank groups per endpoint.  Host-only model of the constraint threads' position loads + slot stores:
    in constraint-id order ~2x the minimum number of wavefronts, in the emitted order within 25 % of it."""
    import ctypes as C
    from velvet_b200 import _capi
    L = _capi.load()
    L.velvet_plan_grid_smem_wavefronts.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    for R, tile in ((63, 256), (255, 256), (255, 128)):
        out = (C.c_ulonglong * 3)()
        assert L.velvet_plan_grid_smem_wavefronts(R, tile, out) == 0
        ideal, id_order, emitted = out[0], out[1], out[2]
        assert id_order >= 1.7 * ideal, (R, tile, list(out))
        assert emitted <= 1.25 * ideal, (R, tile, list(out))


def test_tile_plan_does_not_depend_on_the_number_of_builder_threads():
    """build_tile_plan distributes tiles over worker threads; every array of the plan must be identical to the
    single-threaded build (the kernels' results depend on the plan only through these arrays)."""
    import ctypes as C
    import os
    from velvet_b200 import _capi
    L = _capi.load()
    L.velvet_plan_grid_digest.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]

    def digest(R, tile, attach, threads):
        old = os.environ.get("VELVET_PLAN_THREADS")
        if threads is None:
            os.environ.pop("VELVET_PLAN_THREADS", None)
        else:
            os.environ["VELVET_PLAN_THREADS"] = str(threads)
        try:
            d = C.c_ulonglong()
            assert L.velvet_plan_grid_digest(R, tile, attach, C.byref(d)) == 0
            return d.value
        finally:
            if old is None:
                os.environ.pop("VELVET_PLAN_THREADS", None)
            else:
                os.environ["VELVET_PLAN_THREADS"] = old

    for R, tile, attach in ((255, 256, 0), (255, 128, 1), (300, 256, 1), (511, 256, 0)):
        one = digest(R, tile, attach, 1)
        assert digest(R, tile, attach, 3) == one and digest(R, tile, attach, 8) == one and digest(R, tile, attach, None) == one

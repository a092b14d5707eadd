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


def _grid_check(counts, si, sl, bi, ba, want_rest=True):
    import ctypes as C
    import numpy as np
    L = vb.load()
    L.velvet_grid_plan_check.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_int), C.POINTER(C.c_uint), C.c_void_p, C.c_char_p]
    counts = np.ascontiguousarray(counts, np.uint32)
    si, sl = np.ascontiguousarray(si, np.int32), np.ascontiguousarray(sl, np.float32)
    bi, ba = np.ascontiguousarray(bi, np.uint32), np.ascontiguousarray(ba, np.float32)
    ok, tiles = C.c_int(), C.c_uint()
    rest = np.zeros(4 * int(counts.sum()), np.float32)
    why = C.create_string_buffer(128)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.velvet_grid_plan_check(p(counts), len(counts), p(si), p(sl), len(sl), p(bi), p(ba), len(ba), C.byref(ok), C.byref(tiles),
                                    p(rest) if want_rest else None, why) == 0
    return bool(ok.value), tiles.value, rest.reshape(-1, 4), why.value.decode()


def test_grid_cloths_are_recognised_from_their_constraint_pattern():
    """Host logic of the Jacobi-kernel choice (grid_plan.cpp), no GPU: the lists VtClothObjectGPU generates (here: by the O1
    oracle's restatement of VtClothObjectGPU.hpp L75-132) are recognised, with the rest lengths filed per generating vertex;
    any deviation -- a swapped pair, an extra or missing constraint, another triangulation -- is not."""
    import numpy as np
    from oracle import o1

    def lists(R, off=0):
        s = o1.O1Solver(o1.default_params())
        v, idx = o1.generate_cloth_mesh(R)
        s.cloth_object_start(R, v, idx, o1.transform_matrix((0.1, 1.5, 1.0), (70, 10, 0), (1, 1, 1)), [])
        return (s.buffer("stretchIndices").copy() + off, s.buffer("stretchLengths").copy(),
                s.buffer("bendIndices").copy() + off, s.buffer("bendAngles").copy())

    R = 20
    n = (R + 1) ** 2
    si, sl, bi, ba = lists(R)
    ok, tiles, rest, why = _grid_check([n], si, sl, bi, ba)
    assert ok and tiles == ((R + 1 + 14) // 15) ** 2, why
    # rest lengths per generating vertex: (x,y)-(x,y+1), (x,y)-(x+1,y), (x,y)-(x+1,y+1), (x,y+1)-(x+1,y); 0 where absent
    pairs = si.reshape(-1, 2)
    side = R + 1
    expect = np.zeros((n, 4), np.float32)
    for (a, b), length in zip(pairs, sl):
        lo = min(a, b)
        d = (b - a)
        kind = {1: 0, side: 1, side + 1: 2}.get(d)
        if kind is None:  # anti-diagonal (x,y+1)-(x+1,y): generated at vertex (x,y) = a - 1
            assert d == side - 1
            expect[a - 1, 3] = length
        else:
            expect[lo, kind] = length
    assert np.array_equal(rest, expect)
    assert (rest[:, 0] > 0).sum() == R * (R + 1) and (rest[:, 3] > 0).sum() == R * R

    bad = si.copy(); bad[[6, 7]] = bad[[7, 6]]  # one pair reversed
    assert not _grid_check([n], bad, sl, bi, ba)[0]
    assert not _grid_check([n], si[:-2], sl[:-1], bi, ba)[0]  # one constraint missing
    assert not _grid_check([n], np.r_[si, [0, 5]], np.r_[sl, [0.3]], bi, ba)[0]  # one extra
    other = bi.reshape(-1, 4)[:, [1, 0, 2, 3]].reshape(-1)  # another quad orientation
    assert not _grid_check([n], si, sl, other, ba)[0]
    assert not _grid_check([n - 1], si, sl, bi, ba)[0]  # not a square grid

    # two cloths of different size registered one after the other
    R2 = 7
    s2 = lists(R2, off=n)
    ok, tiles, rest, why = _grid_check([n, (R2 + 1) ** 2], np.r_[si, s2[0]], np.r_[sl, s2[1]], np.r_[bi, s2[2].astype(np.uint32)], np.r_[ba, s2[3]])
    assert ok and tiles == 4 + 1, why
    assert np.array_equal(rest[:n], expect)


def test_grid_tile_shape_follows_the_cloth_side():
    """grid_plan.cpp: 15 x 15 owned particles per tile, or 14 x 16 when that needs at least 5 % fewer tiles (a side of 64:
    5 x 4 instead of 5 x 5; a side of 32: 3 x 2 instead of 3 x 3); one shape for all the cloths of a solver."""
    import numpy as np
    from oracle import o1

    def lists(R, off):
        s = o1.O1Solver(o1.default_params())
        v, idx = o1.generate_cloth_mesh(R)
        s.cloth_object_start(R, v, idx, o1.transform_matrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1)), [])
        return (s.buffer("stretchIndices").copy() + off, s.buffer("stretchLengths").copy(),
                (s.buffer("bendIndices").copy() + off).astype(np.uint32), s.buffer("bendAngles").copy())

    def tiles(sides):
        si, sl, bi, ba, counts, off = [], [], [], [], [], 0
        for side in sides:
            a, b, c, d = lists(side - 1, off)
            si.append(a); sl.append(b); bi.append(c); ba.append(d)
            counts.append(side * side)
            off += side * side
        ok, n, _, why = _grid_check(counts, np.concatenate(si), np.concatenate(sl), np.concatenate(bi), np.concatenate(ba))
        assert ok, why
        return n

    square = lambda s: (-(-s // 15)) ** 2
    rect = lambda s: -(-s // 14) * -(-s // 16)
    assert tiles([64]) == rect(64) == 20 and square(64) == 25
    assert tiles([32]) == rect(32) == 6
    assert tiles([21]) == square(21) == 4          # no gain: the square shape stays
    assert tiles([128]) == square(128) == 81        # 80 rectangular tiles are less than 5 % fewer
    assert tiles([64, 21]) == rect(64) + rect(21)   # one shape per solver: 24 tiles against 29

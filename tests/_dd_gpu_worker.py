"""torchrun worker (one process per GPU): the domain-decomposed cloth must reproduce the single-GPU solver bit for bit.
Usage: torchrun --nproc-per-node G tests/_dd_gpu_worker.py <outdir> [resolution] [frames] [bench_frames] [peer|nccl]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import velvet_b200 as vb  # noqa: E402
from velvet_b200.decomposed import DecomposedCloth  # noqa: E402
from util import gpu_params  # noqa: E402

out_dir = sys.argv[1]
R = int(sys.argv[2]) if len(sys.argv) > 2 else 127
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bench_frames = int(sys.argv[4]) if len(sys.argv) > 4 else 0
transport = sys.argv[5] if len(sys.argv) > 5 else "peer"
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

p = gpu_params(numSubsteps=5, numIterations=10)
cols = vb.sphere_plane_colliders()
# reference: the ordinary single-GPU solver (on every rank for small cloths, on rank 0 only for large ones)
have_ref = rank == 0 or (R + 1) ** 2 <= 300000
ref = None
if have_ref:
    ref = vb.build_scene(R, p, device=local)
    ref.UpdateColliders(cols)
dd_solver = vb.build_scene(R, p, device=local)
dd_solver.UpdateColliders(cols)
dd = DecomposedCloth(dd_solver, local, transport=transport)
ok = True
worst = 0.0
checksum = 0.0
walk_band = os.environ.get("VELVET_DD_WALK_BAND")  # force the band-ordered candidate walk in the decomposed solver only
for f in range(frames):
    if walk_band:
        os.environ["VELVET_WALK_BAND"] = walk_band
    dd.Simulate()
    os.environ.pop("VELVET_WALK_BAND", None)
    b = dd_solver.download("positions")
    checksum = float(np.sum(b.astype(np.float64)))
    if have_ref:
        ref.Simulate()
        a = ref.download("positions")
        worst = max(worst, float(np.max(np.abs(a - b))))
        ok = ok and np.array_equal(a, b) and np.array_equal(ref.download("normals"), dd_solver.download("normals")) \
            and np.array_equal(ref.download("velocities"), dd_solver.download("velocities"))
# every rank must hold the same full state after the per-substep all-gather
sums = [None] * world
dist.all_gather_object(sums, checksum)
ok = ok and all(s == sums[0] for s in sums)
res = {"rank": rank, "world": world, "transport": dd.transport, "transport_requested": transport, "peer_error": dd.peer_error,
       "launches_per_frame": int(dd_solver._L.velvet_solver_last_launch_count(dd_solver._h)), "particles": int(dd_solver.simParams.numParticles), "bit_identical": bool(ok), "max_abs_diff": worst,
       "compared_with_single_gpu": bool(have_ref),
       "tiles": [int(dd.info.tileBegin), int(dd.info.tileEnd), int(dd.info.numTiles)], "owned": int(dd.info.ownedCount),
       "halo_send": int(dd.info.sendTotal), "halo_recv": int(dd.info.recvTotal)}
if bench_frames:
    stream = torch.cuda.ExternalStream(dd_solver.stream)
    for s, name in ((ref, "single_gpu_ms"), (dd, "decomposed_ms")):
        if s is None:
            dist.barrier(); dist.barrier()
            continue
        for _ in range(3):
            s.Simulate(sync=False)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(bench_frames):
            s.Simulate(sync=False)
        torch.cuda.synchronize(); dist.barrier()
        res[name] = (time.perf_counter() - t0) * 1e3 / bench_frames
dd.close()
json.dump(res, open(os.path.join(out_dir, f"dd_gpu{rank}.json"), "w"))
print(json.dumps(res), flush=True)
dist.destroy_process_group()

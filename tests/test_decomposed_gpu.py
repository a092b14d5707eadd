"""Domain-decomposed single cloth (north_star mode 2, BASELINE configs[4]).

* One-GPU box: all shards on one device in one process (velvet_b200.decomposed.LocalShards): ownership, exchange lists,
  owned-only collide / neighbour walk and the stepped schedule against the single-GPU solver, bit for bit.
* Two or more GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_decomposed_gpu.py -m gpu`): one process per GPU with the
  NVLink peer-memory transport and with NCCL."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport,walk_band", [("peer", None), ("nccl", None), ("peer", 5000)])
@pytest.mark.parametrize("resolution", [63, 255])
def test_decomposed_cloth_is_bit_identical_to_single_gpu(tmp_path, resolution, transport, walk_band):
    """walk_band: the band-ordered candidate walk of the owned strip (what a cloth of more than 1.5 M particles uses), forced
    at this size in the decomposed ranks only -- the single-GPU solver they are compared with walks unbanded."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = min(_ngpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_dd_gpu_worker.py"), str(tmp_path), str(resolution), "4", "0", transport]
    env = dict(os.environ)
    if walk_band is not None:
        env["VELVET_DD_WALK_BAND"] = str(walk_band)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    outs = [json.load(open(tmp_path / f"dd_gpu{k}.json")) for k in range(world)]
    for o in outs:
        assert o["bit_identical"], o
        # the NVLink peer-memory transport must really be the one that ran (no silent NCCL fallback on an NVLink box)
        assert o["transport"] == transport, o
        assert o["halo_send"] > 0 and o["halo_recv"] > 0 and o["halo_send"] < o["owned"]
    assert sum(o["owned"] for o in outs) == outs[0]["particles"]


@pytest.mark.parametrize("resolution,shards,attached", [(63, 3, ()), (127, 4, (0, 127)), (255, 2, ())])
def test_logical_shards_on_one_device_match_the_single_gpu_solver(resolution, shards, attached):
    import numpy as np

    import velvet_b200 as vb
    from velvet_b200.decomposed import LocalShards
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import gpu_params

    p = gpu_params(numSubsteps=5, numIterations=10)
    cols = vb.sphere_plane_colliders()
    ref = vb.build_scene(resolution, p, attached=attached)
    ref.UpdateColliders(cols)
    parts = [vb.build_scene(resolution, p, attached=attached) for _ in range(shards)]
    for s in parts:
        s.UpdateColliders(cols)
    dd = LocalShards(parts)
    owned = [int(i.ownedCount) for i in dd.info]
    assert sum(owned) == ref.simParams.numParticles and min(owned) > 0
    assert all(int(i.sendTotal) > 0 and int(i.recvTotal) > 0 for i in dd.info)
    for _ in range(3):
        ref.Simulate()
        dd.Simulate()
    for name in ("positions", "velocities", "normals", "predicted"):
        a = ref.download(name)
        for r, s in enumerate(parts):
            assert np.array_equal(a, s.download(name)), f"{name}: shard {r} of {shards} differs from the single-GPU solver"

"""Domain-decomposed single cloth (north_star mode 2) on real GPUs: needs at least 2 devices (run with
`gpurun --gpus 2 -- python -m pytest tests/test_decomposed_gpu.py -m gpu`); skipped on a 1-GPU box."""
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
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("resolution", [63, 255])
def test_decomposed_cloth_is_bit_identical_to_single_gpu(tmp_path, resolution, transport):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = min(_ngpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_dd_gpu_worker.py"), str(tmp_path), str(resolution), "4", "0", transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    outs = [json.load(open(tmp_path / f"dd_gpu{k}.json")) for k in range(world)]
    for o in outs:
        assert o["bit_identical"], o
        # the NVLink peer-memory transport must really be the one that ran (no silent NCCL fallback on an NVLink box)
        assert o["transport"] == transport, o
        assert o["halo_send"] > 0 and o["halo_recv"] > 0 and o["halo_send"] < o["owned"]
    assert sum(o["owned"] for o in outs) == outs[0]["particles"]

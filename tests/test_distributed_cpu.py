"""CPU-only (gloo, world_size 2): the multi-GPU host logic of the batched-independent-cloths mode -- instance sharding,
barrier, max-over-ranks timing, whole-job aggregation -- and the N > 1 behaviour of bench.py's reference arm."""
import json
import os
import socket
import subprocess
import sys

import pytest

from velvet_b200.distributed import instance_model_height, shard_instances

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(nproc, script_args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port())] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_shard_instances_partitions_exactly():
    for n, world in ((4096, 1), (4096, 2), (4096, 8), (4097, 8), (5, 8), (0, 4)):
        seen = []
        for r in range(world):
            seen += list(shard_instances(n, world, r))
        assert seen == list(range(n))
        sizes = [len(shard_instances(n, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_instances(10, 2, 2)
    assert instance_model_height(0) == 1.5 and abs(instance_model_height(33) - 1.51) < 1e-12


def test_two_rank_gloo_timing_and_aggregation(tmp_path):
    r = _torchrun(2, [os.path.join(ROOT, "tests", "_dist_worker.py"), str(tmp_path)])
    assert r.returncode == 0, r.stderr[-2000:]
    outs = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(2)]
    assert [o["world"] for o in outs] == [2, 2]
    assert outs[0]["first"] == 0 and outs[0]["count"] == 2049 and outs[1]["first"] == 2049 and outs[1]["count"] == 2048
    for o in outs:
        assert o["slowest"] == 2.0, "timing is the max over ranks"
        assert o["total"] == 4097.0
        assert abs(o["thr"] - 4097 * 4096 * 5 / 2.0) < 1e-3, "whole-job units / slowest rank's time"


def test_bench_reference_arm_prints_one_line_under_torchrun():
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                      "--cpu-resolution", "31"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "rank 0 alone prints; the other ranks exit 0 without work"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0


def test_exchange_lists_are_consistent_between_ranks():
    """Domain decomposition (north_star mode 2), host logic only: what rank r sends to q is exactly what q expects from r,
    ids are owned by the sender, nobody needs its own particles, and every rank's owned ranges tile [0, N)."""
    from velvet_b200.decomposed import plan_grid
    import numpy as np
    R, tile = 95, 256
    n = (R + 1) ** 2
    for world in (2, 3, 8):
        plans = [plan_grid(R, r, world, tile) for r in range(world)]
        ranges = [p[2] for p in plans]
        assert ranges[0][0] == 0 and ranges[-1][1] == n and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        total = 0
        for r in range(world):
            send, recv, _ = plans[r]
            assert len(send[r]) == 0 and len(recv[r]) == 0
            for q in range(world):
                assert np.array_equal(send[q], plans[q][1][r]), (world, r, q)
                assert np.all(np.diff(send[q].astype(np.int64)) > 0), "ascending, no duplicates"
                total += len(send[q])
        # a 2D sheet cut into compact patches exchanges O(perimeter) particles, far fewer than it owns
        assert 0 < total < n // 2


def test_two_rank_gloo_halo_exchange_follows_the_lists(tmp_path):
    r = _torchrun(2, [os.path.join(ROOT, "tests", "_dd_worker.py"), str(tmp_path)])
    assert r.returncode == 0, r.stderr[-2000:]
    outs = [json.load(open(tmp_path / f"dd{k}.json")) for k in range(2)]
    assert all(o["ok"] for o in outs) and outs[0]["sent"] == outs[1]["received"] and outs[1]["sent"] == outs[0]["received"]

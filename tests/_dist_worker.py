"""Worker for tests/test_distributed_cpu.py: run under torchrun with the gloo backend (no GPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from velvet_b200.distributed import Group, aggregate_throughput, shard_instances  # noqa: E402


def main():
    g = Group("gloo")
    mine = shard_instances(4097, g.world, g.rank)
    g.barrier()
    # rank r pretends to need (r + 1) seconds for its shard
    slowest = g.max(float(g.rank + 1))
    total = g.sum(float(len(mine)))
    thr = aggregate_throughput(len(mine) * 4096 * 5, float(g.rank + 1), g)
    out = {"rank": g.rank, "world": g.world, "first": mine.start, "count": len(mine), "slowest": slowest, "total": total, "thr": thr}
    with open(os.path.join(sys.argv[1], f"rank{g.rank}.json"), "w") as f:
        json.dump(out, f)
    g.close()


if __name__ == "__main__":
    main()

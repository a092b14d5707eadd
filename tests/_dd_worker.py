"""Worker for test_two_rank_gloo_halo_exchange_follows_the_lists: the halo exchange of the decomposed mode on CPU tensors
(gloo), with the particle id itself as payload so that every received value can be checked."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from velvet_b200.decomposed import plan_grid  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
send, recv, owned = plan_grid(63, rank, world, 128)
ops, bufs = [], {}
for q in range(world):
    if q == rank:
        continue
    if len(send[q]):
        ops.append(dist.P2POp(dist.isend, torch.from_numpy(send[q].astype(np.float32)), q))
    if len(recv[q]):
        bufs[q] = torch.zeros(len(recv[q]), dtype=torch.float32)
        ops.append(dist.P2POp(dist.irecv, bufs[q], q))
for w in dist.batch_isend_irecv(ops):
    w.wait()
ok = all(np.array_equal(bufs[q].numpy().astype(np.uint32), recv[q]) for q in bufs)
json.dump({"ok": bool(ok), "sent": int(sum(len(send[q]) for q in send)), "received": int(sum(len(recv[q]) for q in recv))},
          open(os.path.join(sys.argv[1], f"dd{rank}.json"), "w"))
dist.destroy_process_group()

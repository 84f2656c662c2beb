#!/usr/bin/env python
"""GPU-box tool (torchrun, N ranks): time shard.gather_hit_records on 10 000 hits per rank"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from spaln_b200 import shard  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
rng = np.random.default_rng(rank)
n = 10000
n_skl = rng.integers(4, 20, size=n)
corners = np.zeros((int(n_skl.sum()), 2), np.int32)
corners[:, 0] = rng.integers(0, 3000, size=len(corners))
corners[:, 1] = rng.integers(0, 60000, size=len(corners))
hits = shard.make_hits(np.arange(rank, n * world, world), rng.integers(0, 9999, size=n), n_skl,
                       rng.integers(1000, 3000, size=n), corners)
mods = {"gather_hit_records": shard}
for name, m in mods.items():
    ts = []
    for _ in range(8):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = m.gather_hit_records(hits, corners, 0, dev)
        ts.append(1e3 * (time.perf_counter() - t0))
    if rank == 0:
        print(name, "gather ms:", [round(x, 1) for x in ts], "hits", len(got[0]), flush=True)
dist.destroy_process_group()

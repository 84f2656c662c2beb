"""Query sharding across the GPUs of one box.

DP problems are independent (the reference already deals one query per pthread,
src/spaln.cc:1389-1468), so the multi-GPU path has no data-path collective: every rank takes
a cell-balanced share of the problems, runs its own engine, and the hit records (score +
trace-back corners per problem) are gathered once at the end.  The only other collective is
the one-off broadcast of a formatted genome buffer from rank 0.  One process per GPU;
`torch.distributed` is plumbing only (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def lpt_partition(cells, world: int):
    """Longest-processing-time-first assignment of problems to ranks.
    cells[i] = estimated DP cells of problem i.  Returns a list of index arrays (ascending
    inside each rank) whose cell totals differ by at most the largest problem."""
    cells = np.asarray(cells, np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.empty(len(cells), np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += cells[i]
    return [np.nonzero(owner == r)[0] for r in range(world)]


def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def broadcast_genome(buf: np.ndarray, src: int = 0, device=None) -> np.ndarray:
    """one-off broadcast of a formatted genome / table buffer (uint8) from rank `src`"""
    torch, dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return buf
    dev = device if device is not None else "cpu"
    n = torch.tensor([buf.size if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(buf, np.uint8).ravel()))
    dist.broadcast(t, src)
    return t.cpu().numpy()


def gather_hits(index, scores, skls, dst: int = 0, device=None):
    """Two-phase gather of this rank's hit records to rank `dst`: counts first
    (all_gather of int64), then the fixed-size headers and the variable-size corner payload
    padded to the largest rank.  index[i] = global problem number, scores[i] = DP score,
    skls[i] = (k_i, 2) int32 corners.  Returns on `dst` a dict global index -> (score, corners);
    None elsewhere."""
    torch, dist = _dist()
    index = np.asarray(index, np.int64)
    scores = np.asarray(scores, np.int64)
    lens = np.array([len(s) for s in skls], np.int64)
    payload = (np.concatenate([np.asarray(s, np.int32).reshape(-1, 2) for s in skls])
               if len(skls) and lens.sum() else np.zeros((0, 2), np.int32))
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return _assemble([index], [scores], [lens], [payload])
    dev = device if device is not None else "cpu"
    world, rank = dist.get_world_size(), dist.get_rank()
    cnt = torch.tensor([len(index), len(payload)], dtype=torch.int64, device=dev)
    cnts = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [c.cpu().numpy() for c in cnts]
    max_n = max(int(c[0]) for c in cnts)
    max_p = max(int(c[1]) for c in cnts)
    hdr = torch.zeros((max(max_n, 1), 3), dtype=torch.int64, device=dev)
    if len(index):
        hdr[:len(index)] = torch.from_numpy(np.stack([index, scores, lens], axis=1))
    pay = torch.zeros((max(max_p, 1), 2), dtype=torch.int32, device=dev)
    if len(payload):
        pay[:len(payload)] = torch.from_numpy(payload)
    hdrs = [torch.zeros_like(hdr) for _ in range(world)] if rank == dst else None
    pays = [torch.zeros_like(pay) for _ in range(world)] if rank == dst else None
    dist.gather(hdr, hdrs, dst)
    dist.gather(pay, pays, dst)
    if rank != dst:
        return None
    idx_l, sc_l, len_l, pay_l = [], [], [], []
    for r in range(world):
        n, p = int(cnts[r][0]), int(cnts[r][1])
        h = hdrs[r][:n].cpu().numpy()
        idx_l.append(h[:, 0])
        sc_l.append(h[:, 1])
        len_l.append(h[:, 2])
        pay_l.append(pays[r][:p].cpu().numpy())
    return _assemble(idx_l, sc_l, len_l, pay_l)


def _assemble(idx_l, sc_l, len_l, pay_l):
    out = {}
    for idx, sc, ln, pay in zip(idx_l, sc_l, len_l, pay_l):
        off = np.concatenate([[0], np.cumsum(ln)])
        for j in range(len(idx)):
            out[int(idx[j])] = (int(sc[j]), pay[off[j]:off[j + 1]].copy())
    return out

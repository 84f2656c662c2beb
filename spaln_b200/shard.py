"""Query sharding across the GPUs of one box.

DP problems are independent (the reference already deals one query per pthread,
src/spaln.cc:1389-1468), so the multi-GPU path has no data-path collective: rank 0 owns the
formatted genome and the query set and broadcasts them once, every rank takes a cell-balanced
share of the problems (longest-processing-time partition by `gspaln_task_cells`), runs its own
engine, and the hit records are gathered once at the end.  One process per GPU;
`torch.distributed` is plumbing only (NCCL on the GPU box, gloo in the CPU tests).

Hit records on the wire: a fixed-size header per hit whose leading 72 bytes are the
reference's `GeneRecord` (src/seq.h:1235-1255: what `spaln -O12` writes to its `.grd` file and
`sortgrcd` reads), followed by two words that locate the hit's trace-back corners ((m, n) pairs, the
`SKL` list `skl_rng*_ng` turns into exon records) in one flat payload.  Every rank sends ONE
contiguous header block and ONE contiguous payload block; the receiver assembles with array
operations only.
"""
from __future__ import annotations

import heapq

import numpy as np

# GeneRecord (src/seq.h:1235-1255), then our payload locator
GENE_RECORD_FIELDS = [
    ("Cid", "<i4"), ("Gstart", "<i4"), ("Gend", "<i4"), ("Nrecord", "<u4"), ("nexn", "<u4"),
    ("Rid", "<i4"), ("Rlen", "<i4"), ("Rstart", "<i4"), ("Rend", "<i4"), ("mmc", "<i4"), ("unp", "<i4"),
    ("bmmc", "<i4"), ("bunp", "<i4"), ("ng", "<i4"), ("Gscore", "<f4"), ("Pmatch", "<f4"),
    ("Pcover", "<f4"), ("Csense", "<i2"), ("Rsense", "<i2"),
]
HIT_DTYPE = np.dtype(GENE_RECORD_FIELDS + [("n_skl", "<i4"), ("skl_off", "<i4")])
assert HIT_DTYPE.itemsize == 72 + 8


def lpt_partition(cells, world: int, exact_below: int = 4096):
    """Cell-balanced assignment of problems to ranks, largest problems first.
    cells[i] = DP cells of problem i (gspaln_task_cells).  Up to `exact_below` problems: the
    longest-processing-time rule proper (each problem to the least loaded rank, heap).  Larger
    jobs: the sorted problems are dealt in serpentine order (0 .. N-1, N-1 .. 0, ...), which is
    one vectorised sort and balances the ranks to within the largest problem just the same.
    Returns a list of index arrays (ascending inside each rank)."""
    cells = np.asarray(cells, np.int64)
    order = np.argsort(-cells, kind="stable")
    owner = np.empty(len(cells), np.int64)
    if len(cells) <= exact_below:
        heap = [(0, r) for r in range(world)]
        for i, c in zip(order.tolist(), cells[order].tolist()):
            load, r = heapq.heappop(heap)
            owner[i] = r
            heapq.heappush(heap, (load + c, r))
    else:
        k = np.arange(len(cells)) % (2 * world)
        owner[order] = np.where(k < world, k, 2 * world - 1 - k)
    return [np.nonzero(owner == r)[0] for r in range(world)]


def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def broadcast_buffers(bufs, src: int = 0, device=None, to_host: bool = True):
    """One-off broadcast of the formatted genome / query set / index tables from rank `src`:
    every buffer is sent as raw bytes and comes back with the dtype and shape it had on `src`
    (pass None on the other ranks).  NCCL when `device` is a CUDA device, gloo otherwise.
    `bufs` may hold torch tensors (e.g. pinned host staging buffers) instead of numpy arrays.
    to_host = False leaves the received bytes on the device (uint8 tensors) -- where the scan and
    DP kernels of a run consume them -- and skips the device-to-host copy."""
    torch, dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [b if not to_host or isinstance(b, np.ndarray) else b.numpy() for b in bufs]
    dev = device if device is not None else "cpu"
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        bufs = [b if torch.is_tensor(b) else torch.from_numpy(np.ascontiguousarray(b)) for b in bufs]
        meta = [[(str(b.numpy().dtype.str), tuple(b.shape)) for b in bufs]]
    dist.broadcast_object_list(meta, src)
    out = []
    for k, (dt, shape) in enumerate(meta[0]):
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if rank == src:
            t.copy_(bufs[k].view(torch.uint8).reshape(-1), non_blocking=True)
        dist.broadcast(t, src)
        if not to_host:
            out.append(t)
        elif rank == src:
            out.append(bufs[k].numpy())
        else:
            out.append(t.cpu().numpy().view(np.dtype(dt)).reshape(shape))
    return out


def broadcast_genome(buf: np.ndarray, src: int = 0, device=None) -> np.ndarray:
    """one-off broadcast of a formatted genome / table buffer (uint8) from rank `src`"""
    return broadcast_buffers([buf], src, device)[0].astype(np.uint8, copy=False).ravel()


def make_hits(index, scores, n_skl, q_len, corners, scale=10.0, min_intron=20):
    """Hit headers (HIT_DTYPE) of one rank from flat DP results: index[i] = global query number,
    scores[i] = DP score, corners = the (sum n_skl, 2) int32 corner payload in hit order (alignment
    end first, as the engine returns it).  The GeneRecord fields the DP determines are filled in
    (ids, spans on both sequences, number of exons = 1 + jumps of the genomic coordinate at a
    fixed query coordinate that are at least `min_intron` long, score in the reference's output
    scale); the match statistics come from skl_rng*_ng on the host and stay 0 here."""
    index = np.asarray(index, np.int64)
    n_skl = np.asarray(n_skl, np.int64)
    n = len(index)
    h = np.zeros(n, HIT_DTYPE)
    off = np.concatenate([[0], np.cumsum(n_skl)])
    h["Rid"] = index
    h["Nrecord"] = np.arange(n, dtype=np.uint32)
    h["Rlen"] = np.asarray(q_len, np.int64)
    h["Gscore"] = np.asarray(scores, np.float64) / scale
    h["n_skl"] = n_skl
    h["skl_off"] = off[:-1]
    corners = np.asarray(corners, np.int32).reshape(-1, 2)
    has = n_skl > 0
    if has.any():
        first, last = off[:-1][has], off[1:][has] - 1
        h["Rend"][has] = corners[first, 0]
        h["Gend"][has] = corners[first, 1]
        h["Rstart"][has] = corners[last, 0]
        h["Gstart"][has] = corners[last, 1]
        if len(corners) > 1:
            same_hit = np.ones(len(corners) - 1, bool)
            same_hit[off[1:-1][(off[1:-1] > 0) & (off[1:-1] < len(corners))] - 1] = False
            jump = same_hit & (corners[:-1, 0] == corners[1:, 0]) & \
                (corners[:-1, 1] - corners[1:, 1] >= min_intron)
            owner = np.searchsorted(off[1:], np.nonzero(jump)[0], side="right")
            h["nexn"] = (has.astype(np.uint32) + np.bincount(owner, minlength=n).astype(np.uint32))
        else:
            h["nexn"] = has.astype(np.uint32)
    return h


def gather_hit_records(hits: np.ndarray, corners: np.ndarray, dst: int = 0, device=None):
    """Gather of the hit records of all ranks on rank `dst`: sizes by all_gather, then one
    contiguous header block and one contiguous corner block per rank (padded to the largest
    rank, which the cell-balanced partition keeps within a few per cent).  Returns on `dst`
    (headers sorted by query number `Rid`, corners) with `skl_off` rebased into the merged corner
    array; None elsewhere."""
    torch, dist = _dist()
    hits = np.ascontiguousarray(hits, HIT_DTYPE)
    corners = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        order = np.argsort(hits["Rid"], kind="stable")
        return hits[order], corners
    dev = device if device is not None else "cpu"
    world, rank = dist.get_world_size(), dist.get_rank()
    cnt = torch.tensor([len(hits), len(corners)], dtype=torch.int64, device=dev)
    cnts = torch.zeros(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt)
    cnts = cnts.cpu().numpy().reshape(world, 2)
    max_h, max_c = int(cnts[:, 0].max()), int(cnts[:, 1].max())
    hw = HIT_DTYPE.itemsize // 4
    hbuf = torch.zeros(max(max_h, 1) * hw, dtype=torch.int32, device=dev)
    cbuf = torch.zeros(max(max_c, 1) * 2, dtype=torch.int32, device=dev)
    if len(hits):
        hbuf[: len(hits) * hw] = torch.from_numpy(hits.view(np.int32).reshape(-1))
    if len(corners):
        cbuf[: len(corners) * 2] = torch.from_numpy(corners.reshape(-1))
    hall = [torch.empty_like(hbuf) for _ in range(world)] if rank == dst else None
    call = [torch.empty_like(cbuf) for _ in range(world)] if rank == dst else None
    dist.gather(hbuf, hall, dst)
    dist.gather(cbuf, call, dst)
    if rank != dst:
        return None
    hs, cs = [], []
    base = 0
    for r in range(world):
        nh, nc = int(cnts[r, 0]), int(cnts[r, 1])
        h = hall[r][: nh * hw].cpu().numpy().view(HIT_DTYPE).copy()
        h["skl_off"] += base
        base += nc
        hs.append(h)
        cs.append(call[r][: nc * 2].cpu().numpy().reshape(-1, 2))
    h = np.concatenate(hs) if hs else np.zeros(0, HIT_DTYPE)
    c = np.concatenate(cs) if cs else np.zeros((0, 2), np.int32)
    order = np.argsort(h["Rid"], kind="stable")
    return h[order], c


def gather_hits(index, scores, skls, dst: int = 0, device=None):
    """Convenience form over per-hit Python objects: index[i] = global problem number,
    scores[i] = DP score, skls[i] = (k_i, 2) int32 corners.  Returns on `dst` a dict
    global index -> (score, corners); None elsewhere."""
    lens = np.array([len(s) for s in skls], np.int64)
    payload = (np.concatenate([np.asarray(s, np.int32).reshape(-1, 2) for s in skls])
               if len(skls) and lens.sum() else np.zeros((0, 2), np.int32))
    h = make_hits(index, scores, lens, np.zeros(len(lens), np.int64), payload, scale=1.0)
    h["mmc"] = np.asarray(scores, np.int64)         # exact integer score next to the float one
    got = gather_hit_records(h, payload, dst, device)
    if got is None:
        return None
    hh, cc = got
    return {int(r["Rid"]): (int(r["mmc"]), cc[r["skl_off"]: r["skl_off"] + r["n_skl"]].copy()) for r in hh}

"""CPU, world_size 2 over gloo: the query-shard plumbing of the multi-GPU path
(spaln_b200/shard.py) -- cell-balanced partition, genome broadcast, two-phase gather of
variable-size hit records.  The DP itself needs a GPU; a deterministic stand-in produces the
per-problem records here."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_hit(i):
    rng = np.random.default_rng(i)
    k = int(rng.integers(0, 9))
    return int(rng.integers(-500, 5000)), rng.integers(0, 10000, size=(k, 2)).astype(np.int32)


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from spaln_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genome = np.arange(1000, dtype=np.uint8) if rank == 0 else np.zeros(0, np.uint8)
    g = shard.broadcast_genome(genome)
    # formatted genome + query set + index table in one call, dtypes and shapes preserved
    src = [np.arange(77, dtype=np.uint8), np.arange(12, dtype=np.int64).reshape(6, 2) * 1000003,
           np.arange(9, dtype=np.int16) - 4]
    got = shard.broadcast_buffers(src if rank == 0 else [None] * 3)
    assert all(x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y) for x, y in zip(got, src))
    cells = np.random.default_rng(7).integers(100, 100000, size=n)
    mine = shard.lpt_partition(cells, world)[rank]
    hits = [_fake_hit(int(i)) for i in mine]
    out = shard.gather_hits(mine, [h[0] for h in hits], [h[1] for h in hits], dst=0)
    # the vectorised form: GeneRecord headers + one flat corner block per rank
    lens = np.array([len(h[1]) for h in hits], np.int64)
    flat = np.concatenate([h[1] for h in hits]) if lens.sum() else np.zeros((0, 2), np.int32)
    hdr = shard.make_hits(mine, [h[0] for h in hits], lens, 100 + np.asarray(mine), flat, scale=10.0)
    rec = shard.gather_hit_records(hdr, flat, dst=0)
    if rank == 0:
        hh, cc = rec
        assert hh.dtype == shard.HIT_DTYPE and np.array_equal(hh["Rid"], np.arange(n))
        for i in range(n):
            sc, sk = _fake_hit(i)
            r = hh[i]
            assert abs(float(r["Gscore"]) - sc / 10.0) < 1e-3 and r["Rlen"] == 100 + i
            assert np.array_equal(cc[r["skl_off"]: r["skl_off"] + r["n_skl"]], sk)
            if len(sk):
                assert (r["Rend"], r["Gend"]) == tuple(sk[0]) and (r["Rstart"], r["Gstart"]) == tuple(sk[-1])
    else:
        assert rec is None
    if rank == 0:
        ok = len(out) == n and all(out[i][0] == _fake_hit(i)[0] and
                                   np.array_equal(out[i][1], _fake_hit(i)[1]) for i in range(n))
        q.put((ok, bool(np.array_equal(g, np.arange(1000, dtype=np.uint8)))))
    else:
        assert out is None
        q.put((True, bool(np.array_equal(g, np.arange(1000, dtype=np.uint8)))))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_shard_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=150) for _ in procs]
    [p.join(30) for p in procs]
    assert all(ok and g for ok, g in res)
    assert all(p.exitcode == 0 for p in procs)


def test_lpt_partition_balances_cells():
    from spaln_b200 import shard
    cells = np.random.default_rng(1).integers(1000, 10 ** 7, size=500)
    parts = shard.lpt_partition(cells, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(500))
    loads = [int(cells[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= int(cells.max())


def test_gather_single_process_passthrough():
    from spaln_b200 import shard
    out = shard.gather_hits([3, 5], [10, 20], [np.zeros((2, 2), np.int32), np.zeros((0, 2), np.int32)])
    assert out[3][0] == 10 and out[3][1].shape == (2, 2) and out[5][1].shape == (0, 2)


def test_hit_record_header_is_the_reference_generecord():
    """leading 72 bytes == GeneRecord of src/seq.h:1235-1255 (14 ints, 3 floats, 2 shorts)"""
    from spaln_b200 import shard
    names = [f[0] for f in shard.GENE_RECORD_FIELDS]
    assert names == ["Cid", "Gstart", "Gend", "Nrecord", "nexn", "Rid", "Rlen", "Rstart", "Rend", "mmc", "unp",
                     "bmmc", "bunp", "ng", "Gscore", "Pmatch", "Pcover", "Csense", "Rsense"]
    assert np.dtype(shard.GENE_RECORD_FIELDS).itemsize == 72
    # exons = 1 + genomic jumps at a fixed query coordinate
    corners = np.array([[90, 900], [60, 870], [60, 500], [30, 470], [30, 200], [0, 170]], np.int32)
    h = shard.make_hits([7], [1234], [6], [90], corners, scale=10.0, min_intron=50)
    assert h["nexn"][0] == 3 and h["Gstart"][0] == 170 and h["Gend"][0] == 900 and h["Rid"][0] == 7


def test_partition_of_large_jobs_is_balanced_too():
    """beyond 4096 problems the sorted problems are dealt in serpentine order (vectorised)"""
    from spaln_b200 import shard
    cells = np.random.default_rng(5).lognormal(16, 0.8, size=80000).astype(np.int64)
    parts = shard.lpt_partition(cells, 8)
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(80000))
    loads = np.array([int(cells[p].sum()) for p in parts])
    assert loads.max() - loads.min() <= int(cells.max())
    assert loads.max() / loads.mean() < 1.01

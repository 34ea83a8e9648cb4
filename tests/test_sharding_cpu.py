"""Multi-GPU orchestration on CPU: world_size-2 (and 3) `gloo` process groups drive GenomeRangeShard with a
numpy stand-in for the CUDA ops.  Checks the count all-gather, global offsets, the final pair gather and that
the union of the per-rank joins is exactly the single-process join (owner-of-left-start => no duplicates)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from giql_b200 import synth
from giql_b200.sharding import GenomeRangeShard, ShardedPairs, shard_ranges


class NumpyOps:
    """ops interface of giql_b200.sharding over numpy + the CPU oracle (TEST stand-in, never shipped)."""

    device = torch.device("cpu")

    @staticmethod
    def _key(code, pos):
        return (code.astype(np.int64) << 32) + pos.astype(np.int64)

    def plan(self, left, right, world):
        keys = np.sort(self._key(left["code"], left["start"]))
        n = len(keys)
        cuts = [0] + [int(keys[min(-(-g * n // world), n - 1)]) for g in range(1, world)] + [1 << 62]
        return {"cuts": cuts, "max_width": int((left["end"] - left["start"]).max()) if n else 0}

    def join_shard(self, plan, left, right, g):
        from oracle import c_oracle

        lo, hi = plan["cuts"][g], plan["cuts"][g + 1]
        ks = self._key(left["code"], left["start"])
        rows_l = np.nonzero((ks >= lo) & (ks < hi))[0]
        rs, re = self._key(right["code"], right["start"]), self._key(right["code"], right["end"])
        rows_r = np.nonzero((re > lo) & (rs < hi + plan["max_width"]))[0]
        p = c_oracle.intersects_pairs(left["code"][rows_l], left["start"][rows_l], left["end"][rows_l],
                                      right["code"][rows_r], right["start"][rows_r], right["end"][rows_r], 24, threads=1)
        return (torch.from_numpy(rows_l[p[:, 0]].astype(np.int32)), torch.from_numpy(rows_r[p[:, 1]].astype(np.int32)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, B = synth.peaks(30_000, 5), synth.peaks(25_000, 6)
        shard = GenomeRangeShard(NumpyOps(), rank, world)
        res = shard.intersects(A, B)
        gl, gr = shard.gather_pairs(res)
        out[rank] = {"counts": res.counts, "offset": res.global_offset, "n": len(res),
                     "pairs": ((res.left_rows.long() << 32) | res.right_rows.long()).numpy(),
                     "gathered": ((gl.long() << 32) | gr.long()).numpy()}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_join_over_gloo(world):
    from oracle import c_oracle

    c_oracle.build()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    A, B = synth.peaks(30_000, 5), synth.peaks(25_000, 6)
    exp = c_oracle.intersects_pairs(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"], 24)
    exp = (exp[:, 0] << 32) | exp[:, 1]
    counts = out[0]["counts"]
    assert sum(counts) == len(exp) and all(out[r]["counts"] == counts for r in range(world))
    assert [out[r]["n"] for r in range(world)] == counts
    assert [out[r]["offset"] for r in range(world)] == [sum(counts[:r]) for r in range(world)]
    union = np.sort(np.concatenate([out[r]["pairs"] for r in range(world)]))
    assert np.array_equal(union, exp)                       # exact multiset, nothing twice
    for r in range(world):                                  # final gather: rank-major, identical everywhere
        assert np.array_equal(out[r]["gathered"], np.concatenate([out[g]["pairs"] for g in range(world)]))


def test_shard_ranges_cover_the_genome_once():
    cuts = [0, 1000, 1000, 5000, 0xFFFFFFFF]
    lefts = [shard_ranges(cuts, 70, g)[0] for g in range(4)]
    assert [l[:2] for l in lefts] == [(0, 1000), (1000, 1000), (1000, 5000), (5000, 0xFFFFFFFF)]
    assert shard_ranges(cuts, 70, 2)[1] == (0, 5070, 1000)
    assert shard_ranges(cuts, 70, 3)[1] == (0, 0xFFFFFFFF, 5000)   # saturates


def test_sharded_pairs_offsets():
    r = ShardedPairs(torch.zeros(3, dtype=torch.int32), torch.zeros(3, dtype=torch.int32), [5, 3, 0, 7], 1)
    assert r.offsets == [0, 5, 8, 8] and r.global_total == 15 and r.global_offset == 5


# ---- CLUSTER / MERGE shards over gloo with a numpy stand-in ------------------------------------------------------------------
class NumpyClusterOps:
    """cm_* ops of giql_b200.sharding over numpy (TEST stand-in): linear coordinate = (code << 32) + position."""

    device = torch.device("cpu")

    def cm_plan(self, T, world):
        ks = (T["code"].astype(np.int64) << 32) + T["start"]
        order = np.argsort(ks, kind="stable")
        n = len(ks)
        cuts = [0] + [int(ks[order][min(-(-g * n // world), n - 1)]) for g in range(1, world)] + [1 << 62]
        return {"T": T, "ks": ks, "order": order, "cuts": cuts, "n_codes": int(T["code"].max()) + 1}

    def _rows(self, plan, g):
        o = plan["order"]
        k = plan["ks"][o]
        return o[(k >= plan["cuts"][g]) & (k < plan["cuts"][g + 1])]

    def cm_max_end(self, plan, g):
        r = self._rows(plan, g)
        T = plan["T"]
        return int(((T["code"][r].astype(np.int64) << 32) + T["end"][r]).max()) if len(r) else 0

    def _flags(self, plan, g, distance, carry):
        T, r = plan["T"], self._rows(plan, g)
        d = max(int(distance), 0)
        run, flags = carry, []
        for i in r:
            c, s, e = int(T["code"][i]), int(T["start"][i]), int(T["end"][i])
            prev_in_chrom = (run >> 32) == c and run > 0
            flags.append(0 if prev_in_chrom and (run & 0xFFFFFFFF) + d >= s else 1)
            run = max(run, (c << 32) + e)
        return r, np.array(flags, dtype=np.int64)

    def cm_cluster_counts(self, plan, g, distance, carry):
        r, f = self._flags(plan, g, distance, carry)
        return np.bincount(plan["T"]["code"][r], weights=f, minlength=plan["n_codes"]).astype(np.int64).tolist()

    def cm_cluster_ids(self, plan, g, distance, carry, offsets):
        r, f = self._flags(plan, g, distance, carry)
        ids = torch.zeros(len(plan["T"]["start"]), dtype=torch.int32)
        run = dict(enumerate(offsets))
        for i, fl in zip(r, f):
            c = int(plan["T"]["code"][i])
            run[c] += int(fl)
            ids[int(i)] = run[c]
        return ids

    def cm_merge(self, plan, g, distance, carry):
        T = plan["T"]
        r, f = self._flags(plan, g, distance, carry)
        regs, head = [], [0, -1, 0]
        for i, fl in zip(r, f):
            c, s, e = int(T["code"][i]), int(T["start"][i]), int(T["end"][i])
            if fl:
                regs.append([c, s, e, 1])
            elif regs:
                regs[-1][2] = max(regs[-1][2], e)
                regs[-1][3] += 1
            else:
                head = [head[0] + 1, c, max(head[2], e)]
        cols = list(zip(*regs)) if regs else [(), (), (), ()]
        return ({"code": torch.tensor(cols[0], dtype=torch.int32), "start": torch.tensor(cols[1], dtype=torch.int32),
                 "end": torch.tensor(cols[2], dtype=torch.int32), "count": torch.tensor(cols[3], dtype=torch.int64)}, tuple(head))


def _cm_worker(rank, world, port, out):
    from giql_b200.sharding import sharded_cluster, sharded_merge

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = synth.reads(6000, 9)
        ops = NumpyClusterOps()
        ids = sharded_cluster(ops, rank, world, T, distance=30)
        reg = sharded_merge(ops, rank, world, T, distance=30)
        out[rank] = {"ids": ids.numpy(), "regions": [tuple(int(v) for v in r) for r in zip(reg["code"], reg["start"], reg["end"], reg["count"])]}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_cluster_merge_over_gloo(world):
    from oracle import c_oracle

    c_oracle.build()
    out = mp.Manager().dict()
    mp.spawn(_cm_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    T = synth.reads(6000, 9)
    ids = sum(out[r]["ids"] for r in range(world))
    assert np.array_equal(ids, c_oracle.cluster(T["code"], T["start"], T["end"], 24, distance=30))
    op, os_, oe, oc = c_oracle.merge(T["code"], T["start"], T["end"], 24, distance=30)
    got = sorted(r for k in range(world) for r in out[k]["regions"])
    assert got == sorted(zip(op.tolist(), os_.tolist(), oe.tolist(), oc.tolist()))


def test_stitch_and_offsets_helpers():
    from giql_b200.sharding import carry_before, id_offsets, stitch_heads

    assert stitch_heads([(3, 0, -1, 0), (0, 5, 1, 900), (2, 4, 1, 950), (1, 0, -1, 0)]) == [(9, 950), (0, None), (0, None), (0, None)]
    assert id_offsets([[1, 2], [0, 3], [4, 0]], 2) == [1, 5] and id_offsets([[1, 2]], 0) == [0, 0]
    assert carry_before([5, 9, 2], 0) == 0 and carry_before([5, 9, 2], 2) == 9
    with pytest.raises(ValueError):
        stitch_heads([(0, 2, 0, 10)])


# ---- multi-GPU ingest (1/N upload + all-gather) over gloo ---------------------------------------------------------------------
class _CpuEngine:
    device = torch.device("cpu")


def _ingest_worker(rank, world, port, out):
    from giql_b200.sharding import upload_replicated

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        names = np.array(["chr1", "chr10", "chrX_random_contig", "", "c"], dtype=object)
        rng = np.random.default_rng(3)
        codes = rng.integers(0, len(names), size=1003)
        lens = np.array([len(names[c].encode()) for c in codes])
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        byts = np.frombuffer("".join(names[c] for c in codes).encode(), dtype=np.uint8).copy()
        cols = {"offs": torch.from_numpy(offs), "bytes": torch.from_numpy(byts),
                "start": torch.arange(1003, dtype=torch.int32), "end": torch.arange(1003, dtype=torch.int32) + 7}
        t = upload_replicated(_CpuEngine(), cols, rank, world)
        out[rank] = all(torch.equal(a, b) for a, b in ((t.chrom_offsets, cols["offs"]), (t.chrom_bytes, cols["bytes"]),
                                                         (t.start, cols["start"]), (t.end, cols["end"]))) and t.n == 1003
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_upload_replicated_over_gloo(world):
    out = mp.Manager().dict()
    mp.spawn(_ingest_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))

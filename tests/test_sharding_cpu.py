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

"""Genome-range sharding of the INTERSECTS join over the GPUs of one box.

One process per GPU.  Every rank holds both tables (replicated input); the join is partitioned by
contiguous genome ranges balanced on the LEFT interval count, right intervals that straddle a cut are
replicated to both neighbours, and a pair is owned by the rank that owns the left row's start — so the
union of the per-rank results is the exact multiset of the single-GPU join, without a dedup pass.
The data path uses no collective: every rank derives identical cuts from its own copy of the left table.
NCCL (or gloo on CPU, for the tests) carries only the all-gather of the per-rank pair counts and, on
request, the final gather of the pairs.

The class is written against a small *ops* interface so the orchestration can be exercised on CPU:

    ops.plan(left, right, world)            -> plan (opaque; identical on every rank)
    ops.join_shard(plan, left, right, g)    -> (left_rows, right_rows) of rank g's pairs, original row ids
    ops.device                              -> torch.device the count tensor lives on

``EngineOps`` is the product implementation (CUDA kernels through the C ABI).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

UINT32_MAX = 0xFFFFFFFF


class ShardedPairs:
    """This rank's pairs plus where they sit in the global result.

    ``counts`` (pairs per rank) comes from an all-gather that is ISSUED when the join finishes and awaited only when
    the counts are first read: the rank's own data path never depends on it.
    """

    def __init__(self, left_rows: torch.Tensor, right_rows: torch.Tensor, counts, rank: int):
        self.left_rows, self.right_rows, self.rank = left_rows, right_rows, rank
        self._counts = counts          # list[int], or a pending (work handle, gathered tensor)

    @property
    def counts(self) -> list[int]:
        if not isinstance(self._counts, list):
            work, gathered = self._counts
            work.wait()
            self._counts = [int(x) for x in gathered.cpu().tolist()]
        return self._counts

    def __len__(self) -> int:
        return int(self.left_rows.numel())

    @property
    def offsets(self) -> list[int]:
        out, acc = [], 0
        for c in self.counts:
            out.append(acc)
            acc += c
        return out

    @property
    def global_total(self) -> int:
        return sum(self.counts)

    @property
    def global_offset(self) -> int:
        return self.offsets[self.rank]


def shard_ranges(cuts: list[int], max_width: int, g: int) -> tuple[tuple[int, int, int], tuple[int, int, int]]:
    """(gs_lo, gs_hi, ge_gt) selections of rank g: left rows by owner-of-start, right rows that can overlap."""
    lo, hi = cuts[g], cuts[g + 1]
    left = (lo, hi, 0)
    right = (0, min(hi + max_width, UINT32_MAX), lo)
    return left, right


class EngineOps:
    """ops interface over giql_b200.Engine (hand-written CUDA; no CPU fallback): the rank-local path in ONE library call
    (giqlb_intersects_tables_sharded) — the cuts never leave the device."""

    def __init__(self, engine):
        self.e = engine
        self.device = engine.device

    def plan(self, left, right, world):
        return {"world": world}

    def join_shard(self, plan, left, right, g):
        res = self.e.intersects_shard(left, right, plan["world"], g)
        return res.left_rows, res.right_rows


class EngineOpsStepwise:
    """The same through the individual C-ABI calls (cuts fetched to the host, explicit selections): the reference
    sequence the one-call path is tested against."""

    def __init__(self, engine):
        self.e = engine
        self.device = engine.device

    def plan(self, left, right, world):
        from .engine import _Layout

        d = self.e.encode(left, right)
        lay = _Layout(self.e, len(d))
        lay.observe(left)
        if right is not left:
            lay.observe(right)
        lay.finalize()
        cuts, max_width = self.e.shard_cuts(lay, left, world)
        return {"layout": lay, "cuts": cuts, "max_width": max_width}

    def join_shard(self, plan, left, right, g):
        from .engine import _Index

        lay = plan["layout"]
        sel_l, sel_r = shard_ranges(plan["cuts"], plan["max_width"], g)
        rows_l = self.e.shard_select(lay, left, *sel_l)
        rows_r = self.e.shard_select(lay, right, *sel_r)
        il = _Index(self.e, lay, left, rows=rows_l)
        ir = _Index(self.e, lay, right, rows=rows_r)
        try:
            res = self.e._join_indexes(il, ir)
        finally:
            il.free()
            ir.free()
        return res.left_rows, res.right_rows

    def release(self, plan):
        plan["layout"].free()


class GenomeRangeShard:
    def __init__(self, engine_or_ops, rank: int, world: int, group: Optional[dist.ProcessGroup] = None):
        self.ops = engine_or_ops if hasattr(engine_or_ops, "join_shard") else EngineOps(engine_or_ops)
        self.rank, self.world, self.group = rank, world, group

    def intersects(self, left, right) -> ShardedPairs:
        plan = self.ops.plan(left, right, self.world)
        try:
            out_l, out_r = self.ops.join_shard(plan, left, right, self.rank)
        finally:
            if hasattr(self.ops, "release"):
                self.ops.release(plan)
        return ShardedPairs(out_l, out_r, self.all_gather_counts(int(out_l.numel())), self.rank)

    def all_gather_counts(self, local: int):
        """Starts the all-gather of the per-rank pair counts (one int64 each); see ShardedPairs.counts."""
        if self.world == 1 or not dist.is_initialized():
            return [local]
        mine = torch.tensor([local], dtype=torch.int64, device=self.ops.device)
        allc = torch.empty(self.world, dtype=torch.int64, device=self.ops.device)
        work = dist.all_gather_into_tensor(allc, mine, group=self.group, async_op=True)
        return (work, allc)

    def gather_pairs(self, res: ShardedPairs) -> tuple[torch.Tensor, torch.Tensor]:
        """Final pair gather: every rank receives all pairs, rank-major (rank g's slice starts at offsets[g])."""
        if self.world == 1 or not dist.is_initialized():
            return res.left_rows, res.right_rows
        width = max(res.counts) if res.counts else 0
        dev = res.left_rows.device
        mine = torch.zeros((2, max(width, 1)), dtype=torch.int32, device=dev)
        mine[0, : len(res)] = res.left_rows
        mine[1, : len(res)] = res.right_rows
        everything = torch.empty((self.world, 2, max(width, 1)), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(everything.view(-1), mine.view(-1), group=self.group)
        ls = [everything[g, 0, : res.counts[g]] for g in range(self.world)]
        rs = [everything[g, 1, : res.counts[g]] for g in range(self.world)]
        return torch.cat(ls), torch.cat(rs)

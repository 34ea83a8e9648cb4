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

    def intersects_batches(self, batches, depth: int = 2):
        """The sharded counterpart of ``Engine.intersects_batches``: a stream of HOST batches ``(left, right)`` (pinned
        column dicts, the same on every rank), each uploaded 1/world per rank + all-gathered (``upload_replicated``),
        joined on this rank's genome range, and this rank's pairs copied back — with batch i+1's upload and batch i's
        download overlapping the compute on their own streams.  Yields ``(n_local_pairs, left_rows, right_rows)``
        (pinned host tensors from a ring of ``depth`` buffers).  Every rank must consume the same batches in the same
        order (the uploads are collectives)."""
        eng = self.ops.e
        main = eng.stream
        if getattr(eng, "_copy_streams", None) is None:              # kept across calls, see Engine.intersects_batches
            eng._copy_streams = (torch.cuda.Stream(eng.device), torch.cuda.Stream(eng.device))
        s_in, s_out = eng._copy_streams
        # The ingest all-gathers get their OWN communicator: a process group runs its collectives one after the other on
        # one internal stream, so on the join's group the tiny count exchange of batch i would queue behind the eight
        # ingest all-gathers of batch i+1 (measured at 4 GPUs: 20 ms per batch instead of 6).
        if self.world > 1 and dist.is_initialized() and getattr(self, "_ingest_group", None) is None:
            self._ingest_group = dist.new_group(backend=dist.get_backend(self.group))
        ingest = getattr(self, "_ingest_group", None)

        def upload(batch):
            with torch.cuda.stream(s_in):
                tl = upload_replicated(eng, batch[0], self.rank, self.world, ingest)
                tr = upload_replicated(eng, batch[1], self.rank, self.world, ingest)
                ev = torch.cuda.Event()
                ev.record(s_in)
            return tl, tr, ev

        ring = self.__dict__.setdefault("_result_ring", [])
        ring.extend([None] * (depth - len(ring)))
        inflight = []                                                # (slot, n, done event, tensors kept alive)
        it = iter(batches)
        try:
            nxt = upload(next(it))
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            tl, tr, ev = nxt
            try:
                nxt = upload(next(it))
            except StopIteration:
                nxt = None
            main.wait_event(ev)
            res = self.intersects(tl, tr)
            n = len(res)
            done_c = torch.cuda.Event()
            done_c.record(main)
            slot = i % depth
            if len(inflight) >= depth - 1:
                s0, n0, e0, keep = inflight.pop(0)
                e0.synchronize()
                del keep
                yield n0, ring[s0][0][:n0], ring[s0][1][:n0]
            if ring[slot] is None or ring[slot][0].numel() < n:
                cap = max(int(n * 1.25), 1024)
                ring[slot] = (torch.empty(cap, dtype=torch.int32, pin_memory=True), torch.empty(cap, dtype=torch.int32, pin_memory=True))
            s_out.wait_event(done_c)
            with torch.cuda.stream(s_out):
                ring[slot][0][:n].copy_(res.left_rows[:n], non_blocking=True)
                ring[slot][1][:n].copy_(res.right_rows[:n], non_blocking=True)
                done = torch.cuda.Event()
                done.record(s_out)
            inflight.append((slot, n, done, (tl, tr, res)))
            del tl, tr, res
            i += 1
        while inflight:
            s0, n0, e0, keep = inflight.pop(0)
            e0.synchronize()
            del keep
            yield n0, ring[s0][0][:n0], ring[s0][1][:n0]

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


# ---- CLUSTER / MERGE over genome-range shards ------------------------------------------------------
# A shard owns the rows whose START lies in its range.  Two tiny exchanges stitch the shards together:
#   * the running max of end over each shard (one number per rank): the carry of rank g is the max over ranks < g;
#   * CLUSTER: the number of clusters each shard starts per chromosome -> id offsets;
#     MERGE:   each shard's head fragment (rows continuing a region of an earlier shard) -> folded into the last
#              region of the nearest earlier shard that starts one.
# ops interface (EngineOps below; a numpy stand-in in the CPU tests):
#   ops.cm_plan(table, world) -> plan;  ops.cm_max_end(plan, g) -> int
#   ops.cm_cluster_counts(plan, g, distance, carry) -> list[int] (per chromosome code)
#   ops.cm_cluster_ids(plan, g, distance, carry, offsets) -> int32 tensor [n rows of the table], 0 outside the shard
#   ops.cm_merge(plan, g, distance, carry) -> (regions: dict of equally long tensors code/start/end/count,
#                                              head: (rows, code, end))

def carry_before(max_ends: list[int], g: int) -> int:
    """Running max of end over the shards before g (0 when there is none)."""
    return max(max_ends[:g], default=0)


def id_offsets(counts_per_rank: list[list[int]], g: int) -> list[int]:
    """Clusters started by the shards before g, per chromosome code."""
    n_codes = len(counts_per_rank[0]) if counts_per_rank else 0
    return [sum(counts_per_rank[h][c] for h in range(g)) for c in range(n_codes)]


def stitch_heads(infos: list[tuple[int, int, int, int]]) -> list[tuple[int, Optional[int]]]:
    """infos[g] = (regions started by shard g, head rows, head chromosome code, head max end).  Returns, per shard,
    (extra rows, extra max end or None) to fold into the shard's LAST region: the head fragments of the later shards
    that continue it (a head continues the last region of the nearest earlier shard that starts one)."""
    out: list[list] = [[0, None] for _ in infos]
    for g, (_, head_rows, _, head_end) in enumerate(infos):
        if head_rows <= 0:
            continue
        h = g - 1
        while h >= 0 and infos[h][0] == 0:
            h -= 1
        if h < 0:
            raise ValueError(f"shard {g} continues a region but no earlier shard starts one")
        out[h][0] += head_rows
        out[h][1] = head_end if out[h][1] is None else max(out[h][1], head_end)
    return [(a, b) for a, b in out]


def _all_gather_ints(vec: list[int], world: int, device, group=None) -> list[list[int]]:
    if world == 1 or not dist.is_initialized():
        return [list(vec)]
    mine = torch.tensor(vec, dtype=torch.int64, device=device)
    allv = torch.empty(world * len(vec), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    flat = allv.cpu().tolist()
    return [flat[r * len(vec):(r + 1) * len(vec)] for r in range(world)]


def sharded_cluster(ops, rank: int, world: int, table, distance: int = 0, group=None) -> torch.Tensor:
    """Global CLUSTER ids of this rank's rows (0 for rows owned by other ranks)."""
    plan = ops.cm_plan(table, world)
    try:
        max_ends = [v[0] for v in _all_gather_ints([ops.cm_max_end(plan, rank)], world, ops.device, group)]
        carry = carry_before(max_ends, rank)
        counts = _all_gather_ints(ops.cm_cluster_counts(plan, rank, distance, carry), world, ops.device, group)
        return ops.cm_cluster_ids(plan, rank, distance, carry, id_offsets(counts, rank))
    finally:
        if hasattr(ops, "cm_release"):
            ops.cm_release(plan)


def sharded_merge(ops, rank: int, world: int, table, distance: int = 0, group=None) -> dict:
    """The merged regions that START in this rank's genome range, with the rows of later shards that continue the
    last one folded in.  Returns a dict of equally long tensors: code, start, end, count."""
    plan = ops.cm_plan(table, world)
    if isinstance(plan, dict):
        plan["row_ids"] = False        # MERGE never maps a sorted position back to an input row
    try:
        max_ends = [v[0] for v in _all_gather_ints([ops.cm_max_end(plan, rank)], world, ops.device, group)]
        carry = carry_before(max_ends, rank)
        regions, head = ops.cm_merge(plan, rank, distance, carry)
        n_regions = int(regions["start"].numel())
        infos = [tuple(v) for v in _all_gather_ints([n_regions, head[0], head[1], head[2]], world, ops.device, group)]
        extra_rows, extra_end = stitch_heads(infos)[rank]
        if extra_rows:
            regions["count"][-1] += extra_rows
            regions["end"][-1] = max(int(regions["end"][-1]), int(extra_end))
        return regions
    finally:
        if hasattr(ops, "cm_release"):
            ops.cm_release(plan)


def _engine_cm_plan(self, table, world):
    from .engine import _Layout

    d = self.e.encode(table)
    lay = _Layout(self.e, len(d))
    lay.observe(table, raw=True)
    lay.finalize()
    cuts = (ctypes.c_uint32 * (world + 1))()
    # no widest-interval request: CLUSTER / MERGE shards replicate no rows, and the cuts may then come from a sample of the rows
    self.e._check(self.e.lib.giqlb_shard_cuts(self.e.ctx, lay.h, table.codes.data_ptr(), table.start.data_ptr(), table.end.data_ptr(),
                                              table.n, 0, 0, world, cuts, None))
    return {"layout": lay, "cuts": list(cuts), "table": table, "index": {}, "n_codes": len(d)}


def _engine_cm_index(self, plan, g):
    from .engine import _Index

    if g not in plan["index"]:
        t, lay = plan["table"], plan["layout"]
        # one library call: fused select + keygen of the shard's rows, sort (giqlb_index_build_range); a plan made for MERGE
        # (plan["row_ids"] = False) asks for no row ids: regions, not rows, come out of it
        plan["index"][g] = _Index(self.e, lay, t, raw=True, start_range=(plan["cuts"][g], plan["cuts"][g + 1]),
                                  row_ids=plan.get("row_ids", True))
    return plan["index"][g]


def _engine_cm_max_end(self, plan, g):
    v = ctypes.c_uint32()
    self.e._check(self.e.lib.giqlb_shard_max_end(self.e.ctx, _engine_cm_index(self, plan, g).h, ctypes.byref(v)))
    return int(v.value)


def _engine_cm_cluster_counts(self, plan, g, distance, carry):
    counts = (ctypes.c_uint32 * max(plan["n_codes"], 1))()
    self.e._check(self.e.lib.giqlb_cluster_shard(self.e.ctx, _engine_cm_index(self, plan, g).h, int(distance or 0), carry, None, None, counts))
    return list(counts)[: plan["n_codes"]]


def _engine_cm_cluster_ids(self, plan, g, distance, carry, offsets):
    ids = torch.zeros(plan["table"].n, dtype=torch.int32, device=self.e.device)
    off = (ctypes.c_uint32 * max(plan["n_codes"], 1))(*offsets)
    self.e._check(self.e.lib.giqlb_cluster_shard(self.e.ctx, _engine_cm_index(self, plan, g).h, int(distance or 0), carry, off,
                                                 ids.data_ptr(), None))
    return ids


def _engine_cm_merge(self, plan, g, distance, carry):
    ix = _engine_cm_index(self, plan, g)
    n = int(self.e.lib.giqlb_index_rows(ix.h))
    code, start, end = (self.e._empty(n, torch.int32) for _ in range(3))
    count = self.e._empty(n, torch.int64)
    m, head = ctypes.c_int64(), (ctypes.c_int64 * 3)()
    self.e._check(self.e.lib.giqlb_merge_shard(self.e.ctx, ix.h, int(distance or 0), carry, code.data_ptr(), start.data_ptr(), end.data_ptr(),
                                               count.data_ptr(), n, ctypes.byref(m), head))
    k = m.value
    return {"code": code[:k], "start": start[:k], "end": end[:k], "count": count[:k]}, (int(head[0]), int(head[1]), int(head[2]))


def _engine_cm_release(self, plan):
    for ix in plan["index"].values():
        ix.free()
    plan["layout"].free()


import ctypes  # noqa: E402  (only the Engine-backed ops need it)

for _cls in (EngineOps, EngineOpsStepwise):
    _cls.cm_plan, _cls.cm_max_end, _cls.cm_cluster_counts = _engine_cm_plan, _engine_cm_max_end, _engine_cm_cluster_counts
    _cls.cm_cluster_ids, _cls.cm_merge, _cls.cm_release = _engine_cm_cluster_ids, _engine_cm_merge, _engine_cm_release


# ---- multi-GPU ingest: every rank uploads 1/world of the rows over its own PCIe link, NVLink replicates ----------------
def upload_replicated(engine, host_cols: dict, rank: int, world: int, group=None):
    """Host columns of ONE table (pinned tensors ``offs`` int32[n+1], ``bytes`` uint8[], ``start``/``end`` int32[n], the
    same on every rank) -> a DeviceTable holding the FULL table on every rank.

    With ``world`` > 1 a rank copies only its block of rows host->device and the blocks are all-gathered over
    NVLink (NCCL), so the PCIe traffic per GPU is 1/world of the table instead of all of it.  The join itself still
    needs the whole table on every rank (replicated-input sharding).
    """
    offs, byts, start, end = host_cols["offs"], host_cols["bytes"], host_cols["start"], host_cols["end"]
    n = int(start.numel())
    if world == 1 or not dist.is_initialized():
        return engine.table_from_buffers(offs, byts, start, end)
    dev = engine.device

    def gather_flat(col: torch.Tensor, extra: int = 0) -> torch.Tensor:
        """all-gather of equal blocks of a flat host column straight into its final device array (the last block is
        padded, `extra` spare elements follow): no staging copy, no concatenation"""
        m = int(col.numel())
        per = max(-(-m // world), 1)
        lo, hi = min(rank * per, m), min((rank + 1) * per, m)
        out = torch.empty(world * per + extra, dtype=col.dtype, device=dev)
        mine = out[rank * per: (rank + 1) * per]          # NCCL all-gather in place: a rank's block already sits in its slot
        mine[: hi - lo].copy_(col[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(out[: world * per], mine, group=group)
        return out

    d_start, d_end = gather_flat(start)[:n], gather_flat(end)[:n]
    # offsets are absolute positions into the byte buffer, and the byte buffer is just a flat array: both are cut into
    # equal blocks regardless of row boundaries
    d_offs = gather_flat(offs[:n], extra=1)
    d_offs[n: n + 1].copy_(offs[n:], non_blocking=True)
    d_offs = d_offs[: n + 1]
    n_bytes = int(byts.numel())
    d_bytes = gather_flat(byts, extra=16)[: max(n_bytes, 1) + 16] if n_bytes else torch.zeros(17, dtype=torch.uint8, device=dev)
    from .engine import DeviceTable

    return DeviceTable(engine, n, None, d_start, d_end, None, d_offs, d_bytes[: max(n_bytes, 1)] if n_bytes else d_bytes[:1])

"""CPU restatement of GIQL's interval-operator semantics.  TEST INFRASTRUCTURE ONLY.

This module is the parity oracle for the B200 engine.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  Nothing under ``giql_b200/`` may import it: the
product path is CUDA-only and fails loudly when the extension is missing.

What it restates.  The reference (abdenlab/giql) is a SQL transpiler; the
arithmetic on this path runs inside DuckDB (``duckdb>=1.4.0``, unpinned, not
vendored under /root/reference, not installed in this image).  The functions
below restate the *semantics the reference's SQL text pins*, each citing the
reference file:line it follows.  Parity status: pinned at small scale against
the reference's own golden vectors (``tests/golden/*.json``, harvested from the
reference test-suite by ``tests/golden/harvest_reference_vectors.py``) and
against outputs of the reference's own ``generate_distance_case`` executed on
stdlib SQLite; *unpinned against DuckDB itself* because DuckDB is absent.

Conventions: a table is a dict of equally long python lists / numpy arrays with
keys ``chrom`` (str), ``start`` (int), ``end`` (int) and optionally ``strand``
(str or None).  Row ids are positions in those lists.  Coordinates are canonical
0-based half-open unless a function says otherwise.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

# --------------------------------------------------------------------------- #
# canonical coordinates  (reference: src/giql/canonical.py:16-52)
# --------------------------------------------------------------------------- #

def canonical_offsets(coordinate_system: str = "0based", interval_type: str = "half_open"):
    """Integer (start_adj, end_adj) such that canonical = raw + adj.

    reference: src/giql/canonical.py:16-29 (start), :32-52 (end);
    Table validation src/giql/table.py:16-93.
    """
    if coordinate_system not in ("0based", "1based"):
        raise ValueError(f"coordinate_system must be '0based' or '1based', got {coordinate_system!r}")
    if interval_type not in ("half_open", "closed"):
        raise ValueError(f"interval_type must be 'half_open' or 'closed', got {interval_type!r}")
    s_adj = 0 if coordinate_system == "0based" else -1
    key = (coordinate_system, interval_type)
    if key == ("0based", "closed"):
        e_adj = 1
    elif key == ("1based", "half_open"):
        e_adj = -1
    else:
        e_adj = 0
    return s_adj, e_adj


# --------------------------------------------------------------------------- #
# INTERSECTS family  (reference: src/giql/expanders/intersects.py:136-166;
# per-chromosome IEJoin plan src/giql/expanders/intersects_duckdb.py:1217-1334;
# python restatements in the reference's own tests tests/test_duckdb_iejoin.py:34-81)
# --------------------------------------------------------------------------- #

def _pred(op, ls, le, rs, re):
    if op == "intersects":      # intersects.py:149-154
        return ls < re and le > rs
    if op == "contains":        # intersects.py:155-160
        return ls <= rs and le >= re
    if op == "within":          # intersects.py:161-166
        return ls >= rs and le <= re
    raise ValueError(f"Unknown spatial op_type: {op!r}")


def join_pairs_bruteforce(L, R, op="intersects"):
    """O(N*M) literal predicate; returns sorted list of (left_row, right_row).

    Mirrors ``_python_overlap`` (tests/test_duckdb_iejoin.py:34-46) but keyed by row
    ids so duplicate input rows keep their multiplicity.
    """
    out = []
    for i, (lc, ls, le) in enumerate(zip(L["chrom"], L["start"], L["end"])):
        for j, (rc, rs, re) in enumerate(zip(R["chrom"], R["start"], R["end"])):
            if lc is not None and lc == rc and _pred(op, ls, le, rs, re):
                out.append((i, j))
    return sorted(out)


def _as_cols(T):
    chrom = np.asarray(T["chrom"], dtype=object)
    start = np.asarray(T["start"], dtype=np.int64)
    end = np.asarray(T["end"], dtype=np.int64)
    return chrom, start, end


def _chrom_groups(chrom):
    groups = {}
    for i, c in enumerate(chrom):
        if c is None:
            continue
        groups.setdefault(c, []).append(i)
    return {c: np.asarray(ix, dtype=np.int64) for c, ix in groups.items()}


def join_pairs(L, R, op="intersects", block=4096):
    """Vectorised per-chromosome evaluation of the literal predicate.

    Same result as :func:`join_pairs_bruteforce`; per chromosome
    (intersects_duckdb.py:1283-1299: one branch per chrom in chroms(L) ∩ chroms(R))
    it broadcasts the predicate over blocks.  Returns an int64 array (P, 2) sorted
    lexicographically.
    """
    lc, ls, le = _as_cols(L)
    rc, rs, re = _as_cols(R)
    gl, gr = _chrom_groups(lc), _chrom_groups(rc)
    chunks = []
    for c, li in gl.items():
        ri = gr.get(c)
        if ri is None:
            continue
        order = np.argsort(rs[ri], kind="stable")
        ri = ri[order]
        r_s, r_e = rs[ri], re[ri]
        for b0 in range(0, len(li), block):
            lb = li[b0:b0 + block]
            l_s, l_e = ls[lb][:, None], le[lb][:, None]
            if op == "intersects":
                # prune with sorted starts: r.start < max(l.end)
                hi = int(np.searchsorted(r_s, int(l_e.max()), side="left"))
                m = (l_s < r_e[None, :hi]) & (l_e > r_s[None, :hi])
            elif op == "contains":
                hi = len(ri)
                m = (l_s <= r_s[None, :]) & (l_e >= r_e[None, :])
            elif op == "within":
                hi = len(ri)
                m = (l_s >= r_s[None, :]) & (l_e <= r_e[None, :])
            else:
                raise ValueError(f"Unknown spatial op_type: {op!r}")
            a, b = np.nonzero(m)
            if len(a):
                chunks.append(np.stack([lb[a], ri[:hi][b]], axis=1))
    if not chunks:
        return np.zeros((0, 2), dtype=np.int64)
    out = np.concatenate(chunks)
    key = np.lexsort((out[:, 1], out[:, 0]))
    return out[key]


def count_per_left(L, R, op="intersects"):
    """Matches per left row (int64[N_L])."""
    pairs = join_pairs(L, R, op)
    n = len(L["start"])
    return np.bincount(pairs[:, 0], minlength=n).astype(np.int64) if len(pairs) else np.zeros(n, np.int64)


def semi_rows(L, R, op="intersects"):
    """Left row ids with >=1 match, one per qualifying left *row* (duplicates kept).

    reference: intersects_duckdb.py:1254-1282 (``WHERE EXISTS``) — one output row
    per qualifying left row.
    """
    return np.nonzero(count_per_left(L, R, op) > 0)[0].astype(np.int64)


def anti_rows(L, R, op="intersects"):
    """Left row ids with 0 matches; left-only chromosomes are kept.

    reference: intersects_duckdb.py:1254-1282 (``NOT EXISTS``), partition rule
    :1317-1330 (ANTI partitions on the left table's distinct chroms);
    tests/test_duckdb_iejoin.py:5873-5919.  NULL chrom rows are dropped by the
    per-chrom plan (SURVEY §7 hard parts) — the engine rejects them at ingest.
    """
    cnt = count_per_left(L, R, op)
    keep = np.asarray([c is not None for c in L["chrom"]], dtype=bool) if len(cnt) else np.zeros(0, bool)
    return np.nonzero((cnt == 0) & keep)[0].astype(np.int64)


def count_overlaps(L, R):
    """``(chrom, start, end, n)`` per DISTINCT left key, n = dup(key) * overlaps.

    reference: intersects_duckdb.py:432-548, 806-854; python restatement
    tests/test_duckdb_iejoin.py:66-81.
    """
    cnt = count_per_left(L, R)
    agg = {}
    for i, key in enumerate(zip(L["chrom"], L["start"], L["end"])):
        agg[key] = agg.get(key, 0) + int(cnt[i])
    return sorted((c, int(s), int(e), n) for (c, s, e), n in agg.items())


# --------------------------------------------------------------------------- #
# DISTANCE  (reference: src/giql/expanders/_distance.py:22-117)
# --------------------------------------------------------------------------- #

_NULL_STRANDS = (None, ".", "?")


def distance(a_chrom, a_start, a_end, a_strand, b_chrom, b_start, b_end, b_strand,
             stranded=False, signed=False):
    """bedtools ``closest -d`` style distance; ``None`` is SQL NULL.

    The stranded branch is taken only when *both* strand operands exist
    (_distance.py:67: ``if not stranded or strand_a is None or strand_b is None``
    refers to the *column fragment* being absent, which callers of this oracle
    signal with ``stranded=False``).
    """
    if a_chrom is None or b_chrom is None or a_chrom != b_chrom:
        return None                                     # _distance.py:73 / :92 (NULL != x is NULL -> falls to ... see note)
    if stranded:
        if a_strand in _NULL_STRANDS or b_strand in _NULL_STRANDS:
            return None                                 # _distance.py:93-95
    if a_start < b_end and a_end > b_start:
        return 0                                        # _distance.py:74
    if a_end <= b_start:
        gap = b_start - a_end + 1                       # _distance.py:75 (downstream)
        if stranded and a_strand == "-":
            return -gap                                 # _distance.py:98 / :112
        return gap
    gap = a_start - b_end + 1                           # upstream
    if not stranded:
        return -gap if signed else gap                  # _distance.py:76 / :83
    if signed:
        return gap if a_strand == "-" else -gap         # _distance.py:101-102
    return -gap if a_strand == "-" else gap             # _distance.py:115-116


# --------------------------------------------------------------------------- #
# NEAREST  (reference: src/giql/expanders/nearest.py:255-397)
# --------------------------------------------------------------------------- #

def nearest(Q, T, k=1, max_distance=None, stranded=False, signed=False):
    """Rows ``(query_row, target_row, distance)`` of the LATERAL form.

    For each query row: candidates = target rows on the same chrom
    (nearest.py:327), same strand when ``stranded`` and both sides carry a strand
    column (nearest.py:304-307, 328-329; SQL ``=`` so NULL never matches),
    ``ABS(distance) <= max_distance`` when given (nearest.py:330-331; NULL
    distance fails the filter); order ``ABS(distance), start, end``
    (nearest.py:391-395) with NULLS FIRST (duckdb serialisation,
    targets.py:108-118); first ``k``.  Ties on all three keys are order-ambiguous
    in the reference itself (nearest.py:366-372); this oracle breaks them by
    target row id, callers must compare modulo that residual.
    """
    has_strand = stranded and ("strand" in Q) and ("strand" in T)
    out = []
    nT = len(T["start"])
    for qi in range(len(Q["start"])):
        qc, qs, qe = Q["chrom"][qi], Q["start"][qi], Q["end"][qi]
        qst = Q["strand"][qi] if has_strand else None
        cands = []
        for ti in range(nT):
            tc, ts, te = T["chrom"][ti], T["start"][ti], T["end"][ti]
            if qc is None or tc != qc:
                continue
            tst = T["strand"][ti] if has_strand else None
            if has_strand and (qst is None or tst is None or qst != tst):
                continue
            d = distance(qc, qs, qe, qst, tc, ts, te, tst, stranded=has_strand, signed=signed)
            if max_distance is not None and (d is None or abs(d) > max_distance):
                continue
            key = ((0, 0) if d is None else (1, abs(d)), ts, te, ti)
            cands.append((key, ti, d))
        cands.sort(key=lambda c: c[0])
        for _, ti, d in cands[:k]:
            out.append((qi, ti, d))
    return out


# --------------------------------------------------------------------------- #
# CLUSTER / MERGE  (reference: src/giql/expanders/cluster.py:205-310, 430-434;
#                   src/giql/expanders/merge.py:183-439)
# --------------------------------------------------------------------------- #

_PRED_OPS = {"=": lambda a, b: a == b, "!=": lambda a, b: a != b, "<>": lambda a, b: a != b, "<": lambda a, b: a < b,
             "<=": lambda a, b: a <= b, ">": lambda a, b: a > b, ">=": lambda a, b: a >= b}


def cluster(T, distance=0, stranded=False, predicate=None):
    """1-based cluster ids aligned to input rows; ids restart per partition.

    ``predicate`` = (cur_values, op, prev_values): the row stays in the running cluster only when it is adjacent AND
    ``cur[row] op prev[immediate sorted predecessor]`` is TRUE — ``PREV(col)`` is ``LAG(col)`` over the same
    partition / order (cluster.py:281-299, 587-644); NULL operands (and the first row's NULL LAG) make it not TRUE.

    flag_i = 0 iff MAX(end) over strictly-preceding rows of the partition
    (+ distance when distance > 0, cluster.py:266-272) >= start_i, else 1 (first
    row: window empty -> NULL -> 1; cluster.py:254-279).  id = SUM(flag) with the
    default RANGE frame, so peers on equal ``start`` share the sum
    (cluster.py:430-434).  Raw table coordinates, no canonicalisation
    (expressions.py:223-226).  Partition = chrom (+ strand when stranded,
    cluster.py:243-245; NULL strands group together under PARTITION BY).
    """
    n = len(T["start"])
    ids = [0] * n
    parts = {}
    for i in range(n):
        key = (T["chrom"][i], T["strand"][i] if stranded else None)
        parts.setdefault(key, []).append(i)
    d = distance if distance and distance > 0 else 0
    for rows in parts.values():
        rows = sorted(rows, key=lambda i: (T["start"][i], i))
        flags = []
        run_max = None
        before = None
        for i in rows:
            keep = run_max is not None and run_max + d >= T["start"][i]
            if keep and predicate is not None:
                a, b = predicate[0][i], predicate[2][before]
                keep = a is not None and b is not None and bool(_PRED_OPS[predicate[1]](a, b))
            flags.append(0 if keep else 1)
            run_max = T["end"][i] if run_max is None else max(run_max, T["end"][i])
            before = i
        # RANGE frame: peers (equal start) see the sum through the last peer
        csum, acc = [], 0
        for f in flags:
            acc += f
            csum.append(acc)
        j = 0
        while j < len(rows):
            k2 = j
            while k2 + 1 < len(rows) and T["start"][rows[k2 + 1]] == T["start"][rows[j]]:
                k2 += 1
            for m in range(j, k2 + 1):
                ids[rows[m]] = csum[k2]
            j = k2 + 1
    return ids


def merge(T, distance=0, stranded=False, predicate=None):
    """Merged rows ``(chrom[, strand], start, end, count)`` ordered by (chrom, start).

    reference: merge.py:216-239 (cluster), :293-308 (GROUP BY chrom[,strand],
    cluster id -> MIN(start), MAX(end)), :429-437 (default ORDER BY chrom, start).
    ``count`` is COUNT(*) per region (user aggregate commonly requested,
    tests/integration/datafusion/test_cross_target_oracle.py:1196-1315).
    """
    ids = cluster(T, distance, stranded, predicate)
    agg = {}
    for i, cid in enumerate(ids):
        key = (T["chrom"][i], T["strand"][i] if stranded else None, cid)
        s, e = int(T["start"][i]), int(T["end"][i])
        if key in agg:
            a = agg[key]
            a[0] = min(a[0], s)
            a[1] = max(a[1], e)
            a[2] += 1
        else:
            agg[key] = [s, e, 1]
    rows = []
    for (c, st, _cid), (s, e, cnt) in agg.items():
        rows.append((c, st, s, e, cnt) if stranded else (c, s, e, cnt))
    return sorted(rows, key=lambda r: tuple((v is None, v) for v in r))


# --------------------------------------------------------------------------- #
# DISJOIN  (reference: src/giql/expanders/disjoin.py:102-202)
# --------------------------------------------------------------------------- #

def disjoin(T, REF=None):
    """Rows ``(target_row, seg_start, seg_end)`` in canonical coordinates.

    Breakpoints = distinct ref starts/ends per chrom (disjoin.py:149-154); cuts
    of a target key (chrom,start,end) = {start, end} ∪ {bp: start < bp < end}
    (disjoin.py:159-172); segments = consecutive cuts (LEAD, :173-178), kept when
    seg_end > seg_start and, unless the reference is the target itself, some ref
    row has r.start <= seg_start < r.end (:184-190).  Duplicate target rows share
    a key and each gets every segment (join back on the key, :191-199).
    """
    self_ref = REF is None
    ref = T if self_ref else REF
    bps = {}
    for c, s, e in zip(ref["chrom"], ref["start"], ref["end"]):
        bps.setdefault(c, set()).update((int(s), int(e)))
    out = []
    for i, (c, s, e) in enumerate(zip(T["chrom"], T["start"], T["end"])):
        s, e = int(s), int(e)
        cuts = sorted({s, e} | {p for p in bps.get(c, ()) if s < p < e})
        for a, b in zip(cuts[:-1], cuts[1:]):
            if not b > a:
                continue
            if not self_ref:
                covered = any(rc == c and rs <= a and re > a
                              for rc, rs, re in zip(ref["chrom"], ref["start"], ref["end"]))
                if not covered:
                    continue
            out.append((i, a, b))
    return sorted(out)

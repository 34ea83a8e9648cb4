"""Uniform operator interface over the CPU oracle and the B200 engine, used by the golden-case runner.

A *table* here is a dict ``{"chrom": [...], "start": [...], "end": [...], ["strand": [...]]}`` plus an
optional ``config`` (giql_b200.Table) carried separately.  Row ids are list positions.
"""
from __future__ import annotations

import numpy as np

from giql_b200.table import Table, canonical_offsets


def canonicalise(T, config):
    """dict table in canonical 0-based half-open coordinates (reference: src/giql/canonical.py:16-52)."""
    sadj, eadj = canonical_offsets(config)
    out = dict(T)
    out["start"] = [int(s) + sadj for s in T["start"]]
    out["end"] = [int(e) + eadj for e in T["end"]]
    return out


class OracleBackend:
    name = "oracle"

    def __init__(self):
        from oracle import giql_oracle as go

        self.go = go

    def pairs(self, L, R, lcfg=None, rcfg=None):
        p = self.go.join_pairs(canonicalise(L, lcfg), canonicalise(R, rcfg))
        return sorted((int(a), int(b)) for a, b in p)

    def spatial_pairs(self, L, R, op, lcfg=None, rcfg=None):
        p = self.go.join_pairs(canonicalise(L, lcfg), canonicalise(R, rcfg), op)
        return sorted((int(a), int(b)) for a, b in p)

    def exists(self, L, R, lcfg=None, rcfg=None):
        n = len(L["start"])
        cnt = self.go.count_per_left(canonicalise(L, lcfg), canonicalise(R, rcfg)) if n else np.zeros(0, np.int64)
        return [bool(c > 0) for c in cnt]

    def counts(self, L, R, lcfg=None, rcfg=None):
        n = len(L["start"])
        cnt = self.go.count_per_left(canonicalise(L, lcfg), canonicalise(R, rcfg)) if n else np.zeros(0, np.int64)
        return [int(c) for c in cnt]

    def distance(self, A, B, acfg=None, bcfg=None, stranded=False, signed=False):
        A, B = canonicalise(A, acfg), canonicalise(B, bcfg)
        use = stranded and "strand" in A and "strand" in B
        out = []
        for i in range(len(A["start"])):
            out.append(self.go.distance(A["chrom"][i], A["start"][i], A["end"][i], A["strand"][i] if use else None,
                                        B["chrom"][i], B["start"][i], B["end"][i], B["strand"][i] if use else None,
                                        stranded=use, signed=signed))
        return out

    def nearest(self, Q, T, qcfg=None, tcfg=None, k=1, max_distance=None, stranded=False, signed=False):
        return self.go.nearest(canonicalise(Q, qcfg), canonicalise(T, tcfg), k=k, max_distance=max_distance,
                               stranded=stranded, signed=signed)

    def cluster(self, T, distance=0, stranded=False):
        return self.go.cluster(T, distance, stranded)

    def merge(self, T, distance=0, stranded=False):
        return self.go.merge(T, distance, stranded)

    def disjoin(self, T, REF=None, tcfg=None, rcfg=None):
        return self.go.disjoin(canonicalise(T, tcfg), None if REF is None else canonicalise(REF, rcfg))


class EngineBackend:
    """The product path: giql_b200.Engine -> ctypes -> libgiqlb200.so -> sm_100a kernels."""

    name = "b200"

    def __init__(self, engine, join_mode="auto"):
        self.e = engine
        self.join_mode = join_mode

    def _table(self, T, cfg):
        cfg = cfg or Table("t")
        data = {cfg.chrom_col: T["chrom"], cfg.start_col: T["start"], cfg.end_col: T["end"]}
        if "strand" in T and cfg.strand_col:
            data[cfg.strand_col] = T["strand"]
        return self.e.table(data, cfg)

    def pairs(self, L, R, lcfg=None, rcfg=None):
        res = self.e.intersects(self._table(L, lcfg), self._table(R, rcfg), mode=self.join_mode)
        l, r = res.to_host()
        return sorted(zip(l.tolist(), r.tolist()))

    def spatial_pairs(self, L, R, op, lcfg=None, rcfg=None):
        res = (self.e.contains if op == "contains" else self.e.within)(self._table(L, lcfg), self._table(R, rcfg))
        l, r = res.to_host()
        return sorted(zip(l.tolist(), r.tolist()))

    def range_filter(self, t, op, ranges, cfg):
        """literal-range CONTAINS / WITHIN filter = the column join against the one-row table of the literal"""
        lit = {"chrom": [r[0] for r in ranges], "start": [r[1] for r in ranges], "end": [r[2] for r in ranges]}
        hit = {i for i, _ in self.spatial_pairs(t, lit, op, cfg, None)}
        return [i in hit for i in range(len(t["start"]))]

    def exists(self, L, R, lcfg=None, rcfg=None):
        return [bool(x) for x in self.e.exists(self._table(L, lcfg), self._table(R, rcfg)).cpu().tolist()]

    def counts(self, L, R, lcfg=None, rcfg=None):
        return [int(x) for x in self.e.count_overlaps(self._table(L, lcfg), self._table(R, rcfg)).cpu().tolist()]

    def distance(self, A, B, acfg=None, bcfg=None, stranded=False, signed=False):
        d, v = self.e.distance(self._table(A, acfg), self._table(B, bcfg), stranded=stranded, signed=signed)
        return [int(x) if ok else None for x, ok in zip(d.cpu().tolist(), v.cpu().tolist())]

    def nearest(self, Q, T, qcfg=None, tcfg=None, k=1, max_distance=None, stranded=False, signed=False):
        res = self.e.nearest(self._table(Q, qcfg), self._table(T, tcfg), k=k, max_distance=max_distance,
                             stranded=stranded, signed=signed)
        return res.rows()

    def cluster(self, T, distance=0, stranded=False):
        return [int(x) for x in self.e.cluster(self._table(T, None), distance=distance, stranded=stranded).cpu().tolist()]

    def merge(self, T, distance=0, stranded=False):
        return self.e.merge(self._table(T, None), distance=distance, stranded=stranded).rows()

    def disjoin(self, T, REF=None, tcfg=None, rcfg=None):
        res = self.e.disjoin(self._table(T, tcfg), None if REF is None else self._table(REF, rcfg))
        return res.rows()

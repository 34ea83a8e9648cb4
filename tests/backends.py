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

    def cluster(self, T, distance=0, stranded=False, predicate=None):
        return self.go.cluster(T, distance, stranded, predicate)

    def merge(self, T, distance=0, stranded=False, predicate=None):
        return self.go.merge(T, distance, stranded, predicate)

    def disjoin(self, T, REF=None, tcfg=None, rcfg=None):
        return self.go.disjoin(canonicalise(T, tcfg), None if REF is None else canonicalise(REF, rcfg))


from giql_b200.frontend import EngineBackend  # noqa: E402,F401  (the product backend lives in the package)

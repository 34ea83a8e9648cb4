"""Runs the reference's cross-target golden triples (query, data, expected) against a backend.

The fixtures in ``tests/golden/cross_target_oracle.json`` were harvested verbatim from abdenlab/giql
``tests/integration/datafusion/test_cross_target_oracle.py`` (see harvest_reference_vectors.py).  sqlglot
is absent from this image, so instead of ``giql.transpile`` the product's own front-end
(giql_b200/frontend.py: parser + binder) maps each query onto operator calls of a backend — the CUDA engine, or
the CPU oracle plugged into the same seat by tests/backends.py — and evaluates the SELECT list on the host.  Results are
compared the way the reference does: sorted multisets of tuples (tests/integration/_oracle.py:77-89).
"""
from __future__ import annotations

import json
from pathlib import Path

from giql_b200.table import Table

GOLDEN = Path(__file__).resolve().parent / "golden"
DEFAULT_COLUMNS = ["chrom", "start", "end"]





def load_cases():
    return json.loads((GOLDEN / "cross_target_oracle.json").read_text())


def normalise(rows):
    key = lambda row: tuple((v is None, 0 if v is None else (v if not isinstance(v, str) else 0), "" if not isinstance(v, str) else v) for v in row)
    return sorted((tuple(r) for r in rows), key=key)


from giql_b200.frontend import Session, UnsupportedQuery as Unsupported  # noqa: E402
from giql_b200.ranges import parse_range  # noqa: E402,F401  (re-exported for the range-literal tests)


def session_for(case, backend):
    cols = [c[0] for c in case["columns"]] if case.get("columns") else DEFAULT_COLUMNS
    cfgs = {spec["name"]: Table(**spec) for spec in case.get("tables") or []}
    sess = Session(backend)
    for name, rows in case["data"].items():
        sess.register(name, {c: [r[i] for r in rows] for i, c in enumerate(cols)}, cfgs.get(name), columns=cols)
    return sess


def run_case(case, backend):
    """Returns the rows the query produces through ``backend`` (list of lists)."""
    return [list(r) for r in session_for(case, backend).execute(case["query"]).rows]

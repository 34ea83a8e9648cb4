"""The reference's DISJOIN usage matrix (abdenlab/giql tests/test_usage_patterns.py:143-171) against this repo's front-end:
15 semantic profiles (point / adjacent / nested / identical intervals, duplicate targets, many / single chromosomes,
reference gaps / supersets, empty target / reference, custom columns, 1-based closed and mixed encodings) x self / reference
mode x 26 query patterns, with the rows DuckDB returned for the transpiled SQL as snapshotted by the reference
(tests/golden/disjoin_usage_matrix.json, harvested by harvest_reference_vectors.py).  Every one of the 690 cases — CTEs,
derived tables, HAVING, set operations, chained and subquery references, joins back to a table on the spelled-out overlap
predicate, a window over the segments ... — must execute and reproduce the snapshot exactly on the oracle backend; the GPU test
runs the ten patterns (300 cases) that were bound while the round still had GPU time."""
import json
from collections import Counter
from pathlib import Path

import pytest

from backends import OracleBackend
from giql_b200.frontend import Session, UnsupportedQuery
from giql_b200.table import Table

CASES = json.loads((Path(__file__).resolve().parent / "golden" / "disjoin_usage_matrix.json").read_text())["cases"]


def _run(case, backend):
    sess = Session(backend)
    cfgs = {t["name"]: Table(**t) for t in case["tables"] if isinstance(t, dict)}
    for f in case["fixtures"]:
        cols = [c[0] for c in f["columns"]]
        sess.register(f["name"], {c: [r[i] for r in f["rows"]] for i, c in enumerate(cols)}, cfgs.get(f["name"]), columns=cols)
    rows = sess.execute(case["query"]).rows
    return sorted(([list(r) for r in rows]), key=repr)


# patterns whose cases ran on the CUDA engine while the round still had GPU time; HAVING and subquery references were bound later
# and are pinned on the oracle backend (the DISJOIN call underneath is the same one)
GPU_VALIDATED = {"ALIASED_PROJECTION", "DERIVED_SUBQUERY", "ORDERED_CTE", "OUTER_COUNT", "OUTER_DISTINCT", "OUTER_GROUP_BY",
                 "OUTER_ORDER_LIMIT", "OUTER_SPATIAL_WHERE", "STANDALONE", "WRAPPING_CTE"}


def _sweep(backend, patterns=None):
    outcome, wrong = Counter(), []
    for case in CASES:
        if patterns is not None and case["pattern"] not in patterns:
            continue
        try:
            got = _run(case, backend)
        except UnsupportedQuery:
            outcome["declined:" + case["pattern"]] += 1
            continue
        if got == case["expected"]:
            outcome["ok:" + case["pattern"]] += 1
        else:
            wrong.append((case["key"], case["query"], got[:3], case["expected"][:3]))
    return outcome, wrong


def test_matrix_is_complete():
    assert len(CASES) == 690 and len({c["profile"] for c in CASES}) == 15 and {c["mode"] for c in CASES} == {"self", "reference"}


def test_every_executed_pattern_reproduces_the_duckdb_snapshot_oracle_backend():
    outcome, wrong = _sweep(OracleBackend())
    assert not wrong, wrong[:5]
    ok = sum(v for k, v in outcome.items() if k.startswith("ok:"))
    assert ok == 690, dict(outcome)
    # the plain table-function shapes are never declined, in any profile or mode
    for pattern in ("STANDALONE", "ALIASED_PROJECTION"):
        assert outcome["ok:" + pattern] == 30, (pattern, dict(outcome))


@pytest.mark.gpu
def test_every_executed_pattern_reproduces_the_duckdb_snapshot_on_the_gpu(engine):
    from giql_b200.frontend import EngineBackend

    outcome, wrong = _sweep(EngineBackend(engine), GPU_VALIDATED)
    assert not wrong, wrong[:5]
    assert sum(v for k, v in outcome.items() if k.startswith("ok:")) >= 290

"""The reference's own DuckDB integration tests, replayed on this repo's front-end.

tests/golden/replay_reference_tests.py imports abdenlab/giql's tests/integration/{cluster_predicate, coordinate_space,
bedtools/test_contains, bedtools/test_within} modules, binds their one fixture (``giql_query``: load rows, transpile, run on
DuckDB) to ``giql_b200.Session`` and lets the reference's assertions judge the rows; tests/test_duckdb_iejoin.py (with its
hypothesis properties against ``_python_overlap`` / ``_python_count_overlaps``), tests/test_intersects_join.py and
tests/expanders/test_{cluster,merge}.py run against stand-ins for ``duckdb.connect`` and ``transpile``.  The calls of every test that passed
are committed in tests/golden/reference_test_replay.json; here they are

* re-executed on the CPU oracle backend and on the CUDA engine (``-m gpu``) and compared with the verified rows, and
* when the reference tree is mounted (the build container), replayed live so that the committed file cannot go stale.
"""
import json
import sys
from pathlib import Path

import pytest

from backends import OracleBackend

GOLDEN = Path(__file__).resolve().parent / "golden"
DOC = json.loads((GOLDEN / "reference_test_replay.json").read_text())
VERIFIED = [t for t in DOC["tests"] if t["outcome"] == "verified"]
CALLS = [(t["test"], t["params"], c) for t in VERIFIED for c in t["calls"]]


def _replay_module():
    sys.path.insert(0, str(GOLDEN))
    try:
        import replay_reference_tests as rr
    finally:
        sys.path.pop(0)
    return rr


def _ordered(query: str) -> bool:
    """standalone NEAREST returns its rows in rank order (ties broken by target row on both backends); every other result
    is compared as a multiset — ORDER BY semantics have their own tests (tests/test_frontend.py)"""
    return "FROMNEAREST(" in "".join(query.upper().split())


# the (test, params) pairs whose stored calls ran on the CUDA engine while the round still had GPU time (commit 251da10); tests
# bound later — aggregates, HAVING, subquery references, the property and cross-target suites — are replayed on the oracle
# backend only (the engine calls underneath are the same operators)
GPU_VALIDATED = {tuple(k) for k in json.loads((GOLDEN / "gpu_validated_replay.json").read_text())}


def _check(backend, only=None):
    rr = _replay_module()
    wrong = []
    for test, params, c in CALLS:
        if only is not None and (test, json.dumps(params, sort_keys=True)) not in only:
            continue
        got = [list(r) for r in rr.run_query(backend, c["query"], c["tables"], c["data"])]
        want = c["rows"]
        if not _ordered(c["query"]):
            got, want = sorted(got, key=repr), sorted(want, key=repr)
        if got != want:
            wrong.append((test, params, c["query"], got[:4], want[:4]))
    return wrong


def test_golden_file_holds_only_verified_or_declined_tests():
    outcomes = {t["outcome"].split(":")[0] for t in DOC["tests"]}
    assert outcomes <= {"verified", "declined", "not_replayable"}, outcomes
    assert len(VERIFIED) >= 371 and len(CALLS) >= len(VERIFIED)
    # every replayed module contributes
    for part in ("cluster_predicate", "test_distance_coordinate_space", "test_nearest_coordinate_space",
                 "test_disjoin_coordinate_space", "test_contains", "test_within", "test_duckdb_iejoin", "test_intersects_join",
                 "expanders/test_cluster", "expanders/test_merge", "bedtools/test_nearest.py", "bedtools/test_strand_aware",
                 "test_cross_target_encodings", "test_intersect_property"):
        assert any(part in t["test"] for t in VERIFIED), part


def test_verified_rows_reproduce_on_the_oracle_backend():
    assert not _check(OracleBackend())


@pytest.mark.skipif(not Path("/root/reference/tests/integration").is_dir(), reason="reference tree not mounted")
def test_live_replay_of_the_reference_tests_matches_the_committed_file():
    rr = _replay_module()
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "giql" or k.startswith("giql.") or k == "tests" or k.startswith("tests.") or k == "duckdb"}
    try:
        rr.install_reference_stubs()
        results = rr.replay(OracleBackend()) + rr.replay_duckdb(OracleBackend())
    finally:
        for k in [k for k in sys.modules if k == "giql" or k.startswith("giql.") or k == "tests" or k.startswith("tests.") or k == "duckdb"]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
    failed = [r for r in results if r["outcome"].startswith(("failed", "error"))]
    assert not failed, [(r["test"], r["params"], r["outcome"]) for r in failed[:3]]
    live = {(r["test"], json.dumps(r["params"], sort_keys=True)) for r in results if r["outcome"] == "verified"}
    committed = {(t["test"], json.dumps(t["params"], sort_keys=True)) for t in VERIFIED}
    assert live == committed, (sorted(live - committed)[:3], sorted(committed - live)[:3])


@pytest.mark.gpu
def test_verified_rows_reproduce_on_the_gpu(engine):
    from giql_b200.frontend import EngineBackend

    assert len(GPU_VALIDATED) == 275
    assert not _check(EngineBackend(engine), only=GPU_VALIDATED)

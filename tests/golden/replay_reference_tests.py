#!/usr/bin/env python
"""Run the reference's OWN integration tests against this repo's front-end, and keep what they verified as golden vectors.

The reference's DuckDB-backed integration tests call one fixture, ``giql_query(query, tables=[...], **rows_per_table)``
(tests/integration/cluster_predicate/conftest.py:27-47, tests/integration/coordinate_space/conftest.py:32-50,
tests/integration/bedtools/conftest.py:34-61), and assert on the rows it returns.  This script imports those test
modules from the mounted reference tree, calls every test function whose only fixture is ``giql_query`` with a stand-in
that executes the query on ``giql_b200.Session`` (CPU oracle backend) instead of transpile + DuckDB, and lets the
reference's own assertions judge the rows.  Outcomes per test: ``verified`` (all assertions passed), ``declined`` (the
front-end raised UnsupportedQuery: a shape outside SURVEY Appendix C), ``failed`` (an assertion failed — a semantic bug
here; the script exits non-zero).  Every call of a verified test is written to ``reference_test_replay.json`` as
(query, tables, data, rows); the GPU test replays them on the CUDA engine without the reference tree.

    python tests/golden/replay_reference_tests.py            # in the build container (/root/reference mounted)

Modules replayed: cluster_predicate/test_cluster_predicate.py, coordinate_space/test_{distance,nearest,disjoin}_
coordinate_space.py, bedtools/test_contains.py, bedtools/test_within.py.  The other bedtools tests compare with the
``bedtools`` binary at run time (absent here) and are not replayable.  ``giql`` itself is not importable (sqlglot is
absent): ``giql.table`` / ``giql.constants`` — sqlglot-free — are loaded from the reference tree by path.
"""
from __future__ import annotations

import importlib
import importlib.util
import inspect
import itertools
import json
import sys
import types
from pathlib import Path

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
OUT = HERE / "reference_test_replay.json"
MODULES = [
    "tests.integration.cluster_predicate.test_cluster_predicate",
    "tests.integration.coordinate_space.test_distance_coordinate_space",
    "tests.integration.coordinate_space.test_nearest_coordinate_space",
    "tests.integration.coordinate_space.test_disjoin_coordinate_space",
    "tests.integration.bedtools.test_contains",
    "tests.integration.bedtools.test_within",
]
COLUMNS = ["chrom", "start", "end", "name", "score", "strand"]      # bedtools/utils/duckdb_loader.py:18-27


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def install_reference_stubs():
    """``giql.table`` / ``giql.constants`` by path, and package shells so the test modules' relative imports resolve."""
    giql = types.ModuleType("giql")
    giql.__path__ = []
    sys.modules["giql"] = giql
    giql.constants = _load("giql.constants", REF / "src/giql/constants.py")
    giql.table = _load("giql.table", REF / "src/giql/table.py")
    giql.Table = giql.table.Table

    def transpile(*a, **k):
        raise RuntimeError("replay: tests that transpile directly are not replayable without sqlglot")

    giql.transpile = transpile
    for pkg, path in [("tests", REF / "tests"), ("tests.integration", REF / "tests/integration"),
                      ("tests.integration.cluster_predicate", REF / "tests/integration/cluster_predicate"),
                      ("tests.integration.coordinate_space", REF / "tests/integration/coordinate_space"),
                      ("tests.integration.bedtools", REF / "tests/integration/bedtools"),
                      ("tests.integration.bedtools.utils", REF / "tests/integration/bedtools/utils")]:
        m = types.ModuleType(pkg)
        m.__path__ = [str(path)]
        sys.modules[pkg] = m


def _table_spec(t) -> dict:
    if isinstance(t, str):
        return {"name": t}
    return {k: getattr(t, k) for k in ("name", "genomic_col", "chrom_col", "start_col", "end_col", "strand_col",
                                       "coordinate_system", "interval_type")}


def run_query(backend, query: str, tables: list[dict], data: dict) -> list[tuple]:
    """What the reference's fixture does with DuckDB, on this repo's front-end."""
    from giql_b200.frontend import Session
    from giql_b200.table import Table

    sess = Session(backend)
    cfgs = {t["name"]: Table(**t) for t in tables if len(t) > 1}
    for name, rows in data.items():
        sess.register(name, {c: [r[i] for r in rows] for i, c in enumerate(COLUMNS)}, cfgs.get(name), columns=COLUMNS)
    return [tuple(r) for r in sess.execute(query).rows]


class Recorder:
    def __init__(self, backend):
        self.backend, self.calls = backend, []

    def __call__(self, query, *, tables, **table_data):
        specs = [_table_spec(t) for t in tables]
        data = {name: [list(r.to_tuple() if hasattr(r, "to_tuple") else r) for r in rows] for name, rows in table_data.items()}
        rows = run_query(self.backend, query, specs, data)
        self.calls.append({"query": query, "tables": specs, "data": data, "rows": [list(r) for r in rows]})
        return rows


def _parametrizations(fn, cls):
    marks = list(getattr(fn, "pytestmark", []))
    axes = []
    for m in marks:
        if m.name != "parametrize":
            continue
        names, values = m.args[0], m.args[1]
        names = [n.strip() for n in names.split(",")] if isinstance(names, str) else list(names)
        vals = []
        for v in values:
            v = getattr(v, "values", v)                    # pytest.param(...)
            vals.append(dict(zip(names, v if len(names) > 1 else [v])))
        axes.append(vals)
    if not axes:
        return [{}]
    return [{k: v for d in combo for k, v in d.items()} for combo in itertools.product(*axes)]


def replay(backend):
    from giql_b200.frontend import UnsupportedQuery

    results = []
    for modname in MODULES:
        mod = importlib.import_module(modname)
        src = modname.replace(".", "/") + ".py"
        owners = [(None, mod)] + [(c, c) for n, c in vars(mod).items() if inspect.isclass(c) and n.startswith("Test")]
        for cls, holder in owners:
            for name, fn in vars(holder).items():
                if not (name.startswith("test_") and inspect.isfunction(fn)):
                    continue
                for params in _parametrizations(fn, cls):
                    fixtures = [p for p in inspect.signature(fn).parameters if p not in ("self", *params)]
                    entry = {"test": f"{src}::{cls.__name__ + '::' if cls else ''}{name}", "params": {k: repr(v) for k, v in params.items()},
                             "line": fn.__code__.co_firstlineno}
                    if fixtures != ["giql_query"]:
                        entry["outcome"] = "skipped: fixtures " + ",".join(fixtures)
                        results.append(entry)
                        continue
                    rec = Recorder(backend)
                    try:
                        if cls:
                            fn(cls(), giql_query=rec, **params)
                        else:
                            fn(giql_query=rec, **params)
                        entry["outcome"], entry["calls"] = "verified", rec.calls
                    except UnsupportedQuery as e:
                        entry["outcome"] = f"declined: {e}"
                    except AssertionError as e:
                        entry["outcome"], entry["calls"] = f"failed: {str(e)[:300]}", rec.calls
                    results.append(entry)
    return results


def main():
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    install_reference_stubs()
    from backends import OracleBackend

    results = replay(OracleBackend())
    tally = {}
    for r in results:
        key = r["outcome"].split(":")[0]
        tally[key] = tally.get(key, 0) + 1
    for r in results:
        if not r["outcome"].startswith("verified"):
            print(f"{r['outcome'][:150]:150s}  {r['test']} {r['params'] or ''}")
    print(tally)
    doc = {"source": "abdenlab/giql tests/integration/{cluster_predicate,coordinate_space,bedtools}: test functions executed by "
                     "tests/golden/replay_reference_tests.py with giql_query bound to giql_b200.Session (oracle backend); rows of "
                     "'verified' tests passed every assertion of the reference's test body",
           "tally": tally, "tests": results}
    OUT.write_text(json.dumps(doc, indent=1, default=str) + "\n")
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    return 1 if tally.get("failed") else 0


if __name__ == "__main__":
    sys.exit(main())

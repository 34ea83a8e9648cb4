"""The TRUE reference path, when the box has it: GIQL's transpiled SQL executed by DuckDB.
TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline / --impl reference legs); never imported by giql_b200/.

``duckdb`` and ``sqlglot`` (hence ``giql``) are absent from the build image and the GPU boxes, so everything here is
guarded by run-time probes:

    available()   duckdb importable
    have_giql()   giql importable too -> the SQL comes from giql.transpile itself (both plans: dialect=None naive ON
                  predicate, dialect="duckdb" per-chromosome IEJoin, docs/transpilation/performance.rst:355-381)
                  otherwise the naive predicate is restated by hand (src/giql/expanders/intersects.py:149-154) and DuckDB
                  picks its own range-join plan; the label says "restated SQL".
"""
from __future__ import annotations

import time

import numpy as np

QUERY = "SELECT a.rid AS l, b.rid AS r FROM l a JOIN r b ON a.interval INTERSECTS b.interval"
RESTATED = 'SELECT a.rid AS l, b.rid AS r FROM l a JOIN r b ON a.chrom = b.chrom AND a."start" < b."end" AND a."end" > b."start"'


def available() -> bool:
    try:
        import duckdb  # noqa: F401
        return True
    except ImportError:
        return False


def have_giql() -> bool:
    try:
        import giql  # noqa: F401
        return True
    except ImportError:
        return False


def label(threads: int) -> str:
    import duckdb

    return f"duckdb {duckdb.__version__}" + ("" if have_giql() else " (restated SQL)") + f", threads={threads}"


def _connect(A, B, threads: int):
    import duckdb
    import pyarrow as pa

    conn = duckdb.connect(":memory:")
    conn.execute(f"SET threads TO {int(threads)}")
    for name, T in (("l", A), ("r", B)):
        names = pa.array(list(T["names"]), type=pa.string())
        chrom = pa.DictionaryArray.from_arrays(pa.array(T["code"], type=pa.int32()), names).cast(pa.string())
        tab = pa.table({"chrom": chrom, "start": pa.array(T["start"], type=pa.int32()), "end": pa.array(T["end"], type=pa.int32()),
                        "rid": pa.array(np.arange(len(T["start"]), dtype=np.int32))})
        conn.register(f"{name}_arrow", tab)
        conn.execute(f'CREATE TABLE {name} AS SELECT * FROM {name}_arrow')
    return conn


def statements(plan: str) -> list[str]:
    """The statements to run, last one being the final SELECT (the IEJoin plan sets a session variable first)."""
    if have_giql():
        import giql

        sql = giql.transpile(QUERY, tables=["l", "r"], dialect="duckdb" if plan == "iejoin" else None)
    else:
        sql = RESTATED
    parts = [p.strip() for p in sql.split(";") if p.strip()]
    return parts


def intersects_join_timed(A, B, threads: int, plan: str = "iejoin", repeats: int = 3):
    """(pairs, best seconds, sorted packed pairs) of the INTERSECTS inner join materialised with CREATE TEMP TABLE out AS ...;
    wall clock, best of ``repeats`` after one warm-up."""
    conn = _connect(A, B, threads)
    parts = statements(plan)
    best = float("inf")
    for it in range(repeats + 1):
        conn.execute("DROP TABLE IF EXISTS out")
        t0 = time.perf_counter()
        for stmt in parts[:-1]:
            conn.execute(stmt)                       # SET VARIABLE ... lives on this connection
        conn.execute("CREATE TEMP TABLE out AS " + parts[-1])
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    l, r = (np.asarray(c, dtype=np.int64) for c in conn.execute("SELECT l, r FROM out").fetchnumpy().values())
    packed = np.sort((l << 32) | r)
    return int(len(packed)), best, packed

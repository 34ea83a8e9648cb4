"""The host half of the reference-side plug-in on the GPU: plans of the shape giql_plugin's expanders emit are executed by the
engine, their result relations registered on a SQL connection (stdlib sqlite3 stands in for DuckDB, which is absent from the
image) and the residual SQL finished there; results equal the same statement evaluated by the SQL engine alone."""
import json
import sqlite3

import numpy as np
import pytest

from giql_b200 import giql_plugin as gp
from giql_b200 import synth

pytestmark = pytest.mark.gpu


def _conn(**tables):
    conn = sqlite3.connect(":memory:")
    for name, T in tables.items():
        conn.execute(f'CREATE TABLE "{name}" (chrom TEXT, start INTEGER, "end" INTEGER, strand TEXT, score INTEGER)')
        conn.executemany(f'INSERT INTO "{name}" VALUES (?, ?, ?, ?, ?)',
                         zip(T["chrom"].tolist(), T["start"].tolist(), T["end"].tolist(), T["strand"].tolist(), T["score"].tolist()))
    return conn


def _cfg(name, **kw):
    d = {"name": name, "chrom_col": "chrom", "start_col": "start", "end_col": "end", "strand_col": "strand",
         "coordinate_system": "0based", "interval_type": "half_open"}
    d.update(kw)
    return d


@pytest.fixture(scope="module")
def tables():
    A = synth.with_strings(synth.peaks(3000, 11))
    B = synth.with_strings(synth.genes(4000, 12))
    return A, B


NAIVE = 'a.chrom = b.chrom AND a.start < b."end" AND a."end" > b.start'      # src/giql/expanders/intersects.py:149-154


def test_inner_join_plan_then_residual_sql(engine, tables):
    A, B = tables
    conn = _conn(peaks=A, genes=B)
    plan = {"op": "intersects", "kind": "INNER", "into": "__giql_x_0", "left": _cfg("peaks"), "right": _cfg("genes")}
    text = gp.PLAN_PREFIX + json.dumps([plan]) + ";\n" + (
        'SELECT a.rowid, b.rowid, a.score + b.score FROM peaks AS a JOIN "__giql_x_0" ON a.rowid = "__giql_x_0".l '
        'JOIN genes AS b ON b.rowid = "__giql_x_0".r WHERE b.score > 300 ORDER BY 1, 2')
    got = gp.run(conn, text, engine).fetchall()
    want = conn.execute(f"SELECT a.rowid, b.rowid, a.score + b.score FROM peaks a JOIN genes b ON {NAIVE} WHERE b.score > 300 ORDER BY 1, 2").fetchall()
    assert len(want) > 50 and got == want


@pytest.mark.parametrize("kind", ["SEMI", "ANTI"])
def test_semi_anti_plans(engine, tables, kind):
    A, B = tables
    conn = _conn(peaks=A, genes=B)
    plan = {"op": "intersects", "kind": kind, "into": "__giql_x_1", "left": _cfg("peaks"), "right": _cfg("genes")}
    neg = "NOT " if kind == "ANTI" else ""
    text = gp.PLAN_PREFIX + json.dumps([plan]) + f';\nSELECT a.rowid FROM peaks AS a WHERE a.rowid {neg}IN (SELECT l FROM "__giql_x_1") ORDER BY 1'
    got = gp.run(conn, text, engine).fetchall()
    want = conn.execute(f"SELECT a.rowid FROM peaks a WHERE {neg}EXISTS (SELECT 1 FROM genes b WHERE {NAIVE}) ORDER BY 1").fetchall()
    assert 0 < len(want) < len(A["start"]) and got == want


def test_encoded_tables_use_their_declared_coordinates(engine, tables):
    A, B = tables
    A1 = dict(A)
    A1["start"] = A["start"] + 1                     # the same intervals declared 1-based closed: start + 1, end unchanged
    conn = _conn(peaks=A1, genes=B)
    plan = {"op": "intersects", "kind": "INNER", "into": "__giql_x_2",
            "left": _cfg("peaks", coordinate_system="1based", interval_type="closed"), "right": _cfg("genes")}
    text = gp.PLAN_PREFIX + json.dumps([plan]) + ';\nSELECT l, r FROM "__giql_x_2" ORDER BY 1, 2'
    got = gp.run(conn, text, engine).fetchall()
    want = conn.execute('SELECT a.rowid, b.rowid FROM peaks a JOIN genes b ON a.chrom = b.chrom AND a.start - 1 < b."end" AND a."end" > b.start ORDER BY 1, 2').fetchall()
    assert got == want


def test_nearest_cluster_merge_disjoin_plans_match_the_oracle(engine, tables):
    from oracle import giql_oracle as go

    A, B = tables
    n = 600
    Q = {k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in A.items()}
    conn = _conn(peaks=Q, genes=B)
    plans = [
        {"op": "nearest", "into": "nn", "query": _cfg("peaks"), "target": _cfg("genes"), "k": 3, "max_distance": None, "stranded": True, "signed": True},
        {"op": "cluster", "into": "cl", "table": _cfg("peaks"), "distance": 500, "stranded": False},
        {"op": "merge", "into": "mg", "table": _cfg("peaks"), "distance": 0, "stranded": False},
        {"op": "disjoin", "into": "dj", "target": _cfg("genes"), "reference": _cfg("peaks")},
    ]
    gp.run(conn, gp.PLAN_PREFIX + json.dumps(plans) + ";\nSELECT 1", engine)
    Qd = {"chrom": Q["chrom"].tolist(), "start": Q["start"].tolist(), "end": Q["end"].tolist(), "strand": Q["strand"].tolist()}
    Bd = {"chrom": B["chrom"].tolist(), "start": B["start"].tolist(), "end": B["end"].tolist(), "strand": B["strand"].tolist()}
    got = conn.execute("SELECT q - 1, distance FROM nn ORDER BY q, abs(distance)").fetchall()      # sqlite rowids are 1-based
    want = sorted(((q, d) for q, _t, d in go.nearest(Qd, Bd, k=3, stranded=True, signed=True)), key=lambda r: (r[0], abs(r[1])))
    assert [(q, abs(d)) for q, d in got] == [(q, abs(d)) for q, d in want]
    ids = [r[0] for r in conn.execute("SELECT cluster_id FROM cl ORDER BY row").fetchall()]
    assert ids == [int(x) for x in go.cluster(Qd, 500, False)]
    assert conn.execute("SELECT chrom, start, \"end\", n FROM mg ORDER BY chrom, start").fetchall() == [tuple(r) for r in go.merge(Qd, 0, False)]
    segs = conn.execute("SELECT row - 1, disjoin_start, disjoin_end FROM dj ORDER BY 1, 2").fetchall()
    assert segs == [tuple(int(v) for v in r) for r in go.disjoin(Bd, Qd)]

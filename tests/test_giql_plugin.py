"""The reference-side plug-in (giql_b200/giql_plugin.py): importable without giql, its plan contract, and — where the
reference tree is mounted — that every name it binds from ``giql`` really exists there (round 1's INTEGRATION.md stub
called an API the reference does not have).  With giql + sqlglot importable, the expanders themselves are exercised."""
import ast
import re
import sqlite3
from pathlib import Path

import numpy as np
import pytest

from giql_b200 import giql_plugin as gp

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/src/giql")


def test_module_imports_without_giql_and_says_so():
    if gp.HAVE_GIQL:
        assert gp.B200Target is not None and gp.B200Target().name == "b200"
    else:
        assert gp.B200Target is None
        with pytest.raises(ImportError):
            gp.register_b200()
        with pytest.raises(ImportError):
            gp.transpile("SELECT 1")


def test_plan_header_round_trip():
    plans = [{"op": "intersects", "kind": "INNER", "into": "__giql_x_0", "left": {"name": "a"}, "right": {"name": "b"}}]
    import json

    text = gp.PLAN_PREFIX + json.dumps(plans) + ";\nSELECT a.x FROM a JOIN __giql_x_0 ON a.rowid = __giql_x_0.l"
    got, sql = gp.split(text)
    assert got == plans and sql.startswith("SELECT a.x")
    assert gp.split("SELECT 1") == ([], "SELECT 1")
    with pytest.raises(ValueError):
        gp.split(gp.PLAN_PREFIX + "[]")


def test_table_plan_defaults_and_mapping():
    from giql_b200.table import Table, Tables

    ts = Tables()
    ts.register("v", Table("v", chrom_col="seqid", start_col="pos", end_col="stop", coordinate_system="1based", interval_type="closed"))
    assert gp._table_plan(ts, "v") == {"name": "v", "chrom_col": "seqid", "start_col": "pos", "end_col": "stop", "strand_col": "strand",
                                       "coordinate_system": "1based", "interval_type": "closed"}
    assert gp._table_plan(ts, "unregistered")["chrom_col"] == "chrom"      # reference default columns (intersects_duckdb.py:1181-1188)


def test_register_writes_a_relation_any_dbapi_connection_can_join():
    conn = sqlite3.connect(":memory:")
    gp._register(conn, "__giql_x_0", {"l": np.array([3, 1], dtype=np.int64), "r": np.array([7, 9], dtype=np.int64)})
    assert conn.execute('SELECT l, r FROM "__giql_x_0" ORDER BY l').fetchall() == [(1, 9), (3, 7)]


def _defined_names(path: Path) -> set:
    tree = ast.parse(path.read_text())
    out = set()
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            out.add(node.name)
        elif isinstance(node, ast.Assign):
            out |= {t.id for t in node.targets if isinstance(t, ast.Name)}
        elif isinstance(node, ast.AnnAssign) and isinstance(node.target, ast.Name):
            out.add(node.target.id)
        elif isinstance(node, ast.ImportFrom):
            out |= {a.asname or a.name for a in node.names}
    return out


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted (GPU box)")
def test_every_name_bound_from_giql_exists_in_the_reference():
    src = (ROOT / "giql_b200" / "giql_plugin.py").read_text()
    imports = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.ImportFrom) and n.module and n.module.split(".")[0] == "giql"]
    assert imports, "the plug-in binds nothing from giql?"
    for imp in imports:
        parts = imp.module.split(".")[1:]
        if not parts:
            path = REF / "__init__.py"
        else:
            rel = Path(*parts)
            path = REF / rel.with_suffix(".py") if (REF / rel.with_suffix(".py")).exists() else REF / rel / "__init__.py"
        assert path.exists(), f"{imp.module} is not a module of the reference"
        have = _defined_names(path)
        for a in imp.names:
            assert a.name in have, f"{imp.module}.{a.name} does not exist in the reference ({path})"
    # attributes and methods the expanders call on reference objects
    uses = {
        "expanders/intersects_duckdb.py": ["def from_intersects", "def transform_to_sql", "def _find_target_join", "def is_left_only_join",
                                           "left_user_alias: str", "right_user_alias: str", "left_table: exp.Table", "right_table: exp.Table",
                                           'kind: Literal["INNER", "SEMI", "ANTI"]'],
        "expander.py": ["def add_statement_finalizer", "def alias", "def resolve"],
        "resolver.py": ["def slot", "class ResolvedRef", "class ResolvedInterval"],
        "expanders/nearest.py": ["def _outer_relation", "def _nearest_params"],
        "expanders/_params.py": ["def coerce_bool_param"],
        "table.py": ["def get"],
    }
    for rel, needles in uses.items():
        text = (REF / rel).read_text()
        for needle in needles:
            assert needle in text, f"{rel}: {needle!r} not found"
    # the keys sqlglot stores FROM / WITH under in the version the reference targets
    duck = (REF / "expanders/intersects_duckdb.py").read_text()
    assert 'query.args["from_"]' in duck and 'query.args.get("with_")' in duck
    for key in re.findall(r'args(?:\.get)?\(?\[?"(from_|with_)"', src):
        assert key in ("from_", "with_")


giql = pytest.importorskip("giql", reason="giql / sqlglot not installed in this image") if gp.HAVE_GIQL else None


@pytest.mark.skipif(not gp.HAVE_GIQL, reason="giql / sqlglot not installed in this image")
@pytest.mark.parametrize("query,op,kind", [
    ("SELECT a.name, b.name FROM peaks a JOIN genes b ON a.interval INTERSECTS b.interval", "intersects", "INNER"),
    ("SELECT a.name FROM peaks a SEMI JOIN genes b ON a.interval INTERSECTS b.interval", "intersects", "SEMI"),
    ("SELECT a.name FROM peaks a ANTI JOIN genes b ON a.interval INTERSECTS b.interval", "intersects", "ANTI"),
    ("SELECT a.name, b.name, b.distance FROM peaks a CROSS JOIN LATERAL NEAREST(genes, reference := a.interval, k := 3) b", "nearest", None),
    ("SELECT *, CLUSTER(interval) AS cid FROM peaks", "cluster", None),
    ("SELECT MERGE(interval) FROM peaks", "merge", None),
    ("SELECT * FROM DISJOIN(peaks, reference := genes)", "disjoin", None),
])
def test_b200_dialect_emits_plans_for_supported_shapes(query, op, kind):
    text = gp.transpile(query, tables=["peaks", "genes"])
    plans, sql = gp.split(text)
    assert len(plans) == 1 and plans[0]["op"] == op and plans[0]["into"] in sql
    if kind:
        assert plans[0]["kind"] == kind
    assert "INTERSECTS" not in sql.upper() and "NEAREST(" not in sql.upper()


@pytest.mark.skipif(not gp.HAVE_GIQL, reason="giql / sqlglot not installed in this image")
def test_declined_shapes_are_plain_sql_identical_to_the_duckdb_free_path():
    import giql as g

    q = "SELECT * FROM peaks WHERE interval INTERSECTS 'chr1:1000-2000'"
    assert gp.split(gp.transpile(q, tables=["peaks"]))[0] == []
    q2 = "SELECT a.*, b.* FROM peaks a LEFT JOIN genes b ON a.interval INTERSECTS b.interval"
    assert gp.transpile(q2, tables=["peaks", "genes"]) == g.transpile(q2, tables=["peaks", "genes"], dialect="duckdb")

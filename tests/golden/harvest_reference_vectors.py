#!/usr/bin/env python
"""Harvest the reference's own golden vectors into ``tests/golden/*.json``.

Run in the build container where ``/root/reference`` is mounted (read-only):

    python tests/golden/harvest_reference_vectors.py

Nothing here imports ``giql`` (sqlglot is absent); literal data is read out of the
reference's test modules with ``ast``/regex, and DISTANCE vectors are produced by
*executing the reference's own* ``generate_distance_case``
(src/giql/expanders/_distance.py:22-117, a sqlglot-free string builder) on stdlib
SQLite over a grid of inputs.  The JSON files are committed; the GPU box never
reads /root/reference.
"""
from __future__ import annotations

import ast
import importlib.util
import itertools
import json
import re
import sqlite3
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _lit(node, consts):
    """literal_eval that resolves module-level Name constants."""
    if isinstance(node, ast.Name) and node.id in consts:
        return consts[node.id]
    try:
        return ast.literal_eval(node)
    except Exception:
        return {"__source__": ast.unparse(node)}


def _table_spec(call):
    """Table("peaks", coordinate_system="1based", ...) -> dict."""
    spec = {}
    if call.args:
        spec["name"] = ast.literal_eval(call.args[0])
    for kw in call.keywords:
        spec[kw.arg] = ast.literal_eval(kw.value)
    return spec


def harvest_cross_target():
    path = REF / "tests/integration/datafusion/test_cross_target_oracle.py"
    tree = ast.parse(path.read_text())
    consts = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            try:
                consts[node.targets[0].id] = ast.literal_eval(node.value)
            except Exception:
                pass
    cases = []
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            for node in ast.walk(fn):
                if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Name)
                        and node.func.id == "cross_target_oracle"):
                    continue
                case = {
                    "source": f"tests/integration/datafusion/test_cross_target_oracle.py:{node.lineno}",
                    "class": cls.name,
                    "test": fn.name,
                    "query": _lit(node.args[0], consts),
                    "tables": None,
                    "columns": None,
                    "data": {},
                }
                for kw in node.keywords:
                    if kw.arg == "expected":
                        case["expected"] = _lit(kw.value, consts)
                    elif kw.arg == "tables":
                        specs = []
                        for el in kw.value.elts:
                            if isinstance(el, ast.Call):
                                specs.append(_table_spec(el))
                            else:
                                specs.append({"name": ast.literal_eval(el)})
                        case["tables"] = specs
                    elif kw.arg == "columns":
                        case["columns"] = _lit(kw.value, consts)
                    elif kw.arg in ("targets", "engines"):
                        case[kw.arg] = _lit(kw.value, consts)
                    else:
                        case["data"][kw.arg] = _lit(kw.value, consts)
                cases.append(case)
    return cases


_ROW = re.compile(
    r"SELECT\s+'(?P<chrom>[^']+)'\s+as\s+chrom,\s*(?P<start>-?\d+)\s+as\s+start,\s*(?P<end>-?\d+)\s+as\s+end"
    r"(?:,\s*(?P<strand>'[^']*'|NULL)\s+as\s+strand)?\s*\)\s*(?P<alias>[ab])\b", re.I)
_CALL = re.compile(r"DISTANCE\(a\.interval,\s*b\.interval(?P<args>[^)]*)\)", re.I)
_ASSERT = re.compile(r"assert result(?:\[0\])? (?:== (?P<val>-?\d+)|is (?P<none>None))")


def harvest_distance_known_answers():
    path = REF / "tests/test_distance_udf.py"
    src = path.read_text()
    tree = ast.parse(src)
    lines = src.splitlines()
    cases = []
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            body = "\n".join(lines[fn.lineno - 1: fn.end_lineno])
            if "{" in body.split('"""')[-1] and "{b_start}" in body:
                continue
            rows = {m.group("alias").lower(): m for m in _ROW.finditer(body)}
            call = _CALL.search(body)
            asserts = list(_ASSERT.finditer(body))
            if len(rows) != 2 or call is None or len(asserts) != 1:
                continue
            args = call.group("args").lower()

            def row(m):
                st = m.group("strand")
                has = st is not None
                strand = None if (st is None or st.upper() == "NULL") else st.strip("'")
                return {"chrom": m.group("chrom"), "start": int(m.group("start")), "end": int(m.group("end")),
                        "strand": strand, "has_strand": has}

            a = asserts[0]
            cases.append({
                "source": f"tests/test_distance_udf.py:{fn.lineno}",
                "test": fn.name,
                "a": row(rows["a"]), "b": row(rows["b"]),
                "stranded": bool(re.search(r"stranded\s*:=\s*true", args)),
                "signed": bool(re.search(r"signed\s*:=\s*true", args)),
                "expected": None if a.group("none") else int(a.group("val")),
            })
    return cases


def _load_generate_distance_case():
    spec = importlib.util.spec_from_file_location(
        "_ref_distance", REF / "src/giql/expanders/_distance.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.generate_distance_case


def distance_grid_from_reference():
    """Execute the reference's generate_distance_case on SQLite over a grid."""
    gen = _load_generate_distance_case()
    coords = [(100, 200), (150, 250), (200, 300), (201, 300), (300, 400), (150, 150), (200, 200), (0, 100), (99, 100)]
    strands = ["+", "-", ".", "?", None]
    chroms = [("chr1", "chr1"), ("chr1", "chr2")]
    conn = sqlite3.connect(":memory:")
    out = []
    for stranded, signed in itertools.product((False, True), (False, True)):
        case = gen("a.chrom", "a.s", "a.e", "a.t" if stranded else None,
                   "b.chrom", "b.s", "b.e", "b.t" if stranded else None,
                   stranded=stranded, signed=signed)
        sql = (f"SELECT {case} FROM (SELECT ? AS chrom, ? AS s, ? AS e, ? AS t) a, "
               "(SELECT ? AS chrom, ? AS s, ? AS e, ? AS t) b")
        strand_pairs = list(itertools.product(strands, strands)) if stranded else [(None, None)]
        for (ca, cb), (sa, ea), (sb, eb), (ta, tb) in itertools.product(chroms, coords, coords, strand_pairs):
            val = conn.execute(sql, (ca, sa, ea, ta, cb, sb, eb, tb)).fetchone()[0]
            out.append([ca, sa, ea, ta, cb, sb, eb, tb, stranded, signed, val])
    return {"source": "src/giql/expanders/_distance.py:22-117 generate_distance_case executed on sqlite3 "
                      + sqlite3.sqlite_version,
            "columns": ["a_chrom", "a_start", "a_end", "a_strand", "b_chrom", "b_start", "b_end", "b_strand",
                        "stranded", "signed", "expected"],
            "rows": out}


def harvest_iejoin_fixture():
    """peaks_genes fixture + truth table, tests/test_duckdb_iejoin.py:107-142."""
    path = REF / "tests/test_duckdb_iejoin.py"
    tree = ast.parse(path.read_text())
    out = {"source": "tests/test_duckdb_iejoin.py:107-142"}
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "_EXPECTED_OVERLAPS_PEAKS_GENES":
            out["expected_overlaps"] = ast.literal_eval(node.value)
        if isinstance(node, ast.FunctionDef) and node.name == "peaks_genes":
            for call in ast.walk(node):
                if isinstance(call, ast.Call) and getattr(call.func, "id", "") == "_make_table":
                    out[ast.literal_eval(call.args[1])] = ast.literal_eval(call.args[2])
    return out


def harvest_encodings():
    """encode() table of tests/integration/coordinate_space/encodings.py:15-66 evaluated on a grid."""
    spec = importlib.util.spec_from_file_location(
        "_ref_encodings", REF / "tests/integration/coordinate_space/encodings.py")
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception as e:  # imports giql -> sqlglot
        return {"unavailable": repr(e)}
    return {"unavailable": "not needed"}


def main():
    if not REF.exists():
        sys.exit("/root/reference not mounted; golden vectors are already committed")
    (OUT / "cross_target_oracle.json").write_text(json.dumps(harvest_cross_target(), indent=1))
    (OUT / "distance_known_answers.json").write_text(json.dumps(harvest_distance_known_answers(), indent=1))
    (OUT / "distance_reference_case_grid.json").write_text(json.dumps(distance_grid_from_reference()))
    (OUT / "iejoin_peaks_genes.json").write_text(json.dumps(harvest_iejoin_fixture(), indent=1))
    for f in sorted(OUT.glob("*.json")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()


def harvest_disjoin_usage_matrix():
    """The reference's DISJOIN usage matrix (tests/test_usage_patterns.py:143-171): 15 semantic profiles x 2 modes x 26 query
    patterns, each with the rows DuckDB produced for the transpiled SQL, snapshotted in
    tests/manifests/test_usage_patterns.TestDisjoinUsageMatrix...yaml.  The queries and fixtures are rendered by importing the
    reference's own tests/usage_patterns.py (its only giql import, giql.table, is sqlglot-free and is loaded directly)."""
    import types

    import yaml

    pkg = types.ModuleType("giql")
    pkg.__path__ = [str(REF / "src/giql")]
    sys.modules.setdefault("giql", pkg)
    spec = importlib.util.spec_from_file_location("giql.table", REF / "src/giql/table.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["giql.table"] = mod
    spec.loader.exec_module(mod)
    spec = importlib.util.spec_from_file_location("ref_usage_patterns", REF / "tests/usage_patterns.py")
    up = importlib.util.module_from_spec(spec)
    sys.modules["ref_usage_patterns"] = up
    spec.loader.exec_module(up)
    manifest = yaml.safe_load((REF / "tests/manifests/test_usage_patterns.TestDisjoinUsageMatrix."
                                     "test_disjoin_pattern_executes_to_snapshot_across_profiles.yaml").read_text())
    cases = []
    for profile in up.DISJOIN_PROFILES:
        for mode, usage in (("self", profile.self_usage), ("reference", profile.reference_usage)):
            tables = []
            for t in usage.tables:
                tables.append(t if isinstance(t, str) else {k: getattr(t, k) for k in ("name", "genomic_col", "chrom_col", "start_col", "end_col",
                                                                                         "strand_col", "coordinate_system", "interval_type")})
            fixtures = [{"name": f.name, "columns": [list(c) for c in f.columns], "rows": [list(r) for r in f.rows]} for f in usage.fixtures]
            for pattern in up.templated_patterns(usage):
                key = f"{profile.profile_id}-{mode}-duckdb-{pattern.name}"
                if key not in manifest:
                    continue
                cases.append({"key": key, "profile": profile.profile_id, "mode": mode, "pattern": pattern.name, "xfail": pattern in usage.xfails,
                              "query": up.render(pattern, usage), "tables": tables, "fixtures": fixtures, "expected": manifest[key]})
    (OUT / "disjoin_usage_matrix.json").write_text(json.dumps({
        "source": "tests/test_usage_patterns.py:143-171 + tests/manifests/test_usage_patterns.TestDisjoinUsageMatrix...yaml (duckdb rows)",
        "cases": cases}, indent=0))
    return len(cases)

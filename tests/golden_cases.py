"""Runs the reference's cross-target golden triples (query, data, expected) against a backend.

The fixtures in ``tests/golden/cross_target_oracle.json`` were harvested verbatim from abdenlab/giql
``tests/integration/datafusion/test_cross_target_oracle.py`` (see harvest_reference_vectors.py).  sqlglot
is absent from this image, so instead of ``giql.transpile`` a small pattern matcher maps each query onto
operator calls of a backend (tests/backends.py) and evaluates the SELECT list in Python.  Results are
compared the way the reference does: sorted multisets of tuples (tests/integration/_oracle.py:77-89).
"""
from __future__ import annotations

import json
import re
from pathlib import Path

from giql_b200.table import Table

GOLDEN = Path(__file__).resolve().parent / "golden"
DEFAULT_COLUMNS = ["chrom", "start", "end"]


class Unsupported(Exception):
    pass


def load_cases():
    return json.loads((GOLDEN / "cross_target_oracle.json").read_text())


def normalise(rows):
    key = lambda row: tuple((v is None, 0 if v is None else (v if not isinstance(v, str) else 0), "" if not isinstance(v, str) else v) for v in row)
    return sorted((tuple(r) for r in rows), key=key)


def parse_range(text):
    """'chr1:1000-2000' / 'chr1:[a,b)' / 'chr1:[a,b]' / 'chr1:1500' -> (chrom, start, end) canonical
    (reference: src/giql/range_parser.py:75-188, closed -> end + 1 at :44-62)."""
    text = text.strip().strip("'\"")
    m = re.match(r"^([\w.]+):(\d+)$", text)
    if m:
        return m.group(1), int(m.group(2)), int(m.group(2)) + 1
    m = re.match(r"^([\w.]+):\[(\d+),(\d+)([\)\]])(?::([+-]))?$", text)
    if m:
        s, e = int(m.group(2)), int(m.group(3))
        if s >= e:
            raise ValueError(f"Start must be less than end: {s} >= {e}")
        return m.group(1), s, e + (1 if m.group(4) == "]" else 0)
    m = re.match(r"^([\w.]+):(\d+)-(\d+)(?::([+-]))?$", text)
    if m:
        s, e = int(m.group(2)), int(m.group(3))
        if s >= e:
            raise ValueError(f"Start must be less than end: {s} >= {e}")
        return m.group(1), s, e
    raise ValueError(f"Invalid genomic range format: {text}")


def _tables(case):
    cols = [c[0] for c in case["columns"]] if case.get("columns") else DEFAULT_COLUMNS
    cfgs = {}
    for spec in case.get("tables") or []:
        cfgs[spec["name"]] = Table(**spec)
    out = {}
    for name, rows in case["data"].items():
        t = {c: [r[i] for r in rows] for i, c in enumerate(cols)}
        t["__cols__"] = cols
        t["__rows__"] = rows
        out[name] = t
    return out, cfgs


def _select_items(sel):
    items, depth, cur = [], 0, ""
    for ch in sel:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            items.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        items.append(cur.strip())
    return [re.sub(r"\s+AS\s+\w+$", "", it, flags=re.I).strip().replace('"', "") for it in items]


def _project(items, env, star_order):
    """env: {'a': {col: val}, 'b': {...}}; unqualified names resolve against every alias in order."""
    row = []
    for it in items:
        if it == "*":
            for alias in star_order:
                row.extend(env[alias].values())
        elif it.endswith(".*"):
            row.extend(env[it[:-2]].values())
        elif "." in it:
            alias, col = it.split(".", 1)
            row.append(env[alias][col])
        else:
            for alias in star_order:
                if it in env[alias]:
                    row.append(env[alias][it])
                    break
            else:
                raise Unsupported(f"column {it}")
    return row


def _row_env(t, i):
    return {c: t[c][i] for c in t["__cols__"]}


def _nearest_args(text):
    args = {"k": 1, "max_distance": None, "stranded": False, "signed": False}
    for key, val in re.findall(r"(\w+)\s*:=\s*([\w.']+)", text):
        if key == "k":
            args["k"] = int(val)
        elif key == "max_distance":
            args["max_distance"] = int(val)
        elif key in ("stranded", "signed"):
            args[key] = val.lower() == "true"
    return args


def run_case(case, backend):
    """Returns the rows the query produces through ``backend`` (list of lists)."""
    q = " ".join(case["query"].split())
    tabs, cfgs = _tables(case)

    # --- literal range predicates ---------------------------------------------------------------------
    m = re.match(r'^SELECT (.+?) FROM (\w+) WHERE interval (INTERSECTS|CONTAINS|WITHIN) (ANY|ALL)?\s*\(?(.+?)\)?$', q)
    if m and "JOIN" not in q:
        sel, tname, op, quant, lits = m.groups()
        t = tabs[tname]
        ranges = [parse_range(x) for x in re.findall(r"'[^']+'", lits)]
        lit = {"chrom": [r[0] for r in ranges], "start": [r[1] for r in ranges], "end": [r[2] for r in ranges]}
        n = len(t["start"])
        if op == "INTERSECTS":
            if quant == "ALL":
                cnt = backend.counts(t, lit, cfgs.get(tname), None) if n else []
                keep = [c == len(ranges) for c in cnt]
            else:
                keep = backend.exists(t, lit, cfgs.get(tname), None) if n else []
        else:
            if not hasattr(backend, "range_filter"):
                raise Unsupported(op)
            keep = backend.range_filter(t, op.lower(), ranges, cfgs.get(tname))
        items = _select_items(sel)
        return [_project(items, {"t": _row_env(t, i)}, ["t"]) for i in range(n) if keep[i]]

    # --- INTERSECTS joins -----------------------------------------------------------------------------
    m = re.match(r"^SELECT (.+?) FROM (\w+) a (LEFT )?JOIN (\w+) b ON a\.interval INTERSECTS b\.interval$", q)
    if m:
        sel, ln, left, rn = m.groups()
        L, R = tabs[ln], tabs[rn]
        pairs = backend.pairs(L, R, cfgs.get(ln), cfgs.get(rn)) if len(L["start"]) and len(R["start"]) else []
        items = _select_items(sel)
        rows = [_project(items, {"a": _row_env(L, i), "b": _row_env(R, j)}, ["a", "b"]) for i, j in pairs]
        if left:
            matched = {i for i, _ in pairs}
            null_b = {c: None for c in R["__cols__"]}
            rows += [_project(items, {"a": _row_env(L, i), "b": null_b}, ["a", "b"]) for i in range(len(L["start"])) if i not in matched]
        return rows

    # --- DISTANCE over a chromosome equi-join --------------------------------------------------------------
    m = re.match(r"^SELECT (.+?) FROM (\w+) a JOIN (\w+) b ON a\.chrom = b\.chrom WHERE DISTANCE\(a\.interval, b\.interval\) < (\d+)$", q)
    if m:
        sel, an, bn, lim = m.groups()
        A, B = tabs[an], tabs[bn]
        idx = [(i, j) for i in range(len(A["start"])) for j in range(len(B["start"])) if A["chrom"][i] == B["chrom"][j]]
        if not idx:
            return []
        pa = {c: [A[c][i] for i, _ in idx] for c in A["__cols__"]}
        pb = {c: [B[c][j] for _, j in idx] for c in B["__cols__"]}
        d = backend.distance(pa, pb, cfgs.get(an), cfgs.get(bn))
        items = _select_items(sel)
        return [_project(items, {"a": _row_env(A, i), "b": _row_env(B, j)}, ["a", "b"])
                for (i, j), dist in zip(idx, d) if dist is not None and dist < int(lim)]

    # --- correlated NEAREST -----------------------------------------------------------------------------------
    if "NEAREST(" in q:
        inner = re.search(r"SELECT ((?:(?!SELECT).)+?) FROM (\w+) a ((?:CROSS JOIN LATERAL NEAREST\([^)]*\)(?: \w+)? ?)+)", q)
        if not inner:
            raise Unsupported(q)
        sel, qn, joins = inner.groups()
        Q = tabs[qn]
        calls = re.findall(r"CROSS JOIN LATERAL NEAREST\((\w+), ([^)]*)\)(?: (\w+))?", joins)
        env_rows = [({"a": _row_env(Q, i)}, i) for i in range(len(Q["start"]))]
        order = ["a"]
        for tn, argtext, alias in calls:
            alias = alias or f"__n{len(order)}"
            order.append(alias)
            T = tabs[tn]
            args = _nearest_args(argtext)
            res = backend.nearest(Q, T, cfgs.get(qn), cfgs.get(tn), **args) if len(Q["start"]) and len(T["start"]) else []
            by_q = {}
            for qi, ti, dist in res:
                by_q.setdefault(qi, []).append((ti, dist))
            nxt = []
            for env, qi in env_rows:
                for ti, dist in by_q.get(qi, []):
                    e2 = dict(env)
                    e2[alias] = dict(_row_env(T, ti), distance=dist)
                    nxt.append((e2, qi))
            env_rows = nxt
        items = _select_items(sel)
        rows = [_project(items, env, order) for env, _ in env_rows]
        outer = re.match(r"^SELECT \*, CLUSTER\(interval\) AS cid FROM \(.+\) sub$", q)
        if outer:
            sub = {"chrom": [r[0] for r in rows], "start": [r[1] for r in rows], "end": [r[2] for r in rows]}
            ids = backend.cluster(sub) if rows else []
            rows = [r + [cid] for r, cid in zip(rows, ids)]
        return rows

    # --- CLUSTER ---------------------------------------------------------------------------------------------
    m = re.match(r"^(?:WITH sub AS \(SELECT \* FROM (\w+)\) )?SELECT (DISTINCT )?(.+?) FROM (\w+)(?: \w)?( ORDER BY start LIMIT (\d+))?$", q)
    if m and "CLUSTER(interval) AS cid" in q:
        cte_src, distinct, sel, tname, _, limit = m.groups()
        t = tabs[cte_src or tname]
        n = len(t["start"])
        ids = backend.cluster(t) if n else []
        items = _select_items(sel)
        rows = []
        for i in range(n):
            env = {"t": _row_env(t, i), "p": _row_env(t, i)}
            row = []
            for it in items:
                if it.upper().startswith("CLUSTER("):
                    row.append(ids[i])
                else:
                    row.extend(_project([it], env, ["t"]))
            rows.append(row)
        if distinct:
            rows = [list(r) for r in dict.fromkeys(tuple(r) for r in rows)]
        if limit:
            start_pos = [it for it in items].index("start")
            rows = sorted(rows, key=lambda r: r[start_pos])[: int(limit)]
        return rows

    # --- MERGE ------------------------------------------------------------------------------------------------
    m = re.match(r'^(?:WITH sub AS \(SELECT \* FROM (\w+)\) )?SELECT MERGE\(interval\) FROM (\w+)( ORDER BY end DESC)?( LIMIT (\d+))?$', q.replace('"', ""))
    if m:
        cte_src, tname, desc, _, limit = m.groups()
        t = tabs[cte_src or tname]
        rows = [list(r[:3]) for r in backend.merge(t)] if len(t["start"]) else []
        rows = sorted(rows, key=lambda r: (r[0], r[1]))           # default ORDER BY chrom, start (merge.py:429-437)
        if desc:
            rows = sorted(rows, key=lambda r: -r[2])
        if limit:
            rows = rows[: int(limit)]
        return rows

    # --- DISJOIN ----------------------------------------------------------------------------------------------
    m = re.match(r"^SELECT (.+?) FROM DISJOIN\((\w+)\)$", q)
    if m:
        sel, tname = m.groups()
        t, cfg = tabs[tname], cfgs.get(tname)
        from giql_b200.table import canonical_offsets

        sadj, eadj = canonical_offsets(cfg)
        segs = backend.disjoin(t, None, cfg, None) if len(t["start"]) else []
        items = _select_items(sel)
        rows = []
        for i, s, e in segs:
            env = {"t": dict(_row_env(t, i), disjoin_chrom=t["chrom"][i], disjoin_start=s - sadj, disjoin_end=e - eadj)}
            rows.append(_project(items, env, ["t"]))
        return rows

    raise Unsupported(q)

"""giql_plugin.py — the reference-side plug-in: a ``b200`` target registered beside ``dialect="duckdb"``.

abdenlab/giql plugs engines in through its operator-expander registry
(``src/giql/expander.py:499-546``; targets ``src/giql/targets.py:57-73,189-232``).  Importing this module
while ``giql`` (hence ``sqlglot``) is importable registers :class:`B200Target` and one expander per GIQL operator,
so that ::

    import giql, giql_b200.giql_plugin                      # registration is the import's side effect
    text = giql.transpile(query, tables=[...], dialect="b200")
    rows = giql_b200.giql_plugin.run(conn, text)            # conn: a DuckDB (or sqlite3) connection holding the tables

``text`` is ordinary SQL for every statement (or part of one) the expanders decline — exactly what the reference's own
DuckDB override does for shapes outside its whitelist (``src/giql/expanders/intersects_duckdb.py:1674-1715``) — and
otherwise ``"B200 [<plan>, ...];\\n<residual SQL>"``: a JSON list of engine plans followed by SQL in which every
offloaded operator has been replaced by a join against the relation its plan produces.  :func:`run` executes the plans
on the GPU (``giql_b200.Engine``: dictionary → sort → join kernels behind the C ABI), registers their results on the
connection and lets the SQL engine finish projections, residual predicates, GROUP BY and ORDER BY on row ids.

Division of labour per operator (the expanders below; ``rowid`` is the engines' implicit row number of a base table):

* INTERSECTS column-to-column joins (INNER / SEMI / ANTI) — eligibility is the reference's OWN IEJoin whitelist
  (``IntersectsDuckDBIEJoinTransformer.transform_to_sql(root) is not None``, ``intersects_duckdb.py:618-804``); sides and
  join kind come from ``_IEJoinSides.from_intersects`` (``:360-410``, kinds ``"INNER" | "SEMI" | "ANTI"``).  The plan emits
  ``rel(l, r)`` (INNER) or ``rel(l)`` (rows with a match); the predicate becomes an equi-join on row ids.
* NEAREST under ``CROSS JOIN LATERAL`` with a column reference — the LATERAL is swapped for a derived relation exactly
  where the reference's decorrelated fallback swaps it (``src/giql/expanders/nearest.py:569-640``); plan emits
  ``rel(q, t, distance)``.
* CLUSTER(interval[, d][, stranded := ...]) in the SELECT list of a single registered base table — becomes a column of
  ``rel(row, cluster_id)``.  MERGE(interval[, d][, stranded]) as the only projection (optionally with ``COUNT(*)``) —
  the statement becomes a scan of ``rel(chrom, start, end, n)``.  DISJOIN(t[, reference := r]) as a FROM item — becomes
  ``t JOIN rel ON rowid`` with the ``disjoin_*`` columns.
* DISTANCE is a scalar CASE inside a projection: the SQL engine evaluates it beside the columns it already holds, so the
  registered expander IS the generic one (the GPU kernel ``giqlb_distance`` serves NEAREST's distance column).

Everything else falls through to the generic expanders, so ``dialect="b200"`` is never less correct than ``dialect=None``.
Without ``giql`` installed (this image: ``sqlglot`` is absent) the module still imports; :data:`HAVE_GIQL` is False,
``register_b200()`` raises, and :func:`run` / :func:`execute_plan` — which need only the engine — keep working.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Any, Iterable

PLAN_PREFIX = "B200 "
TARGET_NAME = "b200"

try:  # the reference package and its one hard dependency
    import giql as _giql  # noqa: F401
    from sqlglot import exp, parse_one

    HAVE_GIQL = True
    _IMPORT_ERROR: Exception | None = None
except ImportError as _e:  # pragma: no cover - exercised only where sqlglot exists
    HAVE_GIQL = False
    _IMPORT_ERROR = _e


# ---- plans ---------------------------------------------------------------------------------------------------------
def _table_plan(tables, name: str) -> dict:
    """The ``Table(...)`` column mapping and coordinate encoding of a registered table (``src/giql/table.py:16-93``)."""
    t = tables.get(name) if tables is not None else None
    cfg = {"name": name}
    for field, default in (("chrom_col", "chrom"), ("start_col", "start"), ("end_col", "end"), ("strand_col", "strand"),
                           ("coordinate_system", "0based"), ("interval_type", "half_open")):
        cfg[field] = getattr(t, field, default) if t is not None else default
    return cfg


def split(text: str) -> tuple[list[dict], str]:
    """``"B200 [plans];\\nSQL"`` → (plans, SQL); plain SQL → ([], SQL)."""
    if not text.startswith(PLAN_PREFIX):
        return [], text
    header, sep, sql = text.partition(";\n")
    if not sep:
        raise ValueError("malformed B200 statement: no residual SQL after the plan header")
    return json.loads(header[len(PLAN_PREFIX):]), sql


# ---- host side: executing plans ------------------------------------------------------------------------------------
def _fetch_columns(conn, cfg: dict, with_strand: bool) -> dict:
    """Row ids and genomic columns of a base table through any DB-API-ish connection (DuckDB, sqlite3)."""
    import numpy as np

    cols = [cfg["chrom_col"], cfg["start_col"], cfg["end_col"]] + ([cfg["strand_col"]] if with_strand else [])
    sel = ", ".join('"%s"' % c.replace('"', '""') for c in cols)
    rows = conn.execute('SELECT rowid, %s FROM "%s" ORDER BY rowid' % (sel, cfg["name"].replace('"', '""'))).fetchall()
    out = {"rowid": np.array([r[0] for r in rows], dtype=np.int64),
           "chrom": np.array([r[1] for r in rows], dtype=object),
           "start": np.array([r[2] for r in rows], dtype=np.int64),
           "end": np.array([r[3] for r in rows], dtype=np.int64)}
    if with_strand:
        out["strand"] = np.array([r[4] for r in rows], dtype=object)
    return out


def _mapping(cfg: dict):
    from .table import Table

    return Table(cfg["name"], coordinate_system=cfg["coordinate_system"], interval_type=cfg["interval_type"])


def _engine_table(eng, data: dict, cfg: dict):
    cols = {"chrom": data["chrom"], "start": data["start"], "end": data["end"]}
    if "strand" in data:
        cols["strand"] = data["strand"]
    return eng.table(cols, _mapping(cfg))


def _register(conn, name: str, columns: dict) -> None:
    """Makes ``columns`` (name → numpy array) queryable as relation ``name`` on the connection."""
    names = list(columns)
    if hasattr(conn, "register"):  # DuckDB: zero-copy Arrow registration
        import pyarrow as pa

        conn.register(name, pa.table({k: pa.array(v) for k, v in columns.items()}))
        return
    q = '"%s"' % name.replace('"', '""')  # sqlite3 and other DB-API connections: a temp table
    conn.execute("DROP TABLE IF EXISTS %s" % q)
    conn.execute("CREATE TEMP TABLE %s (%s)" % (q, ", ".join('"%s"' % n for n in names)))
    rows = zip(*[columns[n].tolist() for n in names])
    conn.executemany("INSERT INTO %s VALUES (%s)" % (q, ", ".join("?" for _ in names)), rows)


def execute_plan(eng, conn, plan: dict) -> None:
    """Runs one engine plan and registers its result relation ``plan["into"]`` on ``conn``."""
    import numpy as np

    from .table import canonical_offsets

    op = plan["op"]
    if op == "intersects":
        ld, rd = _fetch_columns(conn, plan["left"], False), _fetch_columns(conn, plan["right"], False)
        L, R = _engine_table(eng, ld, plan["left"]), _engine_table(eng, rd, plan["right"])
        if plan["kind"] == "INNER":
            l, r = eng.intersects(L, R).to_host()
            _register(conn, plan["into"], {"l": ld["rowid"][l], "r": rd["rowid"][r]})
        else:  # SEMI and ANTI both join against the rows that HAVE a match; the join kind stays in the SQL
            flags = eng.exists(L, R).cpu().numpy().astype(bool)
            _register(conn, plan["into"], {"l": ld["rowid"][flags]})
    elif op == "nearest":
        stranded = bool(plan.get("stranded"))
        qd, td = _fetch_columns(conn, plan["query"], stranded), _fetch_columns(conn, plan["target"], stranded)
        Q, T = _engine_table(eng, qd, plan["query"]), _engine_table(eng, td, plan["target"])
        res = eng.nearest(Q, T, k=int(plan.get("k", 1)), max_distance=plan.get("max_distance"), stranded=stranded,
                          signed=bool(plan.get("signed")))
        dist = res.distance.cpu().numpy().astype(object)
        dist[res.valid.cpu().numpy() == 0] = None                # SQL NULL distances (nearest.py:255-333)
        _register(conn, plan["into"], {"q": qd["rowid"][res.query_rows.cpu().numpy()], "t": td["rowid"][res.target_rows.cpu().numpy()],
                                       "distance": dist})
    elif op in ("cluster", "merge"):
        stranded = bool(plan.get("stranded"))
        d = _fetch_columns(conn, plan["table"], stranded)
        T = _engine_table(eng, d, plan["table"])
        if op == "cluster":
            ids = eng.cluster(T, distance=int(plan.get("distance", 0)), stranded=stranded).cpu().numpy()
            _register(conn, plan["into"], {"row": d["rowid"], "cluster_id": ids.astype(np.int64)})
        else:
            m = eng.merge(T, distance=int(plan.get("distance", 0)), stranded=stranded)
            sadj, eadj = canonical_offsets(_mapping(plan["table"]))   # regions go back to the table's declared encoding
            names = np.array(m.dictionary.names, dtype=object)
            cols = {"chrom": names[m.code.cpu().numpy()], "start": m.start.cpu().numpy().astype(np.int64) - sadj,
                    "end": m.end.cpu().numpy().astype(np.int64) - eadj, "n": m.count.cpu().numpy().astype(np.int64)}
            if stranded and m.strand is not None:
                cols["strand"] = np.array([eng.strand_literal(v) for v in m.strand.cpu().tolist()], dtype=object)
            _register(conn, plan["into"], cols)
    elif op == "disjoin":
        td = _fetch_columns(conn, plan["target"], False)
        rd = td if plan.get("reference") is None else _fetch_columns(conn, plan["reference"], False)
        T = _engine_table(eng, td, plan["target"])
        Rf = T if plan.get("reference") is None else _engine_table(eng, rd, plan["reference"])
        seg = eng.disjoin(T, None if plan.get("reference") is None else Rf)
        rows = seg.target_rows.cpu().numpy()
        sadj, eadj = canonical_offsets(_mapping(plan["target"]))     # output in the target's encoding (disjoin.py:139-141)
        _register(conn, plan["into"], {"row": td["rowid"][rows], "disjoin_chrom": td["chrom"][rows],
                                       "disjoin_start": seg.start.cpu().numpy().astype(np.int64) - sadj,
                                       "disjoin_end": seg.end.cpu().numpy().astype(np.int64) - eadj})
    else:
        raise ValueError(f"unknown B200 plan operator {op!r}")


def run(conn, text: str, engine=None):
    """Executes a statement produced by ``giql.transpile(..., dialect="b200")`` (or plain SQL) on ``conn``."""
    plans, sql = split(text)
    if plans:
        if engine is None:
            from .engine import Engine

            engine = Engine(0)  # raises EngineUnavailable without the CUDA library: there is no CPU fallback
        for plan in plans:
            execute_plan(engine, conn, plan)
    return conn.execute(sql)


# ---- reference side: the target and its expanders ------------------------------------------------------------------
def register_b200():
    """Registers ``B200Target`` and its expanders on ``giql.expander.REGISTRY``; returns the target class."""
    if not HAVE_GIQL:
        raise ImportError(f"giql / sqlglot are not importable ({_IMPORT_ERROR}); the b200 target cannot be registered")
    from giql import Capabilities, DuckDBTarget, GenericTarget, Target, register
    from giql.expander import REGISTRY
    from giql.expanders._params import coerce_bool_param
    from giql.expanders.intersects import _expand_spatial_op
    from giql.expanders.intersects_duckdb import (IntersectsDuckDBIEJoinTransformer, _has_sibling_spatial_predicate,
                                                  _has_star_projection, _IEJoinSides, _is_column_intersects,
                                                  _match_count_overlaps)
    from giql.expanders.nearest import _nearest_params, _outer_relation
    from giql.expressions import GIQLCluster, GIQLDisjoin, GIQLDistance, GIQLMerge, GIQLNearest, Intersects
    from giql.resolver import ResolvedInterval, ResolvedRef

    @dataclass(frozen=True)
    class B200Target(Target):
        """GPU execution target; residual SQL is serialised for DuckDB (``src/giql/targets.py:97-133``)."""

        name: str = TARGET_NAME
        sqlglot_dialect: str | None = "duckdb"
        capabilities: Capabilities = DuckDBTarget().capabilities

    def generic(node, ctx):
        return REGISTRY.resolve(GenericTarget(), type(node))(node, ctx)

    def col(table: str, name: str):
        return exp.column(name, table=table, quoted=False)

    def add_plan(ctx, plan: dict) -> None:
        """Statement finalizer: prepend the plan header (``expander.py``: finalizers run after every node-local
        replacement, in registration order, and may return a new root — here a verbatim ``Command``)."""

        def finalize(root):
            if isinstance(root, exp.Command) and str(root.this).startswith(PLAN_PREFIX):
                plans, sql = split(str(root.this))
                return exp.Command(this=PLAN_PREFIX + json.dumps(plans + [plan]) + ";\n" + sql)
            return exp.Command(this=PLAN_PREFIX + json.dumps([plan]) + ";\n" + root.sql(dialect="duckdb"))

        ctx.add_statement_finalizer(finalize)

    def base_table_name(t) -> str | None:
        return t.name if isinstance(t, exp.Table) and t.name else None

    def literal_int(e, default=0):
        return default if e is None else int(str(e.this if isinstance(e, exp.Literal) else e))

    # -- INTERSECTS ------------------------------------------------------------------------------------------------
    @register(B200Target, Intersects)
    def expand_intersects_b200(node, ctx):
        if not (isinstance(node, Intersects) and _is_column_intersects(node)):
            return _expand_spatial_op(node, ctx, "intersects")
        root = node.root()
        if not isinstance(root, exp.Select) or _has_sibling_spatial_predicate(node, root) or _match_count_overlaps(root) is not None:
            return _expand_spatial_op(node, ctx, "intersects")
        shape = IntersectsDuckDBIEJoinTransformer(ctx.tables)
        if shape.transform_to_sql(root) is None:             # the reference's own whitelist (ValueError propagates, as there)
            return _expand_spatial_op(node, ctx, "intersects")
        intersects, the_join = shape._find_target_join(root)
        from_table = root.args["from_"].this
        sides = _IEJoinSides.from_intersects(from_table, the_join, intersects) if intersects is node else None
        if sides is None or (sides.is_left_only_join and the_join.args.get("on") is not node):
            return _expand_spatial_op(node, ctx, "intersects")   # SEMI / ANTI with residual ON-conjuncts on the right table
        rel = ctx.alias()
        add_plan(ctx, {"op": "intersects", "kind": sides.kind, "into": rel,
                       "left": _table_plan(ctx.tables, sides.left_table.name),
                       "right": _table_plan(ctx.tables, sides.right_table.name)})
        if sides.kind == "INNER":   # a JOIN pairs ON a.rowid = pairs.l goes in front; the predicate becomes b.rowid = pairs.r
            joins = root.args["joins"]
            at = next(i for i, j in enumerate(joins) if j is the_join)
            joins.insert(at, exp.Join(this=exp.to_table(rel), on=exp.EQ(this=col(sides.left_user_alias, "rowid"), expression=col(rel, "l"))))
            joins[at].parent = root
            return exp.EQ(this=col(sides.right_user_alias, "rowid"), expression=col(rel, "r"))
        the_join.set("this", exp.to_table(rel))               # SEMI / ANTI JOIN rows-with-a-match ON a.rowid = rel.l
        return exp.EQ(this=col(sides.left_user_alias, "rowid"), expression=col(rel, "l"))

    # -- NEAREST ---------------------------------------------------------------------------------------------------
    @register(B200Target, GIQLNearest)
    def expand_nearest_b200(node, ctx):
        res = ctx.resolution
        target_ref = res.slot("this") if res is not None else None
        ref = res.slot("reference") if res is not None else None
        lateral = node.parent
        join = lateral.parent if isinstance(lateral, exp.Lateral) else None
        root = node.root()
        if not (isinstance(target_ref, ResolvedRef) and isinstance(ref, ResolvedInterval) and ref.kind != "literal_range"
                and isinstance(join, exp.Join) and isinstance(root, exp.Select) and ref.table is not None
                and ctx.tables.get(target_ref.name) is not None and not _has_star_projection(root)):
            return generic(node, ctx)
        relation, outer_alias = _outer_relation(ref)          # physical outer table and its correlation name
        k, max_distance, stranded, signed = _nearest_params(node)
        alias_node = lateral.args.get("alias")
        alias = alias_node.name if alias_node is not None and alias_node.name else ctx.alias()
        rel = ctx.alias()
        add_plan(ctx, {"op": "nearest", "into": rel, "query": _table_plan(ctx.tables, relation),
                       "target": _table_plan(ctx.tables, target_ref.name), "k": k, "max_distance": max_distance,
                       "stranded": bool(stranded), "signed": bool(signed)})
        sub = parse_one(f'SELECT t.*, p.distance AS distance, p.q AS "__b200_q" FROM "{rel}" AS p '
                        f'JOIN "{target_ref.name}" AS t ON t.rowid = p.t', dialect="duckdb").subquery(alias)
        lateral.replace(sub)
        join.set("kind", None)
        join.set("side", None)
        join.set("on", exp.EQ(this=col(alias, "__b200_q"), expression=col(outer_alias or relation, "rowid")))
        return node                                            # detached with its LATERAL, as in the reference's fallback form

    # -- CLUSTER / MERGE -------------------------------------------------------------------------------------------
    def single_table_select(node):
        sel = node.find_ancestor(exp.Select)
        if sel is None or sel.args.get("joins") or sel.args.get("with_"):
            return None, None
        frm = sel.args.get("from_")
        name = base_table_name(frm.this) if frm is not None else None
        return (sel, frm.this) if name else (None, None)

    @register(B200Target, GIQLCluster)
    def expand_cluster_b200(node, ctx):
        sel, tbl = single_table_select(node)
        if sel is None or node.args.get("predicate") is not None or ctx.tables.get(tbl.name) is None or sel.args.get("where"):
            return generic(node, ctx)                          # WHERE changes the clustered row set: left to the SQL path
        rel = ctx.alias()
        add_plan(ctx, {"op": "cluster", "into": rel, "table": _table_plan(ctx.tables, tbl.name),
                       "distance": literal_int(node.args.get("distance")),
                       "stranded": bool(coerce_bool_param(node.args.get("stranded")))})
        sel.append("joins", exp.Join(this=exp.to_table(rel), on=exp.EQ(this=col(tbl.alias_or_name, "rowid"), expression=col(rel, "row"))))
        return col(rel, "cluster_id")

    @register(B200Target, GIQLMerge)
    def expand_merge_b200(node, ctx):
        sel, tbl = single_table_select(node)
        plain = sel is not None and all(
            (s.this if isinstance(s, exp.Alias) else s) is node
            or (isinstance(s.this if isinstance(s, exp.Alias) else s, exp.Count) and isinstance((s.this if isinstance(s, exp.Alias) else s).this, exp.Star))
            for s in sel.expressions)
        if not plain or node.args.get("predicate") is not None or ctx.tables.get(tbl.name) is None or any(
                sel.args.get(k) for k in ("where", "group", "having", "order", "limit", "distinct")) or sel is not node.root():
            return generic(node, ctx)
        rel = ctx.alias()
        stranded = bool(coerce_bool_param(node.args.get("stranded")))
        cfg = _table_plan(ctx.tables, tbl.name)
        add_plan(ctx, {"op": "merge", "into": rel, "table": cfg, "distance": literal_int(node.args.get("distance")), "stranded": stranded})
        names = [("chrom", cfg["chrom_col"]), ("start", cfg["start_col"]), ("end", cfg["end_col"])] + ([("strand", cfg["strand_col"])] if stranded else [])
        proj = [exp.alias_(col(rel, src), dst, quoted=True) for src, dst in names]
        for s in sel.expressions:                              # COUNT(*) [AS x] -> the region's row count
            inner = s.this if isinstance(s, exp.Alias) else s
            if inner is not node:
                proj.append(exp.alias_(col(rel, "n"), s.alias if isinstance(s, exp.Alias) else "count_star()", quoted=True))
        sel.set("expressions", proj)
        sel.set("from_", exp.From(this=exp.to_table(rel)))
        order = [exp.Ordered(this=col(rel, "chrom")), exp.Ordered(this=col(rel, "start"))]
        sel.set("order", exp.Order(expressions=order))         # the reference's MERGE orders by chrom, start (merge.py:429-437)
        return node

    # -- DISJOIN ---------------------------------------------------------------------------------------------------
    @register(B200Target, GIQLDisjoin)
    def expand_disjoin_b200(node, ctx):
        res = ctx.resolution
        target_ref = res.slot("this") if res is not None else None
        reference = res.slot("reference") if res is not None and node.args.get("reference") is not None else None
        holder = node.parent                                   # DISJOIN(...) parses as a table function: Table(this=node)
        if not (isinstance(target_ref, ResolvedRef) and isinstance(holder, exp.Table) and ctx.tables.get(target_ref.name) is not None
                and (node.args.get("reference") is None or (isinstance(reference, ResolvedRef) and ctx.tables.get(reference.name) is not None))):
            return generic(node, ctx)
        rel = ctx.alias()
        add_plan(ctx, {"op": "disjoin", "into": rel, "target": _table_plan(ctx.tables, target_ref.name),
                       "reference": None if node.args.get("reference") is None else _table_plan(ctx.tables, reference.name)})
        alias = holder.alias or target_ref.name
        sub = parse_one(f'SELECT t.*, p.disjoin_chrom, p.disjoin_start, p.disjoin_end FROM "{target_ref.name}" AS t '
                        f'JOIN "{rel}" AS p ON t.rowid = p.row', dialect="duckdb").subquery(alias)
        holder.replace(sub)
        return node

    # -- DISTANCE: a scalar CASE the SQL engine evaluates in place ------------------------------------------------------
    register(B200Target, GIQLDistance)(lambda node, ctx: generic(node, ctx))

    return B200Target


B200Target = register_b200() if HAVE_GIQL else None


def transpile(query: str, tables: Iterable[Any] | None = None) -> str:
    """``giql.transpile(query, tables=tables, dialect="b200")`` (raises ImportError without giql)."""
    if not HAVE_GIQL:
        raise ImportError(f"giql / sqlglot are not importable ({_IMPORT_ERROR})")
    import giql

    return giql.transpile(query, tables=list(tables or []), dialect=TARGET_NAME)

"""A minimal GIQL front-end for the B200 engine: query string in, rows out, without sqlglot.

The reference's parser/AST (``giql.dialect``, ``giql.expressions``) is reused as-is whenever ``giql`` is importable
(INTEGRATION.md shows the ``@giql.register(B200Target, ...)`` binding).  This module is the fallback SURVEY.md §8(f)-1
asks for: a small recursive-descent parser for the operator surface users actually write and a
binder that maps every recognised shape onto ONE engine operator call, evaluating projections, DISTINCT, ORDER BY and
LIMIT on the host.  Anything else raises :class:`UnsupportedQuery` — the analogue of the reference's expanders
declining a shape to the generic SQL path (``src/giql/expanders/intersects_duckdb.py:618-804``).

Recognised shapes (GIQL semantics cited per branch in :class:`Session`):

* ``FROM a [x] [INNER|LEFT|SEMI|ANTI] JOIN b [y] ON x.interval INTERSECTS|CONTAINS|WITHIN y.interval`` and the comma form
  ``FROM a x, b y WHERE x.interval INTERSECTS y.interval``
* the same join with the overlap predicate spelled out over any columns, ``ON x.chrom = y.c AND x.start < y.e AND x.end > y.s``
  (what the generic expander writes for INTERSECTS), also against ``DISJOIN(...)`` as a join source
* ``FROM t WHERE interval INTERSECTS|CONTAINS|WITHIN 'chr1:100-200'`` and ``... INTERSECTS|CONTAINS|WITHIN ANY|ALL(...)``
* ``FROM a x CROSS JOIN LATERAL NEAREST(b, reference := x.interval[, k, max_distance, stranded, signed]) [y]`` (repeatable);
  the standalone placement ``FROM NEAREST(b, reference := 'chr1:350-360'[, k, max_distance, signed]) [y]``
* ``FROM a x JOIN b y ON x.chrom = y.chrom WHERE DISTANCE(x.interval, y.interval[, ...]) < n``; the plain cross join
  ``FROM a x, b y`` / ``a x CROSS JOIN b y`` feeding ``DISTANCE`` in the select list (bounded product)
* ``SELECT ..., CLUSTER(interval[, d][, stranded := true]) AS c FROM t``; ``SELECT MERGE(interval[, d][, stranded := true])[, COUNT(*) | SUM | MIN | MAX | AVG(col) ...] FROM t [HAVING COUNT(*) ...]``
* ``<select> UNION [ALL] | EXCEPT | INTERSECT <select>`` (also as a CTE or derived table feeding an operator)
* ``FROM DISJOIN(t[, reference := r | (<query>)])``;  ``WITH name AS (SELECT * FROM t) ...``;  a derived table ``FROM (<query>) alias``
* ``SELECT <left cols>, COUNT(*) | COUNT(y.col) FROM a x [LEFT] JOIN b y ON x.interval INTERSECTS y.interval GROUP BY <left cols>``
  (count_overlaps: one per-row count probe, no pair list); any other ``GROUP BY cols`` with ``COUNT / SUM / MIN / MAX / AVG /
  COUNT(DISTINCT)`` over the joined rows and ``HAVING <aggregate> <op> <number> [AND ...]``
* residual ON-conjuncts ``AND x.col <op> y.col`` (same / opposite strand joins), ``WHERE col <op> col | literal``
* ``CLUSTER / MERGE(interval, predicate := col <op> PREV(col))`` and conjunctions of such equalities
* projections ``*``, ``x.*``, ``[x.]col [AS name]``, ``DISTANCE(x.interval, y.interval[, ...]) [AS name]``, literals and arithmetic
  (``end - start AS length``, ``LEAST`` / ``GREATEST`` / ``ABS``) — also inside aggregates and WHERE comparisons; ``DISTINCT``;
  ``ORDER BY col [DESC]``; ``LIMIT n [OFFSET m]``; ``col BETWEEN a AND b``, ``col [NOT] IN (...)``, ``NOT <literal-range predicate>``

A *backend* supplies the operators on plain dict tables (``{"chrom": [...], "start": [...], "end": [...], ...}``):
:class:`EngineBackend` is the product (CUDA through the C ABI); the test-suite plugs its CPU oracle in the same seat.
"""
from __future__ import annotations

import re
from typing import Any, Optional

import numpy as np
import torch

from .ranges import parse_range
from .table import Table, canonical_offsets


_AGGREGATES = ("COUNT", "SUM", "MIN", "MAX", "AVG", "STRING_AGG")
_SET_OPS = ("UNION", "EXCEPT", "INTERSECT")
_CROSS_JOIN_LIMIT = 1 << 24     # rows of `FROM a, b` without any join predicate the front-end will materialise


class UnsupportedQuery(ValueError):
    """The query shape is outside the built-in front-end; run it through giql.transpile + a SQL engine."""


# ---- backend over the engine -----------------------------------------------------------------------------------------
class EngineBackend:
    """Operator calls on dict tables -> giql_b200.Engine -> ctypes -> libgiqlb200.so -> sm_100a kernels.

    Two flavours per operator: the list-returning one (``pairs``, ``exists`` ...) is the small-case interface the parity
    tests share with the CPU oracle; the ``*_arrays`` one returns numpy arrays without per-row Python work and is what
    :class:`Session` uses.  Uploaded tables are cached on the dict they came from, so a registered relation is sent to
    the device once, not once per statement."""

    name = "b200"

    def __init__(self, engine, join_mode: str = "auto"):
        self.e = engine
        self.join_mode = join_mode

    def _table(self, T, cfg):
        cfg = cfg or Table("t")
        cache = T.setdefault("__device__", {}) if isinstance(T, dict) else {}
        key = (id(self.e), cfg.chrom_col, cfg.start_col, cfg.end_col, cfg.strand_col, cfg.coordinate_system, cfg.interval_type)
        if key not in cache:
            data = {cfg.chrom_col: T["chrom"], cfg.start_col: T["start"], cfg.end_col: T["end"]}
            if "strand" in T and cfg.strand_col:
                data[cfg.strand_col] = T["strand"]
            cache[key] = self.e.table(data, cfg)
        return cache[key]

    # -- array flavour (numpy in, numpy out)
    def pair_arrays(self, L, R, op="intersects", lcfg=None, rcfg=None, residuals=()):
        """``residuals``: extra ON-conjuncts as dicts ``{l, r, op[, lvalid, rvalid]}`` of per-row host columns, applied
        on the device (giqlb_pairs_filter) before the pairs travel back."""
        l, r = self._table(L, lcfg), self._table(R, rcfg)
        res = self.e.intersects(l, r, mode=self.join_mode) if op == "intersects" else (self.e.contains if op == "contains" else self.e.within)(l, r)
        for x in residuals:
            res = self.e.filter_pairs(res, x["l"], x["r"], x["op"], left_valid=x.get("lvalid"), right_valid=x.get("rvalid"))
        return res.to_host()

    def exists_array(self, L, R, lcfg=None, rcfg=None):
        return self.e.exists(self._table(L, lcfg), self._table(R, rcfg)).cpu().numpy().astype(bool)

    def counts_array(self, L, R, lcfg=None, rcfg=None):
        return self.e.count_overlaps(self._table(L, lcfg), self._table(R, rcfg)).cpu().numpy()

    def distance_arrays(self, A, B, a_rows, b_rows, acfg=None, bcfg=None, stranded=False, signed=False):
        """(distance int64, valid bool) for the paired rows ``A[a_rows[i]]``, ``B[b_rows[i]]``"""
        dev = self.e.device
        ia, ib = (torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)).to(dev) for x in (a_rows, b_rows))
        d, v = self.e.distance(self._table(A, acfg), self._table(B, bcfg), ia, ib, stranded=stranded, signed=signed)
        return d.cpu().numpy().astype(np.int64), v.cpu().numpy().astype(bool)

    def nearest_arrays(self, Q, T, qcfg=None, tcfg=None, k=1, max_distance=None, stranded=False, signed=False):
        res = self.e.nearest(self._table(Q, qcfg), self._table(T, tcfg), k=k, max_distance=max_distance, stranded=stranded, signed=signed)
        return tuple(x.cpu().numpy() for x in (res.query_rows, res.target_rows, res.distance, res.valid))

    def cluster_array(self, T, distance=0, stranded=False, predicate=None):
        return self.e.cluster(self._table(T, None), distance=distance, stranded=stranded, predicate=predicate).cpu().numpy()

    # -- list flavour (shared with the oracle backend of the test-suite)
    def pairs(self, L, R, lcfg=None, rcfg=None):
        l, r = self.pair_arrays(L, R, "intersects", lcfg, rcfg)
        return sorted(zip(l.tolist(), r.tolist()))

    def spatial_pairs(self, L, R, op, lcfg=None, rcfg=None):
        l, r = self.pair_arrays(L, R, op, lcfg, rcfg)
        return sorted(zip(l.tolist(), r.tolist()))

    def exists(self, L, R, lcfg=None, rcfg=None):
        return self.exists_array(L, R, lcfg, rcfg).tolist()

    def counts(self, L, R, lcfg=None, rcfg=None):
        return self.counts_array(L, R, lcfg, rcfg).tolist()

    def distance(self, A, B, acfg=None, bcfg=None, stranded=False, signed=False):
        d, v = self.e.distance(self._table(A, acfg), self._table(B, bcfg), stranded=stranded, signed=signed)
        return [int(x) if ok else None for x, ok in zip(d.cpu().tolist(), v.cpu().tolist())]

    def nearest(self, Q, T, qcfg=None, tcfg=None, k=1, max_distance=None, stranded=False, signed=False):
        res = self.e.nearest(self._table(Q, qcfg), self._table(T, tcfg), k=k, max_distance=max_distance,
                             stranded=stranded, signed=signed)
        return res.rows()

    def cluster(self, T, distance=0, stranded=False, predicate=None):
        return self.cluster_array(T, distance, stranded, predicate).tolist()

    def merge(self, T, distance=0, stranded=False):
        return self.e.merge(self._table(T, None), distance=distance, stranded=stranded).rows()

    def disjoin(self, T, REF=None, tcfg=None, rcfg=None):
        res = self.e.disjoin(self._table(T, tcfg), None if REF is None else self._table(REF, rcfg))
        return res.rows()


# ---- tokenizer / parser -------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""\s*(?:(?P<str>'(?:[^']|'')*')|(?P<qid>"[^"]*")|(?P<num>-?\d+(?:\.\d+)?)|(?P<op>:=|=>|<=|>=|<>|!=|[(),.*=<>;+\-/])|(?P<id>[A-Za-z_][A-Za-z_0-9]*))""")
_COMMENT = re.compile(r"('(?:[^']|'')*'|\"[^\"]*\")|--[^\n]*")
_KEYWORDS = {"SELECT", "DISTINCT", "FROM", "WHERE", "JOIN", "INNER", "LEFT", "SEMI", "ANTI", "CROSS", "LATERAL", "ON", "AS", "AND",
             "ORDER", "BY", "LIMIT", "DESC", "ASC", "WITH", "INTERSECTS", "CONTAINS", "WITHIN", "ANY", "ALL", "SOME", "IS", "NULL", "NOT",
             "OUTER", "GROUP"}


def _tokenize(text: str):
    out, pos = [], 0
    text = _COMMENT.sub(lambda m: m.group(1) or " ", text).strip().rstrip(";")   # `-- ...` comments (string literals are kept)
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise UnsupportedQuery(f"cannot tokenize near {text[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("str") is not None:
            out.append(("str", m.group("str")[1:-1].replace("''", "'")))
        elif m.group("qid") is not None:
            out.append(("id", m.group("qid")[1:-1]))
        elif m.group("num") is not None:
            text_num = m.group("num")
            if text_num[0] == "-" and out and (out[-1][0] in ("num", "str") or out[-1] == ("op", ")") or
                                               (out[-1][0] == "id" and out[-1][1].upper() not in ("BETWEEN", "IN", "OFFSET", "HAVING"))):
                out.append(("op", "-"))                                   # `a -5`, `(x) -1`: binary minus, not a signed literal
                text_num = text_num[1:]
            out.append(("num", float(text_num) if "." in text_num else int(text_num)))
        elif m.group("op") is not None:
            out.append(("op", m.group("op")))
        else:
            word = m.group("id")
            out.append(("kw", word.upper()) if word.upper() in _KEYWORDS else ("id", word))
    return out


def _int_arg(v, what: str, default=None):
    if v is None:
        return default
    if isinstance(v, bool) or not isinstance(v, int):
        raise ValueError(f"{what} must be an integer literal")
    return v


def _bool_arg(v, what: str) -> bool:
    if v is None:
        return False
    if not isinstance(v, bool):
        raise ValueError(f"{what} must be true or false")
    return v


class _Parser:
    def __init__(self, tokens):
        self.t, self.i = tokens, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def at(self, kind, val=None, k=0):
        tk = self.peek(k)
        return tk[0] == kind and (val is None or tk[1] == val)

    def take(self, kind=None, val=None):
        tk = self.peek()
        if (kind and tk[0] != kind) or (val is not None and tk[1] != val):
            raise UnsupportedQuery(f"expected {val or kind}, found {tk[1]!r}")
        self.i += 1
        return tk[1]

    def accept(self, kind, val=None):
        if self.at(kind, val):
            self.i += 1
            return True
        return False

    # query := [WITH name AS (query) {, ...}] select
    def query(self):
        ctes = {}
        if self.accept("kw", "WITH"):
            while True:
                name = self.take("id")
                self.take("kw", "AS")
                self.take("op", "(")
                ctes[name] = self.query()
                self.take("op", ")")
                if not self.accept("op", ","):
                    break
        q = self.select()
        q["ctes"] = ctes
        # <select> {UNION [ALL] | EXCEPT | INTERSECT <select>}: a trailing ORDER BY / LIMIT / OFFSET belongs to the whole set operation
        while self.at("id") and self.peek()[1].upper() in _SET_OPS:
            op = self.take("id").upper()
            keep_all = self.accept("kw", "ALL")
            if keep_all and op != "UNION":
                raise UnsupportedQuery(f"{op} ALL is finished by the SQL engine")
            right = self.select()
            right["ctes"] = ctes
            node = {"setop": op, "all": keep_all, "left": q, "right": right, "ctes": ctes,
                    "order": right["order"], "limit": right["limit"], "offset": right.get("offset", 0)}
            right["order"], right["limit"], right["offset"] = [], None, 0
            q = node
        return q

    def select(self):
        self.take("kw", "SELECT")
        q: dict[str, Any] = {"distinct": self.accept("kw", "DISTINCT"), "proj": [], "joins": [], "where": [], "order": [], "limit": None}
        while True:
            q["proj"].append(self.projection())
            if not self.accept("op", ","):
                break
        self.take("kw", "FROM")
        q["source"] = self.source()
        while True:
            if self.accept("op", ","):
                q["joins"].append({"kind": "comma", "source": self.source()})
            elif self.at("kw", "CROSS"):
                self.take("kw", "CROSS"); self.take("kw", "JOIN")
                if self.accept("kw", "LATERAL"):
                    q["joins"].append({"kind": "nearest", **self.nearest_call()})
                else:                                                    # `a CROSS JOIN b` == `a, b`
                    q["joins"].append({"kind": "comma", "source": self.source()})
            elif self.at("kw", "JOIN") or self.at("kw", "INNER") or self.at("kw", "LEFT") or self.at("kw", "SEMI") or self.at("kw", "ANTI"):
                kind = "inner"
                if not self.at("kw", "JOIN"):
                    kind = self.take("kw").lower()
                    self.accept("kw", "OUTER")
                self.take("kw", "JOIN")
                src = self.source()
                self.take("kw", "ON")
                on = [self.condition()]
                while self.accept("kw", "AND"):
                    on.append(self.condition())
                q["joins"].append({"kind": kind, "source": src, "on": on})
            else:
                break
        if self.accept("kw", "WHERE"):
            while True:
                q["where"].append(self.condition())
                if not self.accept("kw", "AND"):
                    break
        q["group"] = []
        if self.accept("kw", "GROUP"):
            self.take("kw", "BY")
            while True:
                q["group"].append(self.column_ref())
                if not self.accept("op", ","):
                    break
        q["having"] = []
        if self.at("id") and self.peek()[1].upper() == "HAVING":
            # HAVING <aggregate> <op> <number> [AND ...]: a filter over the groups
            self.take("id")
            while True:
                if not (self.at("id") and self.peek()[1].upper() in _AGGREGATES and self.at("op", "(", 1)):
                    raise UnsupportedQuery("HAVING compares aggregates with numbers; anything else is finished by the SQL engine")
                item = self.aggregate_args(self.take("id").upper())
                op = self.take("op")
                if not self.at("num"):
                    raise UnsupportedQuery("HAVING compares aggregates with numbers; anything else is finished by the SQL engine")
                q["having"].append({"item": item, "op": op, "value": self.take("num")})
                if not self.accept("kw", "AND"):
                    break
        if self.accept("kw", "ORDER"):
            self.take("kw", "BY")
            while True:
                col = self.column_ref()
                desc = self.accept("kw", "DESC")
                if not desc:
                    self.accept("kw", "ASC")
                q["order"].append((col, desc))
                if not self.accept("op", ","):
                    break
        if self.accept("kw", "LIMIT"):
            q["limit"] = _int_arg(self.take("num"), "LIMIT")
        q["offset"] = 0
        if self.at("id") and self.peek()[1].upper() == "OFFSET":
            self.take("id")
            q["offset"] = _int_arg(self.take("num"), "OFFSET")
        return q

    def source(self):
        if self.accept("op", "("):
            sub = self.query()
            self.take("op", ")")
            had_as = self.accept("kw", "AS")
            named = had_as or (self.at("id") and self.peek()[1].upper() not in ("OFFSET", "HAVING"))
            alias = self.take("id") if named else f"__subquery{self.i}"                    # `FROM (<query>)` without a name
            return {"kind": "subquery", "query": sub, "alias": alias}
        name = self.take("id")
        if name.upper() == "NEAREST" and self.at("op", "("):
            # standalone placement: `FROM NEAREST(target, reference := 'chr1:350-360', k := 2)` — the literal range is the one
            # query interval (src/giql/expanders/nearest.py:343-350, 782-817; the literal is canonical 0-based half-open)
            self.take("op", "(")
            target = self.take("id")
            args = self.named_args()
            self.take("op", ")")
            ref = args.get("reference")
            if not isinstance(ref, str):
                raise UnsupportedQuery("a NEAREST in FROM takes a literal range as its reference")
            if _bool_arg(args.get("stranded"), "stranded"):
                raise UnsupportedQuery("stranded NEAREST with a literal reference is finished by the SQL engine")
            k = _int_arg(args.get("k"), "NEAREST k", 1)
            if k < 1:
                raise ValueError("NEAREST k must be >= 1")
            return {"kind": "nearest_literal", "table": target, "range": ref, "k": k, "stranded": False,
                    "max_distance": _int_arg(args.get("max_distance"), "NEAREST max_distance"),
                    "signed": _bool_arg(args.get("signed"), "signed"), "alias": self.alias() or target}
        if name.upper() == "DISJOIN" and self.at("op", "("):
            self.take("op", "(")
            target = self.take("id")
            ref = None
            if self.accept("op", ","):
                # DISJOIN(t, reference := r) / reference => r (src/giql/expressions.py:386-413: one positional argument only)
                if not (self.at("id") and (self.at("op", ":=", 1) or self.at("op", "=>", 1))):
                    raise ValueError("DISJOIN accepts at most one positional argument (the target table); got 2. "
                                     "Pass the reference set as 'reference := ...'.")
                key = self.take("id").lower()
                if key != "reference":
                    raise ValueError(f"DISJOIN got unexpected named argument(s): {key}. Valid arguments: reference.")
                self.take("op")
                if self.accept("op", "("):
                    # reference := (<query>): the subquery's chrom / start / end columns are the reference set, read as canonical
                    # 0-based half-open coordinates — a derived relation has no Table encoding (disjoin.py:255-270)
                    ref = {"subquery": self.query()}
                    self.take("op", ")")
                else:
                    ref = self.take("id")
            self.take("op", ")")
            return {"kind": "disjoin", "table": target, "reference": ref, "alias": self.alias() or target}
        return {"kind": "table", "table": name, "alias": self.alias() or name}

    def alias(self):
        if self.accept("kw", "AS"):
            return self.take("id")
        if self.at("id") and self.peek()[1].upper() not in _SET_OPS + ("OFFSET", "HAVING"):
            return self.take("id")
        return None

    def named_args(self):
        """`name := value` / `name => value` pairs up to the closing parenthesis (src/giql/expressions.py:184-194)."""
        args = {}
        while self.accept("op", ","):
            if self.at("id") and (self.at("op", ":=", 1) or self.at("op", "=>", 1)):
                key = self.take("id").lower()
                self.take("op")
                args[key] = self.predecessor_predicate() if key == "predicate" else self.scalar()
            else:
                args.setdefault("_positional", []).append(self.scalar())
        return args

    def predecessor_predicate(self):
        """`predicate := <col> <op> PREV(<col>)` (either side may carry the PREV): the CLUSTER / MERGE gate against the
        immediate sorted predecessor (src/giql/expanders/cluster.py:281-299, 587-644; PREV takes exactly one column and
        does not nest, :618-630).  Normalised to {"cur", "op", "prev"}."""
        def operand():
            if self.at("id") and self.peek()[1].upper() == "PREV" and self.at("op", "(", 1):
                self.take("id")
                self.take("op", "(")
                if self.at("id") and self.peek()[1].upper() == "PREV":
                    raise ValueError("PREV() cannot be nested; a CLUSTER/MERGE predicate compares only the immediate predecessor.")
                col = self.column_ref()
                if self.at("op", ","):
                    raise ValueError("PREV() takes exactly one column argument; got 2.")
                self.take("op", ")")
                return True, col[1]
            return False, self.column_ref()[1]
        def comparison():
            lp, lcol = operand()
            op = self.take("op")
            rp, rcol = operand()
            op = {"<>": "!=", "==": "="}.get(op, op)
            if op not in ("=", "!=", "<", "<=", ">", ">=") or lp == rp:
                raise UnsupportedQuery("predicate := must compare one column with PREV(column)")
            if lp:                                                        # PREV(a) op b  ==  b op' PREV(a)
                op = {"<": ">", "<=": ">=", ">": "<", ">=": "<="}.get(op, op)
                return {"cur": rcol, "op": op, "prev": lcol}
            return {"cur": lcol, "op": op, "prev": rcol}
        terms = [comparison()]
        while self.accept("kw", "AND"):
            terms.append(comparison())
        if len(terms) == 1:
            return terms[0]
        # `a = PREV(a) AND b = PREV(b)`: a conjunction of equalities is ONE equality of the column tuples (a tuple with a NULL
        # component is NULL: the conjunction is then not TRUE).  Other conjunctions need a boolean expression evaluator.
        if any(t["op"] != "=" for t in terms):
            raise UnsupportedQuery("predicate := takes one comparison with PREV(column), or a conjunction of equalities")
        return {"cur": [t["cur"] for t in terms], "op": "=", "prev": [t["prev"] for t in terms]}

    def scalar(self):
        if self.at("num"):
            return self.take("num")
        if self.at("str"):
            return self.take("str")
        if self.at("id") and self.at("op", ".", 1):
            return ("col",) + self.column_ref()
        word = self.take("id")
        if word.lower() in ("true", "false"):
            return word.lower() == "true"
        return ("col", None, word)

    def nearest_call(self):
        if self.take("id").upper() != "NEAREST":
            raise UnsupportedQuery("only NEAREST(...) may follow CROSS JOIN LATERAL")
        self.take("op", "(")
        target = self.take("id")
        args = self.named_args()
        self.take("op", ")")
        ref = args.get("reference")
        if not (isinstance(ref, tuple) and ref[0] == "col" and ref[2] == "interval"):
            raise UnsupportedQuery("NEAREST needs reference := <alias>.interval")
        alias = self.alias()
        k = _int_arg(args.get("k"), "NEAREST k", 1)
        if k < 1:
            raise ValueError("NEAREST k must be >= 1")
        return {"target": target, "ref_alias": ref[1], "alias": alias, "k": k, "max_distance": _int_arg(args.get("max_distance"), "NEAREST max_distance"),
                "stranded": _bool_arg(args.get("stranded"), "stranded"), "signed": _bool_arg(args.get("signed"), "signed")}

    def column_ref(self):
        a = self.take("id")
        if self.accept("op", "."):
            if self.accept("op", "*"):
                return (a, "*")
            return (a, self.take("id"))
        return (None, a)

    def distance_call(self):
        self.take("op", "(")
        a = self.column_ref()
        self.take("op", ",")
        b = self.column_ref()
        args = self.named_args()
        self.take("op", ")")
        if a[1] != "interval" or b[1] != "interval":
            raise UnsupportedQuery("DISTANCE takes two <alias>.interval operands")
        return {"a": a[0], "b": b[0], "stranded": _bool_arg(args.get("stranded"), "stranded"), "signed": _bool_arg(args.get("signed"), "signed")}

    # arithmetic over columns and numbers — `end - start`, `0.5 * (a.end - a.start)`, LEAST / GREATEST — as the reference's
    # recipes write it (docs/recipes/*.rst: SUM(end - start) coverage, overlap-fraction filters)
    def at_arith(self) -> bool:
        return self.at("op", "+") or self.at("op", "-") or self.at("op", "*") or self.at("op", "/")

    def expr(self, first=None):
        left = self.term(first)
        while self.at("op", "+") or self.at("op", "-"):
            op = self.take("op")
            left = ("bin", op, left, self.term())
        return left

    def term(self, first=None):
        left = first if first is not None else self.factor()
        while self.at("op", "*") or self.at("op", "/"):
            op = self.take("op")
            left = ("bin", op, left, self.factor())
        return left

    def factor(self):
        if self.accept("op", "-"):
            return ("neg", self.factor())
        if self.accept("op", "("):
            e = self.expr()
            self.take("op", ")")
            return e
        if self.at("num"):
            return ("num", self.take("num"))
        if self.at("id") and self.at("op", "(", 1):
            fn = self.take("id").upper()
            if fn not in ("LEAST", "GREATEST", "ABS"):
                raise UnsupportedQuery(f"function {fn} is not handled by the built-in front-end")
            self.take("op", "(")
            args = [self.expr()]
            while self.accept("op", ","):
                args.append(self.expr())
            self.take("op", ")")
            if fn == "ABS" and len(args) != 1:
                raise ValueError("ABS takes one argument")
            return ("fn", fn, args)
        ref = self.column_ref()
        if ref[1] == "*":
            raise UnsupportedQuery("`*` is not an operand")
        return ("col", ref[0], ref[1])

    def aggregate_args(self, fn: str) -> dict:
        """`( * | [DISTINCT] col )` after COUNT / SUM / MIN / MAX / AVG: the aggregates of the reference's IEJoin whitelist
        (src/giql/expanders/intersects_duckdb.py:618-804 lets GROUP BY + aggregates + HAVING through to the fast path)"""
        self.take("op", "(")
        if fn == "COUNT" and self.accept("op", "*"):
            item = {"kind": "count_star"}
        else:
            distinct = self.accept("kw", "DISTINCT")
            e = self.expr()
            if fn == "STRING_AGG":                                        # STRING_AGG(col, 'sep'): the names merged into a region
                if e[0] != "col" or distinct:
                    raise UnsupportedQuery("STRING_AGG takes a column and a separator")
                self.take("op", ",")
                item = {"kind": "agg", "fn": "STRING_AGG", "alias": e[1], "col": e[2], "sep": self.take("str")}
            elif e[0] != "col":                                           # SUM(end - start)
                if distinct or fn == "COUNT":
                    raise UnsupportedQuery(f"{fn}({'DISTINCT ' if distinct else ''}<expression>) is finished by the SQL engine")
                item = {"kind": "agg", "fn": fn, "alias": None, "col": "<expression>", "expr": e}
            elif fn == "COUNT" and not distinct:
                item = {"kind": "count_col", "alias": e[1], "col": e[2]}
            else:
                if distinct and fn != "COUNT":
                    raise UnsupportedQuery(f"{fn}(DISTINCT ...) is finished by the SQL engine")
                item = {"kind": "agg", "fn": "COUNT_DISTINCT" if distinct else fn, "alias": e[1], "col": e[2]}
        self.take("op", ")")
        return item

    def projection(self):
        if self.accept("op", "*"):
            return {"kind": "star", "alias": None}
        if self.at("str") or (self.at("num") and not (self.at("op", "+", 1) or self.at("op", "-", 1) or self.at("op", "*", 1) or self.at("op", "/", 1))):
            item = {"kind": "const", "value": self.take()}                 # a literal select-list item: `1 AS extra`, `'x' AS tag`
            if self.accept("kw", "AS"):
                item["name"] = self.take("id")
            return item
        if self.at("num") or self.at("op", "(") or self.at("op", "-") or \
                (self.at("id") and self.at("op", "(", 1) and self.peek()[1].upper() in ("LEAST", "GREATEST", "ABS")):
            item = {"kind": "expr", "expr": self.expr()}
            if self.accept("kw", "AS"):
                item["name"] = self.take("id")
            return item
        if self.at("id") and self.at("op", "(", 1):
            fn = self.take("id").upper()
            if fn == "DISTANCE":
                item = {"kind": "distance", **self.distance_call()}
            elif fn in ("CLUSTER", "MERGE"):
                self.take("op", "(")
                col = self.column_ref()
                if col[1] != "interval":
                    raise UnsupportedQuery(f"{fn} takes the genomic column")
                args = self.named_args()
                self.take("op", ")")
                pos = args.get("_positional", [])
                if len(pos) > 1:
                    raise ValueError(f"{fn} takes at most one positional distance")
                item = {"kind": fn.lower(), "distance": _int_arg(pos[0] if pos else args.get("distance"), f"{fn} distance", 0),
                        "stranded": _bool_arg(args.get("stranded"), "stranded"), "predicate": args.get("predicate")}
                if item["predicate"] is not None and not isinstance(item["predicate"], dict):
                    raise ValueError("predicate := must be a comparison")
            elif fn in _AGGREGATES:
                item = self.aggregate_args(fn)
                if self.at("id") and self.peek()[1].upper() == "OVER":
                    # <aggregate> OVER (PARTITION BY cols): the group's value on every row of the group; an ORDER BY inside the
                    # window (running aggregates) is the SQL engine's business
                    self.take("id")
                    self.take("op", "(")
                    part = []
                    if self.at("id") and self.peek()[1].upper() == "PARTITION":
                        self.take("id")
                        self.take("kw", "BY")
                        part.append(self.column_ref())
                        while self.accept("op", ","):
                            part.append(self.column_ref())
                    if not self.at("op", ")"):
                        raise UnsupportedQuery("a window with ORDER BY / a frame is finished by the SQL engine")
                    self.take("op", ")")
                    item = {"kind": "window", "agg": item, "partition": part}
            else:
                raise UnsupportedQuery(f"function {fn} is not handled by the built-in front-end")
        else:
            ref = self.column_ref()
            item = {"kind": "star", "alias": ref[0]} if ref[1] == "*" else {"kind": "col", "alias": ref[0], "col": ref[1]}
            if item["kind"] == "col" and self.at_arith():                  # `end - start AS length`
                item = {"kind": "expr", "expr": self.expr(("col", ref[0], ref[1]))}
        if self.accept("kw", "AS"):
            item["name"] = self.take("id")
        return item

    def condition(self):
        if self.at("num") or self.at("op", "(") or self.at("op", "-") or \
                (self.at("id") and self.at("op", "(", 1) and self.peek()[1].upper() in ("LEAST", "GREATEST", "ABS")):
            return self.expr_comparison(self.expr())                       # `( LEAST(a.end, b.end) - GREATEST(a.start, b.start) ) >= ...`
        if self.accept("kw", "NOT"):
            # `WHERE NOT interval INTERSECTS ANY(...)` (docs/dialect/quantifiers.rst): the literal-range predicates can be negated
            inner = self.condition()
            if inner["kind"] != "literal":
                raise UnsupportedQuery("NOT is bound for predicates against literal ranges")
            inner["negated"] = not inner.get("negated", False)
            return inner
        if self.at("id") and self.peek()[1].upper() == "DISTANCE" and self.at("op", "(", 1):
            self.take("id")
            d = self.distance_call()
            op = self.take("op")
            return {"kind": "distance_cmp", **d, "op": op, "value": self.take("num")}
        left = self.column_ref()
        if self.at_arith():                                               # `disjoin_end - disjoin_start >= 50`
            return self.expr_comparison(self.expr(("col", left[0], left[1])))
        negated_in = self.at("kw", "NOT") and self.at("id", None, 1) and self.peek(1)[1].upper() == "IN"
        if negated_in or (self.at("id") and self.peek()[1].upper() == "IN"):
            if negated_in:
                self.take("kw", "NOT")
            self.take("id")
            self.take("op", "(")
            values = []
            while True:
                if not (self.at("num") or self.at("str")):
                    raise UnsupportedQuery("IN takes a list of literals")
                values.append(self.take())
                if not self.accept("op", ","):
                    break
            self.take("op", ")")
            return {"kind": "in_list", "col": left, "values": values, "negated": bool(negated_in)}
        if self.at("id") and self.peek()[1].upper() == "BETWEEN":
            self.take("id")
            if not (self.at("num") or self.at("str")):
                raise UnsupportedQuery("BETWEEN takes literal bounds")
            lo = self.take()
            self.take("kw", "AND")
            if not (self.at("num") or self.at("str")):
                raise UnsupportedQuery("BETWEEN takes literal bounds")
            return {"kind": "between", "col": left, "lo": lo, "hi": self.take()}
        if self.accept("kw", "IS"):
            neg = self.accept("kw", "NOT")
            self.take("kw", "NULL")
            return {"kind": "is_null", "col": left, "negated": neg}
        if self.at("kw", "INTERSECTS") or self.at("kw", "CONTAINS") or self.at("kw", "WITHIN"):
            op = self.take("kw").lower()
            if left[1] != "interval":
                raise UnsupportedQuery("spatial predicates take the genomic column")
            if self.at("str"):
                return {"kind": "literal", "op": op, "alias": left[0], "quant": None, "ranges": [self.take("str")]}
            if self.at("kw", "ANY") or self.at("kw", "ALL") or self.at("kw", "SOME"):
                quant = self.take("kw")
                self.take("op", "(")
                ranges = [self.take("str")]
                while self.accept("op", ","):
                    ranges.append(self.take("str"))
                self.take("op", ")")
                return {"kind": "literal", "op": op, "alias": left[0], "quant": "ALL" if quant == "ALL" else "ANY", "ranges": ranges}
            right = self.column_ref()
            if right[1] != "interval":
                raise UnsupportedQuery("spatial predicates compare two genomic columns or a column and a literal range")
            return {"kind": "spatial", "op": op, "left": left[0], "right": right[0]}
        op = self.take("op")
        if self.at("str") or (self.at("num") and not (self.at("op", "+", 1) or self.at("op", "-", 1) or self.at("op", "*", 1) or self.at("op", "/", 1))):
            return {"kind": "compare_literal", "op": op, "col": left, "value": self.take()}
        if self.at("num") or self.at("op", "(") or self.at("op", "-") or (self.at("id") and self.at("op", "(", 1)):
            return {"kind": "compare_expr", "op": op, "left": ("col", left[0], left[1]), "right": self.expr()}
        right = self.column_ref()
        if self.at_arith():
            return {"kind": "compare_expr", "op": op, "left": ("col", left[0], left[1]), "right": self.expr(("col", right[0], right[1]))}
        return {"kind": "compare", "op": op, "left": left, "right": right}

    def expr_comparison(self, left):
        op = self.take("op")
        if op in ("+", "-", "*", "/", "(", ")", ",", "."):
            raise UnsupportedQuery(f"expected a comparison, found {op!r}")
        return {"kind": "compare_expr", "op": op, "left": left, "right": self.expr()}


def parse(query: str) -> dict:
    p = _Parser(_tokenize(query))
    q = p.query()
    if not p.at("eof"):
        raise UnsupportedQuery(f"unexpected {p.peek()[1]!r}")
    return q


# ---- binder / evaluator ---------------------------------------------------------------------------------------------------
_CMP = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal, "=": np.equal, "!=": np.not_equal}


def _column(values) -> np.ndarray:
    """host column: int64 / float / bool arrays stay typed, anything holding strings or NULLs becomes an object array"""
    if isinstance(values, np.ndarray) and values.dtype != object and values.dtype.kind not in "US":
        return values
    if hasattr(values, "to_numpy") and not isinstance(values, np.ndarray):
        values = values.to_numpy()
    seq = values.tolist() if isinstance(values, np.ndarray) else list(values)
    if seq and all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in seq):
        return np.asarray(seq, dtype=np.int64)
    out = np.empty(len(seq), dtype=object)
    out[:] = seq
    return out


class Result:
    """Columnar result: ``columns`` (names, duplicates allowed as in SQL) and ``arrays``; ``rows`` = list of tuples."""

    def __init__(self, columns: list[str], arrays: list[np.ndarray], sources: Optional[list] = None):
        # sources[i] = (relation alias, column) an output column was projected from, None for computed columns: what a
        # qualified ORDER BY key (``ORDER BY b.start``) resolves against when output names repeat
        self.columns, self.arrays = columns, arrays
        self.sources = sources if sources is not None else [None] * len(columns)

    @property
    def rows(self) -> list[tuple]:
        if not self.arrays:
            return []
        return list(zip(*(a.tolist() for a in self.arrays)))

    def take(self, sel) -> "Result":
        return Result(self.columns, [a[sel] for a in self.arrays], self.sources)

    def to_pydict(self) -> dict:
        return {c: a.tolist() for c, a in zip(self.columns, self.arrays)}

    def to_arrow(self):
        """pyarrow Table (duplicate column names, legal in SQL, get a numeric suffix)"""
        import pyarrow as pa

        names, seen = [], {}
        for c in self.columns:
            seen[c] = seen.get(c, 0) + 1
            names.append(c if seen[c] == 1 else f"{c}_{seen[c] - 1}")
        return pa.table({n: pa.array(a.tolist()) for n, a in zip(names, self.arrays)})

    def to_pandas(self):
        return self.to_arrow().to_pandas()

    def __iter__(self):
        return iter(self.rows)

    def __len__(self):
        return len(self.arrays[0]) if self.arrays else 0


class _Rel:
    """A host relation: ordered column names + equally long numpy columns, and the Table config it came from."""

    def __init__(self, cols: list[str], data: dict, config: Optional[Table] = None):
        self.cols, self.data, self.config = cols, data, config
        self._genomic = None

    def __len__(self):
        return len(self.data[self.cols[0]]) if self.cols else 0

    def genomic(self) -> dict:
        """the dict-table view the backends take (cached: the engine backend parks its device copy on it)"""
        if self._genomic is None:
            c = self.config or Table("t")
            for col in (c.chrom_col, c.start_col, c.end_col):
                if col not in self.data:
                    raise ValueError(f"genomic column {col!r} is missing")
            g = {"chrom": self.data[c.chrom_col], "start": self.data[c.start_col], "end": self.data[c.end_col]}
            if c.strand_col and c.strand_col in self.data:
                g["strand"] = self.data[c.strand_col]
            self._genomic = g
        return self._genomic


class _Frame:
    """Joined rows in columnar form: per alias an int64 row-index array into its relation (-1 = the NULL row of an
    outer join) plus computed columns (NEAREST's distance, DISJOIN's segment)."""

    def __init__(self):
        self.order: list[str] = []
        self.rels: dict[str, _Rel] = {}
        self.idx: dict[str, np.ndarray] = {}
        self.extra: dict[str, dict[str, np.ndarray]] = {}

    def add(self, alias, rel, idx, extra=None):
        if alias in self.rels:
            raise UnsupportedQuery(f"alias {alias!r} is used twice")
        self.order.append(alias)
        self.rels[alias], self.idx[alias], self.extra[alias] = rel, idx, dict(extra or {})

    def __len__(self):
        return len(self.idx[self.order[0]])

    def take(self, sel):
        for a in self.order:
            self.idx[a] = self.idx[a][sel]
            self.extra[a] = {c: v[sel] for c, v in self.extra[a].items()}

    def schema(self, alias) -> list[str]:
        return list(self.rels[alias].cols) + list(self.extra[alias])

    def resolve(self, alias, col) -> str:
        for a in ([alias] if alias else self.order):
            if a not in self.rels:
                raise ValueError(f"unknown alias {a!r}")
            if col in self.extra[a] or col in self.rels[a].data:
                return a
        raise ValueError(f"unknown column {col!r}")

    def column(self, alias, col) -> np.ndarray:
        if col in self.extra[alias]:
            return self.extra[alias][col]
        src, ix = self.rels[alias].data[col], self.idx[alias]
        if len(ix) and ix.min() < 0:
            out = np.empty(len(ix), dtype=object)
            ok = ix >= 0
            out[ok] = src[ix[ok]].tolist()
            out[~ok] = None
            return out
        return src[ix]

    def genomic(self, alias) -> dict:
        g = self.rels[alias].genomic()
        return {k: v[self.idx[alias]] for k, v in g.items() if not k.startswith("__")}


def _expand(frame_rows: np.ndarray, l: np.ndarray, r: np.ndarray, keep_unmatched: bool):
    """Join the frame (its per-row id ``frame_rows`` into the left relation) with the pair list (l, r): returns
    (frame row to repeat, matched right row or -1 for the NULL row of a LEFT join)."""
    l, r = np.asarray(l, dtype=np.int64), np.asarray(r, dtype=np.int64)
    order = np.lexsort((r, l))
    l, r = l[order], r[order]
    lo, hi = np.searchsorted(l, frame_rows, "left"), np.searchsorted(l, frame_rows, "right")
    cnt = hi - lo
    out = np.maximum(cnt, 1) if keep_unmatched else cnt
    rep = np.repeat(np.arange(len(frame_rows)), out)
    within = np.arange(len(rep)) - np.repeat(np.cumsum(out) - out, out)
    pos = np.repeat(lo, out) + within
    matched = np.repeat(cnt, out) > 0
    right = np.where(matched, r[np.minimum(pos, max(len(r) - 1, 0))] if len(r) else -1, -1)
    return rep, right


def _pairs(be, L, R, op, lcfg, rcfg, residuals=()):
    if hasattr(be, "pair_arrays"):
        return be.pair_arrays(L, R, op, lcfg, rcfg, residuals) if residuals else be.pair_arrays(L, R, op, lcfg, rcfg)
    p = be.pairs(L, R, lcfg, rcfg) if op == "intersects" else be.spatial_pairs(L, R, op, lcfg, rcfg)
    a = np.asarray(p, dtype=np.int64).reshape(-1, 2)
    l, r = a[:, 0], a[:, 1]
    for x in residuals:                                                  # a backend without a device filter: same rule on the host
        keep = _CMP[x["op"]](x["l"][l], x["r"][r])
        if x.get("lvalid") is not None:
            keep &= x["lvalid"][l].astype(bool)
        if x.get("rvalid") is not None:
            keep &= x["rvalid"][r].astype(bool)
        l, r = l[keep], r[keep]
    return l, r


def _residual(left: _Rel, right: _Rel, la, ra, cond) -> dict:
    """One `a.col <op> b.col` ON-conjunct as comparable per-row columns: numbers as they are, anything else as the rank
    of its value in the sorted union of both sides (equality and order preserved); NULLs through validity masks."""
    (a1, c1), (a2, c2), op = cond["left"], cond["right"], cond["op"]
    op = {"<>": "!=", "==": "="}.get(op, op)
    if op not in _CMP:
        raise UnsupportedQuery(f"comparison {op}")
    if {a1, a2} != {la, ra}:
        raise UnsupportedQuery("a residual join predicate must compare one column of each side")
    if a1 == ra:                                                         # written `b.col <op> a.col`: mirror it
        (a1, c1), (a2, c2) = (a2, c2), (a1, c1)
        op = {"<": ">", "<=": ">=", ">": "<", ">=": "<="}.get(op, op)
    for rel, c in ((left, c1), (right, c2)):
        if c not in rel.data:
            raise ValueError(f"unknown column {c!r}")
    lc, rc = left.data[c1], right.data[c2]
    out = {"op": op}
    if lc.dtype != object and rc.dtype != object:
        out["l"], out["r"] = lc, rc
        return out
    lv = np.array([v is not None for v in lc], dtype=bool) if lc.dtype == object else np.ones(len(lc), bool)
    rv = np.array([v is not None for v in rc], dtype=bool) if rc.dtype == object else np.ones(len(rc), bool)
    try:
        values = sorted(set(lc[lv].tolist()) | set(rc[rv].tolist()))
    except TypeError as exc:                                             # mixed incomparable types: a binder error in SQL
        raise ValueError(f"cannot compare {c1!r} with {c2!r}: {exc}") from None
    rank = {v: i for i, v in enumerate(values)}
    out["l"] = np.fromiter((rank[v] if ok else 0 for v, ok in zip(lc.tolist(), lv)), dtype=np.int32, count=len(lc))
    out["r"] = np.fromiter((rank[v] if ok else 0 for v, ok in zip(rc.tolist(), rv)), dtype=np.int32, count=len(rc))
    if not lv.all():
        out["lvalid"] = lv.astype(np.uint8)
    if not rv.all():
        out["rvalid"] = rv.astype(np.uint8)
    return out


def _exists(be, L, R, lcfg, rcfg):
    return be.exists_array(L, R, lcfg, rcfg) if hasattr(be, "exists_array") else np.asarray(be.exists(L, R, lcfg, rcfg), dtype=bool)


def _counts(be, L, R, lcfg, rcfg):
    return be.counts_array(L, R, lcfg, rcfg) if hasattr(be, "counts_array") else np.asarray(be.counts(L, R, lcfg, rcfg), dtype=np.int64)


def _distance(be, a: _Rel, b: _Rel, ia, ib, stranded, signed):
    """object column of distances (None = SQL NULL) for the paired rows"""
    if not len(ia):
        return np.empty(0, dtype=object)
    if hasattr(be, "distance_arrays"):
        d, v = be.distance_arrays(a.genomic(), b.genomic(), ia, ib, a.config, b.config, stranded=stranded, signed=signed)
        if v.all():
            return d
        out = np.empty(len(d), dtype=object)
        out[:] = d.tolist()
        out[~v] = None
        return out
    ga, gb = a.genomic(), b.genomic()
    pa = {c: x[ia].tolist() for c, x in ga.items() if not c.startswith("__")}
    pb = {c: x[ib].tolist() for c, x in gb.items() if not c.startswith("__")}
    return _column(be.distance(pa, pb, a.config, b.config, stranded=stranded, signed=signed))


def _nearest(be, Q: _Rel, T: _Rel, j):
    kw = dict(k=j["k"], max_distance=j["max_distance"], stranded=j["stranded"], signed=j["signed"])
    if hasattr(be, "nearest_arrays"):
        q, t, d, v = be.nearest_arrays(Q.genomic(), T.genomic(), Q.config, T.config, **kw)
        dist = d.astype(np.int64)
        if not v.all():
            dist = _column([x if ok else None for x, ok in zip(d.tolist(), v.tolist())])
        return q.astype(np.int64), t.astype(np.int64), dist
    res = be.nearest(Q.genomic(), T.genomic(), Q.config, T.config, **kw)
    return (np.asarray([x[0] for x in res], dtype=np.int64), np.asarray([x[1] for x in res], dtype=np.int64), _column([x[2] for x in res]))


def _cluster_ids(be, f, p) -> np.ndarray:
    """CLUSTER ids of the frame's rows (src/giql/expanders/cluster.py:205-310, 430-434), with the optional
    `predicate := cur <op> PREV(prev)` gate (:281-299)."""
    a = f.order[0]
    g = f.genomic(a)
    if p["stranded"] and "strand" not in g:
        raise ValueError("stranded := true needs a strand column")
    if not p["stranded"]:
        g.pop("strand", None)
    if not len(f):
        return np.zeros(0, np.int64)
    pred = None
    if p.get("predicate"):
        q = p["predicate"]

        def operand(cols):
            if not isinstance(cols, list):
                return f.column(f.resolve(None, cols), cols).tolist()
            parts = [f.column(f.resolve(None, c), c).tolist() for c in cols]
            return [None if any(v is None for v in t) else t for t in zip(*parts)]
        pred = (operand(q["cur"]), q["op"], operand(q["prev"]))
    if hasattr(be, "cluster_array"):
        return be.cluster_array(g, p["distance"], p["stranded"], pred).astype(np.int64) if pred else be.cluster_array(g, p["distance"], p["stranded"]).astype(np.int64)
    lists = {k: v.tolist() for k, v in g.items()}
    ids = be.cluster(lists, p["distance"], p["stranded"], pred) if pred else be.cluster(lists, p["distance"], p["stranded"])
    return np.asarray(ids, dtype=np.int64)


def _eval_expr(f: "_Frame", e) -> np.ndarray:
    """An arithmetic expression over the frame's rows.  NULL operands give NULL (object column with None); integer arithmetic
    stays in int64, `/` is a DOUBLE division as in DuckDB."""
    kind = e[0]
    if kind == "num":
        return np.full(len(f), e[1], dtype=np.float64 if isinstance(e[1], float) else np.int64)
    if kind == "col":
        return f.column(f.resolve(e[1], e[2]), e[2])
    if kind == "neg":
        return _arith("-", np.zeros(len(f), np.int64), _eval_expr(f, e[1]))
    if kind == "bin":
        return _arith(e[1], _eval_expr(f, e[2]), _eval_expr(f, e[3]))
    if kind == "fn":
        args = [_eval_expr(f, a) for a in e[2]]
        if e[1] == "ABS":
            return _arith("abs", args[0], args[0])
        out = args[0]
        for a in args[1:]:
            out = _arith("min" if e[1] == "LEAST" else "max", out, a)
        return out
    raise UnsupportedQuery(f"expression node {kind}")


def _arith(op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    fns = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.true_divide, "min": np.minimum, "max": np.maximum,
           "abs": lambda x, _: np.abs(x)}
    if a.dtype != object and b.dtype != object:
        if a.dtype.kind not in "iuf" or b.dtype.kind not in "iuf":
            raise ValueError(f"arithmetic needs numeric operands, got {a.dtype} and {b.dtype}")
        return fns[op](a, b)
    py = {"+": lambda x, y: x + y, "-": lambda x, y: x - y, "*": lambda x, y: x * y, "/": lambda x, y: x / y, "min": min, "max": max,
          "abs": lambda x, _: abs(x)}[op]
    try:                                                                  # NULL in, NULL out
        return _column([None if x is None or y is None else py(x, y) for x, y in zip(a.tolist(), b.tolist())])
    except TypeError as exc:
        raise ValueError(f"arithmetic needs numeric operands: {exc}") from None


def _aggregate(f: "_Frame", p, members) -> np.ndarray:
    """one value per group of frame rows (``members``: row-index arrays) for a COUNT / SUM / MIN / MAX / AVG item"""
    if p["kind"] == "count_star":
        return np.asarray([len(v) for v in members], dtype=np.int64)
    vals = _eval_expr(f, p["expr"]) if p.get("expr") is not None else f.column(f.resolve(p["alias"], p["col"]), p["col"])
    if p["kind"] == "count_col":
        present = np.array([v is not None for v in vals], dtype=bool) if vals.dtype == object else np.ones(len(vals), bool)
        return np.asarray([int(present[v].sum()) for v in members], dtype=np.int64)
    if p["fn"] in ("SUM", "AVG") and vals.dtype.kind not in "iufb" and any(isinstance(x, str) for x in vals.tolist()):
        raise ValueError(f"{p['fn']} needs a numeric column, {p['col']!r} holds strings")
    out = []
    for v in members:                                         # NULLs are skipped; no value at all -> NULL (COUNT: 0)
        xs = [x for x in vals[v].tolist() if x is not None]
        fn = p["fn"]
        if fn == "COUNT_DISTINCT":
            out.append(len(set(xs)))
        elif fn == "STRING_AGG":
            out.append(p["sep"].join(str(x) for x in xs) if xs else None)
        elif not xs:
            out.append(None)
        elif fn == "SUM":
            out.append(sum(xs))
        elif fn == "MIN":
            out.append(min(xs))
        elif fn == "MAX":
            out.append(max(xs))
        else:
            out.append(sum(xs) / len(xs))                     # AVG is a DOUBLE
    if p["fn"] == "AVG" and all(x is not None for x in out):
        return np.asarray(out, dtype=np.float64)
    return _column(out)


class Session:
    """Registered tables + ``execute(giql_query)``.  ``backend`` defaults to the CUDA engine on device 0."""

    def __init__(self, backend=None):
        self._backend = backend
        self.tables: dict[str, _Rel] = {}

    @property
    def backend(self):
        if self._backend is None:
            from .engine import Engine

            self._backend = EngineBackend(Engine(0))
        return self._backend

    def register(self, name: str, data, table: Optional[Table] = None, columns: Optional[list[str]] = None):
        """``data``: mapping column -> sequence / numpy array (a pandas DataFrame or pyarrow Table works too).
        ``table``: the reference's column-mapping contract (src/giql/table.py:16-93); defaults to chrom/start/end/strand."""
        if hasattr(data, "to_pydict"):
            data = data.to_pydict()
        elif hasattr(data, "to_dict") and not isinstance(data, dict):
            data = data.to_dict(orient="list")
        cols = columns or list(data.keys())
        rel = _Rel(cols, {c: _column(data[c]) for c in cols}, table)
        if len({len(v) for v in rel.data.values()}) > 1:
            raise ValueError("columns differ in length")
        self.tables[name] = rel

    def register_bed(self, name: str, source, fields: Optional[int] = None):
        """Register a BED file (path or bytes), parsed on the device (``Engine.read_bed``); columns are named chrom, start,
        end, name, score, strand as far as the file has them."""
        be = self.backend
        if not isinstance(be, EngineBackend):
            raise UnsupportedQuery("register_bed needs the CUDA engine backend")
        t = be.e.read_bed(source, fields)
        import pyarrow as pa

        def strings(offs, byts):
            o, b = offs.cpu().numpy(), byts.cpu().numpy()
            return pa.StringArray.from_buffers(len(o) - 1, pa.py_buffer(o), pa.py_buffer(b)).to_numpy(zero_copy_only=False)

        data = {"chrom": strings(t.chrom_offsets, t.chrom_bytes), "start": t.start.cpu().numpy().astype(np.int64), "end": t.end.cpu().numpy().astype(np.int64)}
        if t.bed_name is not None:
            data["name"] = strings(*t.bed_name)
        if t.bed_score is not None:
            data["score"] = t.bed_score.cpu().numpy().astype(np.int64)
        if t.strand is not None:
            data["strand"] = np.array([be.e.strand_literal(c) for c in range(3)], dtype=object)[t.strand.cpu().numpy()]
        self.register(name, data)
        cfg = Table("t")
        key = (id(be.e), cfg.chrom_col, cfg.start_col, cfg.end_col, cfg.strand_col, cfg.coordinate_system, cfg.interval_type)
        self.tables[name].genomic().setdefault("__device__", {})[key] = t      # the parsed columns are already on the device

    def execute(self, query: str) -> Result:
        return self._run(parse(query))

    # -- sources
    @staticmethod
    def _derived(res: Result) -> _Rel:
        if len(set(res.columns)) != len(res.columns):
            raise UnsupportedQuery("a derived table with duplicate column names is ambiguous")
        return _Rel(list(res.columns), dict(zip(res.columns, res.arrays)))

    def _relation(self, q, src) -> _Rel:
        if src["kind"] == "subquery":
            return self._derived(self._run(src["query"]))
        name = src["table"]
        if name in q["ctes"]:
            return self._derived(self._run(q["ctes"][name]))
        if name not in self.tables:
            raise ValueError(f"unknown table {name!r}")
        return self.tables[name]

    def _run(self, q) -> Result:
        if "setop" in q:
            return self._set_operation(q)
        be = self.backend
        src = q["source"]
        f = _Frame()
        if src["kind"] == "disjoin":
            # DISJOIN(t[, ref])   (src/giql/expanders/disjoin.py:102-202; output in the target's encoding, :139-141)
            t = self._relation(q, {"kind": "table", "table": src["table"]})
            if isinstance(src["reference"], dict):
                sub = src["reference"]["subquery"]
                sub.setdefault("ctes", {})
                sub["ctes"] = {**q["ctes"], **sub["ctes"]}
                ref = self._derived(self._run(sub))
            else:
                ref = self._relation(q, {"kind": "table", "table": src["reference"]}) if src["reference"] else None
            sadj, eadj = canonical_offsets(t.config)
            segs = []
            if len(t) and (ref is None or len(ref)):
                segs = be.disjoin(t.genomic(), None if ref is None else ref.genomic(), t.config, None if ref is None else ref.config)
            rows = np.asarray([s[0] for s in segs], dtype=np.int64)
            f.add(src["alias"], t, rows, {
                "disjoin_chrom": t.genomic()["chrom"][rows],
                "disjoin_start": np.asarray([s[1] - sadj for s in segs], dtype=np.int64),
                "disjoin_end": np.asarray([s[2] - eadj for s in segs], dtype=np.int64)})
        elif src["kind"] == "nearest_literal":
            tgt = self._relation(q, {"kind": "table", "table": src["table"]})
            chrom, start, end = parse_range(src["range"])
            one = _Rel(["chrom", "start", "end"], {"chrom": _column([chrom]), "start": _column([start]), "end": _column([end])})
            ti, dist = np.zeros(0, np.int64), np.zeros(0, np.int64)
            if len(tgt):
                _, ti, dist = _nearest(be, one, tgt, src)              # one query row: the engine's rank order is the output order
            f.add(src["alias"], tgt, ti, {"distance": dist})
        else:
            base = self._relation(q, src)
            f.add(src["alias"], base, np.arange(len(base), dtype=np.int64))
        where = list(q["where"])
        fast = self._count_overlaps(q, f, where)
        if fast is not None:
            return fast
        for j in q["joins"]:
            if j["kind"] == "nearest":
                self._join_nearest(q, f, j)
                continue
            if j["source"]["kind"] == "disjoin":                      # `... JOIN DISJOIN(t[, ref]) AS d ON ...`: the segments as a relation
                right = self._derived(self._run({"distinct": False, "proj": [{"kind": "star", "alias": None}], "joins": [], "where": [],
                                                 "order": [], "limit": None, "offset": 0, "group": [], "source": j["source"], "ctes": q["ctes"]}))
            else:
                right = self._relation(q, j["source"])
            ra, la = j["source"]["alias"], f.order[0]
            left = f.rels[la]
            conds, kind = list(j.get("on") or []), j["kind"]
            comma = kind == "comma"
            if kind == "comma":
                # `FROM a, b WHERE a.interval INTERSECTS b.interval [AND a.x = b.x]`: the WHERE conjuncts over both tables are the ON
                between = lambda w: (w["kind"] == "spatial" and {w["left"], w["right"]} == {la, ra}) or \
                    (w["kind"] == "compare" and {w["left"][0], w["right"][0]} == {la, ra})
                conds, kind = [w for w in where if between(w)], "inner"
                where = [w for w in where if not between(w)]
            single_side = [c for c in conds if c["kind"] in ("compare_literal", "between", "in_list", "is_null", "compare_expr")]
            if single_side and kind == "inner":
                # `ON a.interval INTERSECTS b.interval AND a.score > 10`: on an INNER join a conjunct over one row (or an expression
                # of the joined row) filters the joined rows exactly like WHERE does (intersects_duckdb.py:1239-1243 keeps such
                # residuals inside every per-chromosome branch); outer / SEMI / ANTI joins change meaning and stay declined
                where = where + single_side
                conds = [c for c in conds if c not in single_side]
            spatial = [c for c in conds if c["kind"] == "spatial"]
            compares = [c for c in conds if c["kind"] == "compare"]
            if len(spatial) + len(compares) != len(conds) or len(spatial) > 1:
                raise UnsupportedQuery("JOIN ... ON takes one spatial predicate and column comparisons between the two tables")
            if comma and not conds:
                # plain cross join (`FROM a, b` feeding DISTANCE in the select list, docs/dialect/distance-operators.rst): every
                # pairing is an output row, so the product is built as two row-id columns; large products belong to the SQL engine
                if len(f) * len(right) > _CROSS_JOIN_LIMIT:
                    raise UnsupportedQuery(f"a cross join of {len(f)} x {len(right)} rows is finished by the SQL engine")
                n_left = len(f)
                f.take(np.repeat(np.arange(n_left), len(right)))
                f.add(ra, right, np.tile(np.arange(len(right), dtype=np.int64), n_left))
                continue
            overlap = self._literal_overlap_join(f, right, ra, compares) if not spatial and kind in ("inner", "left") else None
            if overlap is not None:
                # `ON x.chrom = y.chrom AND x.start < y.end AND x.end > y.start` over any three columns of each side: the predicate
                # the generic expander writes for INTERSECTS (src/giql/expanders/intersects.py:149-154), spelled out by the user —
                # e.g. DISJOIN's segments joined back to an annotation table.  Raw columns: no coordinate canonicalisation.
                Lg, Rg = overlap
                l, r = _pairs(be, Lg, Rg, "intersects", None, None) if len(f) and len(right) else (np.zeros(0, np.int64),) * 2
                rep, rrow = _expand(np.arange(len(f), dtype=np.int64), l, r, keep_unmatched=(kind == "left"))
                f.take(rep)
                f.add(ra, right, rrow)
                continue
            if not spatial:
                # chromosome equi-join feeding DISTANCE (docs/dialect/distance-operators.rst:130-155): quadratic by nature
                on = compares[0] if len(compares) == 1 else None
                chrom = (left.config or Table("t")).chrom_col
                if not (on and on["op"] == "=" and kind == "inner" and on["left"][1] == chrom and on["right"][1] == (right.config or Table("t")).chrom_col
                        and {on["left"][0], on["right"][0]} == {la, ra}):
                    raise UnsupportedQuery("without a spatial predicate only the `a.chrom = b.chrom` equi-join is handled")
                lc, rc = f.genomic(la)["chrom"], right.genomic()["chrom"]
                _, codes = np.unique(np.concatenate([lc.astype(str), rc.astype(str)]), return_inverse=True)
                cl, cr = codes[: len(lc)], codes[len(lc):]
                by = np.argsort(cr, kind="stable")
                lo, hi = np.searchsorted(cr[by], cl, "left"), np.searchsorted(cr[by], cl, "right")
                cnt = hi - lo
                rep = np.repeat(np.arange(len(lc)), cnt)
                pos = np.repeat(lo, cnt) + np.arange(len(rep)) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                f.take(rep)
                f.add(ra, right, by[pos].astype(np.int64))
                continue
            on = spatial[0]
            if {on["left"], on["right"]} != {la, ra}:
                raise UnsupportedQuery("JOIN ... ON needs `<left>.interval <op> <right>.interval`")
            residuals = [_residual(left, right, la, ra, c) for c in compares]      # intersects_duckdb.py:1239-1243
            op = on["op"]
            if on["left"] == ra and op != "intersects":                 # `b.interval CONTAINS a.interval` == a WITHIN b
                op = "within" if op == "contains" else "contains"
            nonempty = len(left) and len(right)
            if kind in ("semi", "anti") and not residuals:
                # src/giql/expanders/intersects_duckdb.py:1254-1282: one row per qualifying LEFT row; ANTI keeps left-only chroms
                if op != "intersects":
                    raise UnsupportedQuery("SEMI / ANTI joins are bound for INTERSECTS only")
                flags = _exists(be, left.genomic(), right.genomic(), left.config, right.config) if nonempty else np.zeros(len(left), bool)
                f.take(flags[f.idx[la]] == (kind == "semi"))
                continue
            # src/giql/expanders/intersects.py:149-154 (INTERSECTS), :155-166 (CONTAINS / WITHIN)
            l, r = _pairs(be, left.genomic(), right.genomic(), op, left.config, right.config, residuals) if nonempty else (np.zeros(0, np.int64),) * 2
            if kind in ("semi", "anti"):                                 # [NOT] EXISTS over the residual-filtered pairs
                flags = np.zeros(len(left), bool)
                flags[np.asarray(l, dtype=np.int64)] = True
                f.take(flags[f.idx[la]] == (kind == "semi"))
                continue
            rep, rrow = _expand(f.idx[la], l, r, keep_unmatched=(kind == "left"))
            f.take(rep)
            f.add(ra, right, rrow)
        for w in where:
            self._filter(f, w)
        return self._finish(q, f)

    def _set_operation(self, q) -> Result:
        """UNION [ALL] / EXCEPT / INTERSECT of two results (bag for UNION ALL, sets otherwise, as in SQL); names come from the left."""
        a, b = self._run(q["left"]), self._run(q["right"])
        if len(a.columns) != len(b.columns):
            raise ValueError(f"{q['setop']}: the two queries return {len(a.columns)} and {len(b.columns)} columns")
        ra, rb = a.rows, b.rows
        if q["setop"] == "UNION":
            rows = ra + rb if q["all"] else list(dict.fromkeys(ra + rb))
        elif q["setop"] == "EXCEPT":
            gone = set(rb)
            rows = [r for r in dict.fromkeys(ra) if r not in gone]
        else:
            both = set(rb)
            rows = [r for r in dict.fromkeys(ra) if r in both]
        res = Result(list(a.columns), [_column([r[i] for r in rows]) for i in range(len(a.columns))])
        return self._order_limit(q, res, None)

    @staticmethod
    def _literal_overlap_join(f: _Frame, right: _Rel, ra: str, compares):
        """Recognises the three-conjunct overlap predicate between the rows joined so far and table ``ra``; returns the two
        (chrom, start, end) column sets — the left one aligned with the frame's rows — or None."""
        if len(compares) != 3:
            return None
        flip = {"<": ">", ">": "<", "=": "="}
        eq = lt = gt = None
        for c in compares:
            (la_, lc), (ra_, rc), op = c["left"], c["right"], c["op"]
            if op not in flip or la_ is None or ra_ is None:
                return None
            if la_ == ra and ra_ != ra:                                  # orient: frame side first
                la_, lc, ra_, rc, op = ra_, rc, la_, lc, flip[op]
            if ra_ != ra or la_ == ra or la_ not in f.rels or rc not in right.data:
                return None
            if not (lc in f.extra[la_] or lc in f.rels[la_].data):
                return None
            slot = {"=": "eq", "<": "lt", ">": "gt"}[op]
            if slot == "eq" and eq is None:
                eq = (la_, lc, rc)
            elif slot == "lt" and lt is None:
                lt = (la_, lc, rc)                                       # frame.start < right.end
            elif slot == "gt" and gt is None:
                gt = (la_, lc, rc)                                       # frame.end > right.start
            else:
                return None
        if eq is None or lt is None or gt is None:
            return None
        Lg = {"chrom": f.column(eq[0], eq[1]), "start": f.column(lt[0], lt[1]), "end": f.column(gt[0], gt[1])}
        Rg = {"chrom": right.data[eq[2]], "start": right.data[gt[2]], "end": right.data[lt[2]]}
        for g in (Lg, Rg):
            if len(g["chrom"]) and (g["start"].dtype.kind not in "iu" or g["end"].dtype.kind not in "iu" or any(v is None for v in g["chrom"].tolist())):
                return None                                              # NULLs or non-integer coordinates: the SQL engine's business
        return Lg, Rg

    def _count_overlaps(self, q, f: _Frame, where) -> Optional[Result]:
        """The count_overlaps shape (src/giql/expanders/intersects_duckdb.py:432-548, 806-854):

            SELECT <left columns>, COUNT(*) | COUNT(b.col) FROM a [LEFT] JOIN b ON a.interval INTERSECTS b.interval
            GROUP BY <the same left columns>

        bound to ONE per-left-row count probe (giqlb_intersects_count_rows) — no pair list is materialised — followed by the
        sum over equal left keys.  Returns None when the statement is not of this shape (the generic path then groups the
        joined rows)."""
        if not q.get("group") or len(q["joins"]) != 1 or where or q["distinct"] or len(f.order) != 1:
            return None
        j = q["joins"][0]
        if j["kind"] not in ("inner", "left") or j["source"]["kind"] != "table":
            return None
        conds = j.get("on") or []
        la, ra = f.order[0], j["source"]["alias"]
        if len(conds) != 1 or conds[0]["kind"] != "spatial" or conds[0]["op"] != "intersects" or {conds[0]["left"], conds[0]["right"]} != {la, ra}:
            return None
        left, right = f.rels[la], self._relation(q, j["source"])
        keys, counters = [], []
        for p in q["proj"]:
            if p["kind"] == "col" and (p["alias"] in (None, la)) and p["col"] in left.data:
                keys.append(p)
            elif p["kind"] == "count_star" or (p["kind"] == "count_col" and p["alias"] == ra and p["col"] in right.data
                                               and right.data[p["col"]].dtype != object):
                counters.append(p)                                    # COUNT(b.col) over a NULL-free column counts the matches
            else:
                return None
        group_cols = [c for a, c in q["group"] if a in (None, la)]
        if len(group_cols) != len(q["group"]) or len(counters) != 1 or sorted(group_cols) != sorted(p["col"] for p in keys) or not keys:
            return None
        be = self.backend
        n = len(left)
        per_row = _counts(be, left.genomic(), right.genomic(), left.config, right.config).astype(np.int64) if n and len(right) else np.zeros(n, np.int64)
        cnt = counters[0]
        if j["kind"] == "left" and cnt["kind"] == "count_star":
            per_row = np.maximum(per_row, 1)                            # the NULL-extended row of an unmatched left row counts once
        cols = [left.data[p["col"]] for p in keys]
        first: dict[tuple, int] = {}
        total: list[int] = []
        for i, key in enumerate(zip(*(c.tolist() for c in cols))):
            g = first.setdefault(key, len(total))
            if g == len(total):
                total.append(0)
            total[g] += int(per_row[i])
        rows = [k + (total[g],) for k, g in first.items() if j["kind"] == "left" or total[g] > 0]
        names = {id(p): p.get("name", p["col"]) for p in keys}
        out_cols, arrays, sources = [], [], []
        for p in q["proj"]:
            if p is cnt:
                out_cols.append(p.get("name", "count"))
                arrays.append(np.asarray([r[-1] for r in rows], dtype=np.int64))
                sources.append(None)
            else:
                k = keys.index(p)
                out_cols.append(names[id(p)])
                arrays.append(_column([r[k] for r in rows]))
                sources.append((la, p["col"]))
        return self._order_limit(q, Result(out_cols, arrays, sources), f)

    def _group_count(self, q, f: _Frame) -> Result:
        """generic GROUP BY <columns> with COUNT / SUM / MIN / MAX / AVG over the joined rows, and HAVING over the groups (the
        aggregation the reference's fast path leaves to DuckDB around the pair relation, intersects_duckdb.py:1339-1400)"""
        keys = [(f.resolve(a, c), c) for a, c in q["group"]]
        key_cols = [f.column(a, c).tolist() for a, c in keys]
        groups: dict[tuple, list[int]] = {}
        for i, key in enumerate(zip(*key_cols)) if key_cols else []:
            groups.setdefault(key, []).append(i)
        if not keys:
            groups[()] = list(range(len(f)))                               # the global aggregate: one row, also over no rows
        members = [np.asarray(v, dtype=np.int64) for v in groups.values()]

        aggregate = lambda p: _aggregate(f, p, members)

        out_cols, arrays, sources = [], [], []
        for p in q["proj"]:
            sources.append(None)
            if p["kind"] == "col":
                a = f.resolve(p["alias"], p["col"])
                if (a, p["col"]) not in keys:
                    raise UnsupportedQuery(f"column {p['col']!r} must appear in GROUP BY")
                k = keys.index((a, p["col"]))
                out_cols.append(p.get("name", p["col"]))
                arrays.append(_column([key[k] for key in groups]))
                sources[-1] = (a, p["col"])
            elif p["kind"] in ("count_star", "count_col", "agg"):
                default = "count" if p["kind"] != "agg" else p["fn"].lower()
                out_cols.append(p.get("name", default))
                arrays.append(aggregate(p))
            else:
                raise UnsupportedQuery("GROUP BY supports grouped columns and COUNT / SUM / MIN / MAX / AVG; other expressions are finished by the SQL engine")
        res = Result(out_cols, arrays, sources)
        for h in q.get("having") or []:
            if h["op"] not in _CMP:
                raise UnsupportedQuery(f"HAVING comparison {h['op']}")
            vals = aggregate(h["item"])
            if vals.dtype == object:
                keep = np.array([x is not None and bool(_CMP[h["op"]](x, h["value"])) for x in vals.tolist()], dtype=bool)
            else:
                keep = _CMP[h["op"]](vals, h["value"])
            res = res.take(keep)
            members = [m for m, k in zip(members, keep.tolist()) if k]
        return self._order_limit(q, res, f)

    def _join_nearest(self, q, f: _Frame, j):
        """src/giql/expanders/nearest.py:336-397 — one engine call per NEAREST, joined back on the query row"""
        tgt = self._relation(q, {"kind": "table", "table": j["target"]})
        qa = j["ref_alias"] or f.order[0]
        if qa not in f.rels or f.extra[qa]:
            raise UnsupportedQuery(f"NEAREST reference alias {qa!r} is not a base table of this query")
        Q = f.rels[qa]
        if len(Q) and len(tgt):
            qi, ti, dist = _nearest(self.backend, Q, tgt, j)
        else:
            qi, ti, dist = np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
        order = np.argsort(qi, kind="stable")                              # keep the engine's per-query rank order
        qi, ti, dist = qi[order], ti[order], dist[order]
        lo, hi = np.searchsorted(qi, f.idx[qa], "left"), np.searchsorted(qi, f.idx[qa], "right")
        cnt = hi - lo
        rep = np.repeat(np.arange(len(f)), cnt)
        pos = np.repeat(lo, cnt) + np.arange(len(rep)) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        f.take(rep)
        f.add(j["alias"] or f"__nearest{len(f.order)}", tgt, ti[pos], {"distance": dist[pos]})

    def _filter(self, f: _Frame, w):
        be = self.backend
        if w["kind"] == "literal":
            # src/giql/expanders/intersects.py:85-133 (literal range), :243-270 (ANY / ALL)
            alias = w["alias"] or f.order[0]
            if alias not in f.rels:
                raise ValueError(f"unknown alias {alias!r}")
            rel = f.rels[alias]
            ranges = [parse_range(r) for r in w["ranges"]]
            lit = {"chrom": _column([r[0] for r in ranges]), "start": _column([r[1] for r in ranges]), "end": _column([r[2] for r in ranges])}
            if not len(rel):
                keep = np.zeros(0, bool)
            elif w["op"] == "intersects":
                if w["quant"] == "ALL":
                    keep = _counts(be, rel.genomic(), lit, rel.config, None) == len(ranges)
                else:
                    keep = _exists(be, rel.genomic(), lit, rel.config, None)
            else:
                # the column join against the table of the literals (intersects.py:155-166 on a literal operand); several
                # ranges: per-range predicates OR-ed for ANY, AND-ed for ALL (:243-270) = at least one / every literal row hit
                hit, _ = _pairs(be, rel.genomic(), lit, w["op"], rel.config, None)
                per_row = np.bincount(np.asarray(hit, dtype=np.int64), minlength=len(rel))
                keep = per_row == len(ranges) if w["quant"] == "ALL" else per_row > 0
            if w.get("negated"):
                keep = ~keep
            ix = f.idx[alias]
            f.take((ix >= 0) & keep[np.maximum(ix, 0)] if len(keep) else np.zeros(len(ix), bool))
        elif w["kind"] == "compare_expr":                              # arithmetic on either side; a NULL operand drops the row
            op = {"<>": "!=", "==": "="}.get(w["op"], w["op"])
            if op not in _CMP:
                raise UnsupportedQuery(f"comparison {op}")
            a, b = _eval_expr(f, w["left"]), _eval_expr(f, w["right"])
            if a.dtype == object or b.dtype == object:
                py = {"<": lambda x, y: x < y, "<=": lambda x, y: x <= y, ">": lambda x, y: x > y, ">=": lambda x, y: x >= y,
                      "=": lambda x, y: x == y, "!=": lambda x, y: x != y}[op]
                try:
                    keep = np.fromiter((x is not None and y is not None and py(x, y) for x, y in zip(a.tolist(), b.tolist())), dtype=bool, count=len(a))
                except TypeError as exc:
                    raise ValueError(f"cannot compare the operands of {op!r}: {exc}") from None
            else:
                keep = _CMP[op](a, b)
            f.take(keep)
        elif w["kind"] == "in_list":                                   # col [NOT] IN (literals); a NULL operand drops the row either way
            col = f.column(f.resolve(*w["col"]), w["col"][1])
            values = set(w["values"])
            hit = np.fromiter((x in values for x in col.tolist()), dtype=bool, count=len(col))
            known = np.fromiter((x is not None for x in col.tolist()), dtype=bool, count=len(col))
            f.take(known & (hit != w["negated"]))
        elif w["kind"] == "between":                                   # col BETWEEN lo AND hi == col >= lo AND col <= hi
            self._filter(f, {"kind": "compare_literal", "op": ">=", "col": w["col"], "value": w["lo"]})
            self._filter(f, {"kind": "compare_literal", "op": "<=", "col": w["col"], "value": w["hi"]})
        elif w["kind"] == "distance_cmp":
            for a in (w["a"], w["b"]):
                if a not in f.rels:
                    raise ValueError(f"unknown alias {a!r}")
            d = _distance(be, f.rels[w["a"]], f.rels[w["b"]], f.idx[w["a"]], f.idx[w["b"]], w["stranded"], w["signed"])
            if w["op"] not in _CMP:
                raise UnsupportedQuery(f"DISTANCE comparison {w['op']}")
            if d.dtype == object:
                ok = np.array([x is not None for x in d], dtype=bool)
                keep = np.zeros(len(d), bool)
                keep[ok] = _CMP[w["op"]](d[ok].astype(np.int64), w["value"])
            else:
                keep = _CMP[w["op"]](d, w["value"])
            f.take(keep)
        elif w["kind"] == "is_null":
            alias = f.resolve(*w["col"])
            col = f.column(alias, w["col"][1])
            null = np.array([x is None for x in col], dtype=bool) if col.dtype == object else np.zeros(len(col), bool)
            f.take(null != w["negated"])
        elif w["kind"] in ("compare", "compare_literal"):
            # a plain column comparison over the joined rows, SQL three-valued: NULL operands drop the row
            op = {"<>": "!=", "==": "="}.get(w["op"], w["op"])
            if op not in _CMP:
                raise UnsupportedQuery(f"comparison {op}")
            lc = w["left"] if w["kind"] == "compare" else w["col"]
            a = f.column(f.resolve(*lc), lc[1])
            b = f.column(f.resolve(*w["right"]), w["right"][1]) if w["kind"] == "compare" else np.full(len(a), w["value"], dtype=object if isinstance(w["value"], str) else np.float64 if isinstance(w["value"], float) else np.int64)
            if a.dtype == object or b.dtype == object:
                fn = {"<": lambda x, y: x < y, "<=": lambda x, y: x <= y, ">": lambda x, y: x > y, ">=": lambda x, y: x >= y,
                      "=": lambda x, y: x == y, "!=": lambda x, y: x != y}[op]
                try:
                    keep = np.fromiter((x is not None and y is not None and fn(x, y) for x, y in zip(a.tolist(), b.tolist())), dtype=bool, count=len(a))
                except TypeError as exc:
                    raise ValueError(f"cannot compare the operands of {op!r}: {exc}") from None
            else:
                keep = _CMP[op](a, b)
            f.take(keep)
        else:
            raise UnsupportedQuery("a residual spatial predicate between already joined tables is not handled")

    def _finish(self, q, f: _Frame) -> Result:
        """projection (incl. CLUSTER / MERGE / DISTANCE), DISTINCT, ORDER BY, LIMIT"""
        be = self.backend
        proj = q["proj"]
        kinds = {p["kind"] for p in proj}
        if kinds & {"merge", "cluster"} and len(f.order) != 1:
            raise UnsupportedQuery("CLUSTER / MERGE are bound over a single FROM source")
        if "merge" in kinds:
            # src/giql/expanders/merge.py:183-439: (chrom[, strand], start, end[, COUNT(*)]), default ORDER BY chrom, start
            m = next(p for p in proj if p["kind"] == "merge")
            g = f.genomic(f.order[0])
            if m["stranded"] and "strand" not in g:
                raise ValueError("MERGE(stranded := true) needs a strand column")
            # the partition keys may be named beside MERGE and in a GROUP BY — `SELECT chrom, MERGE(interval) FROM t GROUP BY chrom`:
            # the reference folds them into its own grouping and emits each once (merge.py:320-439); anything else changes the
            # grouping and belongs to the SQL engine
            cfg = f.rels[f.order[0]].config or Table("t")
            keys = {cfg.chrom_col.lower()} | ({cfg.strand_col.lower()} if m["stranded"] and cfg.strand_col else set())
            is_key = lambda p: p["kind"] == "col" and p["col"].lower() in keys and p.get("name", p["col"]).lower() == p["col"].lower()
            user_aggs = [p for p in proj if p["kind"] in ("count_col", "agg")]
            if any(p["kind"] not in ("merge", "count_star", "count_col", "agg") and not is_key(p) for p in proj) or q["distinct"] or \
                    any(c.lower() not in keys for _, c in q.get("group") or []):
                raise UnsupportedQuery("MERGE is bound for `SELECT [chrom, [strand,]] MERGE(interval...)[, <aggregates>] FROM t [GROUP BY chrom[, strand]]`")
            width = 4 if m["stranded"] else 3
            members: list[np.ndarray] = []
            if (m.get("predicate") or user_aggs) and len(f):
                # merge.py:201-239: MERGE is GROUP BY chrom[, strand], cluster id over the (predicate-gated) CLUSTER.  With a
                # predicate, or with user aggregates beside MERGE (`MERGE(interval), AVG(score)`, merge.py:320-439), the cluster ids
                # come from the engine and the small per-region aggregation is finished here
                ids = _cluster_ids(be, f, m)
                agg: dict[tuple, list] = {}
                strands = g["strand"].tolist() if m["stranded"] else [None] * len(ids)
                for i_row, (c, st, cid, s0, e0) in enumerate(zip(g["chrom"].tolist(), strands, ids.tolist(), g["start"].tolist(), g["end"].tolist())):
                    a = agg.setdefault((c, st, cid), [s0, e0, 0, []])
                    a[0], a[1], a[2] = min(a[0], s0), max(a[1], e0), a[2] + 1
                    a[3].append(i_row)
                rows = [((c, st, v[0], v[1], v[2]) if m["stranded"] else (c, v[0], v[1], v[2])) for (c, st, _), v in agg.items()]
                members = [np.asarray(v[3], dtype=np.int64) for v in agg.values()]
            else:
                rows = be.merge(g, m["distance"], m["stranded"]) if len(f) else []
            extra = [_aggregate(f, p, members).tolist() if members else [] for p in user_aggs]      # one list per aggregate, region-aligned
            rows = [tuple(r) + tuple(col[k] for col in extra) for k, r in enumerate(rows)]
            n_fixed = width + 1                                             # chrom[, strand], start, end, row count
            for h in q.get("having") or []:                                 # HAVING COUNT(*) <op> n over the merged regions
                if h["item"]["kind"] != "count_star" or h["op"] not in _CMP:
                    raise UnsupportedQuery("HAVING over MERGE is bound for COUNT(*) comparisons")
                rows = [r for r in rows if bool(_CMP[h["op"]](r[width], h["value"]))]
            out = [tuple(r[:width]) + ((r[width],) if "count_star" in kinds else ()) + tuple(r[n_fixed:]) for r in rows]
            out.sort(key=lambda r: (r[0], r[width - 2]))
            # the regions carry the table's own column names (merge.py:302-307: MIN(start) AS <start_col>, MAX(end) AS <end_col>)
            count_name = next((p.get("name", "count") for p in proj if p["kind"] == "count_star"), None)
            cols = [cfg.chrom_col] + ([cfg.strand_col] if m["stranded"] else []) + [cfg.start_col, cfg.end_col] + \
                   ([count_name] if count_name is not None else []) + \
                   [p.get("name", "count" if p["kind"] == "count_col" else p["fn"].lower()) for p in user_aggs]
            return self._order_limit(q, Result(cols, [_column([r[i] for r in out]) for i in range(len(cols))]), f)
        if q.get("group"):
            return self._group_count(q, f)
        if kinds == {"count_star"} and not q.get("having"):
            # SELECT COUNT(*) [AS n] FROM <rows the FROM clause produced>: one row, also over an empty input
            cols = [p.get("name", "count") for p in proj]
            return self._order_limit(q, Result(cols, [np.asarray([len(f)], dtype=np.int64) for _ in proj]), f)
        if kinds and kinds <= {"count_star", "count_col", "agg"}:
            return self._group_count(q, f)                                 # aggregates only, no GROUP BY: one group over all rows
        if kinds & {"count_star", "count_col", "agg"}:
            raise UnsupportedQuery("aggregates mixed with other projections without GROUP BY are finished by the SQL engine")
        if q.get("having"):
            raise UnsupportedQuery("HAVING without aggregates in the select list is finished by the SQL engine")
        cols: list[str] = []
        arrays: list[np.ndarray] = []
        sources: list = []
        for p in proj:
            if p["kind"] == "star":
                for a in ([p["alias"]] if p["alias"] else f.order):
                    if a not in f.rels:
                        raise ValueError(f"unknown alias {a!r}")
                    for c in f.schema(a):
                        cols.append(c)
                        arrays.append(f.column(a, c))
                        sources.append((a, c))
            elif p["kind"] == "col":
                a = f.resolve(p["alias"], p["col"])
                cols.append(p.get("name", p["col"]))
                arrays.append(f.column(a, p["col"]))
                sources.append((a, p["col"]))
            elif p["kind"] == "window":
                part = [f.column(f.resolve(a, c), c).tolist() for a, c in p["partition"]]
                groups: dict[tuple, list[int]] = {}
                for i, key in enumerate(zip(*part)) if part else []:
                    groups.setdefault(key, []).append(i)
                if not part:
                    groups[()] = list(range(len(f)))
                members = [np.asarray(v, dtype=np.int64) for v in groups.values()]
                per_group = _aggregate(f, p["agg"], members)
                out = np.empty(len(f), dtype=per_group.dtype)
                for g, v in enumerate(members):
                    out[v] = per_group[g]
                default = "count" if p["agg"]["kind"] != "agg" else p["agg"]["fn"].lower()
                cols.append(p.get("name", default))
                arrays.append(out)
                sources.append(None)
            elif p["kind"] == "expr":
                cols.append(p.get("name", "expr"))
                arrays.append(_eval_expr(f, p["expr"]))
                sources.append(None)
            elif p["kind"] == "const":
                v = p["value"]
                cols.append(p.get("name", str(v)))
                arrays.append(np.full(len(f), v, dtype=object if isinstance(v, str) else np.int64))
                sources.append(None)
            elif p["kind"] == "cluster":
                ids = _cluster_ids(be, f, p)                            # over the rows the FROM clause produced
                cols.append(p.get("name", "cluster"))
                arrays.append(ids)
                sources.append(None)
            elif p["kind"] == "distance":
                # src/giql/expanders/_distance.py:67-117 over the joined pairs
                for a in (p["a"], p["b"]):
                    if a not in f.rels:
                        raise ValueError(f"unknown alias {a!r}")
                ia, ib = f.idx[p["a"]], f.idx[p["b"]]
                if len(ia) and (ia.min() < 0 or ib.min() < 0):
                    raise UnsupportedQuery("DISTANCE over the NULL side of an outer join")
                cols.append(p.get("name", "distance"))
                arrays.append(_distance(be, f.rels[p["a"]], f.rels[p["b"]], ia, ib, p["stranded"], p["signed"]))
                sources.append(None)
            else:
                raise UnsupportedQuery(f"projection {p['kind']}")
        res = Result(cols, arrays, sources)
        if q["distinct"]:
            first: dict[tuple, int] = {}
            for i, r in enumerate(res.rows):
                first.setdefault(r, i)
            res = res.take(np.fromiter(first.values(), dtype=np.int64, count=len(first)))
            return self._order_limit(q, res, f)
        # ORDER BY over a column the select list does not carry (legal without DISTINCT; the reference's IEJoin path projects
        # such keys itself, src/giql/expanders/intersects_duckdb.py:1339-1400): sort on a hidden column of the joined rows
        shown = len(cols)
        for (alias, col), _ in q["order"]:
            try:
                self._order_key(res, alias, col, f)
            except ValueError as exc:
                if "ambiguous" in str(exc):
                    raise
                a = f.resolve(alias, col)
                res = Result(res.columns + [f"__order_{len(res.columns)}"], res.arrays + [f.column(a, col)], res.sources + [(a, col)])
                if alias is None:                                     # an unqualified hidden key is found by its name
                    res.columns[-1] = col
        res = self._order_limit(q, res, f)
        return Result(res.columns[:shown], res.arrays[:shown], res.sources[:shown])

    @staticmethod
    def _order_key(res: Result, alias, col, f) -> int:
        """Output column an ORDER BY key names.  A qualified key (``b.start``) matches the column projected from that
        relation; an unqualified one matches an output name (alias or column), which must then be unambiguous."""
        if alias is not None:
            a = f.resolve(alias, col) if f is not None else alias
            hits = [i for i, src in enumerate(res.sources) if src == (a, col)]
            if not hits:
                raise ValueError(f"ORDER BY column {alias}.{col} is not in the select list")
            return hits[0]                                                   # the same source projected twice holds equal values
        hits = [i for i, name in enumerate(res.columns) if name == col]
        if not hits:
            raise ValueError(f"ORDER BY column {col!r} is not in the select list")
        if len(hits) > 1 and len({res.sources[i] for i in hits}) > 1:
            raise ValueError(f"ORDER BY column {col!r} is ambiguous: {len(hits)} output columns carry that name; qualify it")
        return hits[0]

    @staticmethod
    def _order_limit(q, res: Result, f=None) -> Result:
        for (alias, col), desc in reversed(q["order"]):
            key = res.arrays[Session._order_key(res, alias, col, f)]
            if key.dtype == object:                                         # NULLs last, as DuckDB orders by default
                vals = key.tolist()
                some = sorted((i for i in range(len(vals)) if vals[i] is not None), key=lambda i: vals[i], reverse=desc)
                perm = np.asarray(some + [i for i in range(len(vals)) if vals[i] is None], dtype=np.int64)
            elif desc:
                # descending and stable without negating the key (negation wraps unsigned / bool keys and INT64_MIN):
                # sort on the negated dense rank
                rank = np.unique(key, return_inverse=True)[1].astype(np.int64).reshape(-1)
                perm = np.argsort(-rank, kind="stable")
            else:
                perm = np.argsort(key, kind="stable")
            res = res.take(perm)
        off = q.get("offset") or 0
        if q["limit"] is not None or off:
            res = res.take(slice(off, None if q["limit"] is None else off + q["limit"]))
        return res

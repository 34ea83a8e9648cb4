"""The CPU oracle pinned against the reference's own golden vectors (no GPU).

Sources (all harvested by tests/golden/harvest_reference_vectors.py from /root/reference):
  * tests/integration/datafusion/test_cross_target_oracle.py — 54 (query, data, expected) triples
  * tests/test_distance_udf.py — scalar DISTANCE known answers
  * src/giql/expanders/_distance.py:22-117 executed on SQLite over an input grid
  * tests/test_duckdb_iejoin.py:107-142 — peaks x genes expected overlaps
"""
import json

import numpy as np
import pytest

from golden_cases import GOLDEN, Unsupported, load_cases, normalise, run_case
from backends import OracleBackend
from oracle import giql_oracle as go
from oracle import sqlite_oracle as so

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[f"{i}:{c['test']}" for i, c in enumerate(CASES)])
def test_cross_target_golden(case):
    if "expected" not in case or isinstance(case["expected"], dict):
        pytest.skip("reference test has no literal expected rows")
    rows = run_case(case, OracleBackend())
    assert normalise(rows) == normalise(case["expected"]), case["source"]


def test_distance_known_answers():
    ka = json.loads((GOLDEN / "distance_known_answers.json").read_text())
    assert len(ka) >= 20
    for k in ka:
        a, b = k["a"], k["b"]
        stranded = k["stranded"] and a["has_strand"] and b["has_strand"]
        got = go.distance(a["chrom"], a["start"], a["end"], a["strand"], b["chrom"], b["start"], b["end"], b["strand"],
                          stranded=stranded, signed=k["signed"])
        assert got == k["expected"], k["source"]


def test_distance_matches_reference_case_on_sqlite_grid():
    grid = json.loads((GOLDEN / "distance_reference_case_grid.json").read_text())
    rows = grid["rows"]
    assert len(rows) > 1000
    for ac, as_, ae, ast, bc, bs, be, bst, stranded, signed, exp in rows:
        got = go.distance(ac, as_, ae, ast, bc, bs, be, bst, stranded=bool(stranded), signed=bool(signed))
        assert got == exp, (ac, as_, ae, ast, bc, bs, be, bst, stranded, signed)


def test_iejoin_peaks_genes_fixture():
    g = json.loads((GOLDEN / "iejoin_peaks_genes.json").read_text())
    def tab(rows):
        return {"chrom": [r[0] for r in rows], "start": [r[1] for r in rows], "end": [r[2] for r in rows]}
    P, G = tab(g["peaks"]), tab(g["genes"])
    pairs = go.join_pairs(P, G)
    got = sorted((P["chrom"][i], P["start"][i], P["end"][i], G["chrom"][j], G["start"][j], G["end"][j]) for i, j in pairs)
    assert got == sorted(tuple(r) for r in g["expected_overlaps"])


def _random_table(rng, n, chroms=("chr1", "chr2", "chrX"), span=400, max_len=60, zero_width=False, strands=None):
    start = rng.integers(0, span, size=n)
    length = rng.integers(0 if zero_width else 1, max_len, size=n)
    t = {"chrom": [chroms[i] for i in rng.integers(0, len(chroms), size=n)], "start": start.tolist(), "end": (start + length).tolist()}
    if strands:
        t["strand"] = [strands[i] for i in rng.integers(0, len(strands), size=n)]
    return t


@pytest.mark.parametrize("seed", range(6))
def test_oracle_agrees_with_sqlite_and_bruteforce(seed):
    """Second, independent oracle: the reference's SQL text shapes on stdlib SQLite (not DuckDB)."""
    rng = np.random.default_rng(seed)
    L, R = _random_table(rng, 60, zero_width=seed % 2 == 1), _random_table(rng, 50, zero_width=seed % 2 == 1)
    fast = [tuple(map(int, p)) for p in go.join_pairs(L, R)]
    assert fast == go.join_pairs_bruteforce(L, R)
    assert fast == so.intersects_pairs(L, R)
    assert sorted(go.semi_rows(L, R).tolist()) == so.semi_anti_rows(L, R, anti=False)
    assert sorted(go.anti_rows(L, R).tolist()) == so.semi_anti_rows(L, R, anti=True)


@pytest.mark.parametrize("seed", range(4))
def test_c_oracle_agrees_with_python_oracle(seed, c_oracle):
    rng = np.random.default_rng(100 + seed)
    L, R = _random_table(rng, 300, span=3000), _random_table(rng, 280, span=3000)
    names = sorted(set(L["chrom"]) | set(R["chrom"]))
    code = {c: i for i, c in enumerate(names)}
    lc, rc = [code[c] for c in L["chrom"]], [code[c] for c in R["chrom"]]
    got = c_oracle.intersects_pairs(lc, L["start"], L["end"], rc, R["start"], R["end"], len(names))
    exp = go.join_pairs(L, R)
    assert np.array_equal(got, exp)
    ex = c_oracle.intersects_exists(lc, L["start"], L["end"], rc, R["start"], R["end"], len(names))
    assert np.array_equal(np.nonzero(ex)[0], go.semi_rows(L, R))
    ids = c_oracle.cluster(lc, L["start"], L["end"], len(names), distance=5)
    assert ids.tolist() == go.cluster(L, distance=5)


@pytest.mark.parametrize("seed,op,distance,stranded", [(0, "=", 0, False), (1, "!=", 0, True), (2, ">=", 30, False), (3, "<", 0, False), (4, "=", 100, True)])
def test_cluster_predicate_matches_reference_sql_shape_on_sqlite(seed, op, distance, stranded):
    """`predicate := cur <op> PREV(prev)`: the oracle against the reference's LAG rewrite (cluster.py:281-310, 587-644)
    executed on SQLite.  Starts are unique within a partition — with ties the reference's result depends on the engine's
    tie order."""
    rng = np.random.default_rng(seed)
    n = 400
    chrom = [["chr1", "chr2", "chr3"][i] for i in rng.integers(0, 3, size=n)]
    start = rng.permutation(4000)[:n]
    T = {"chrom": chrom, "start": start.tolist(), "end": (start + rng.integers(1, 60, size=n)).tolist(),
         "strand": [["+", "-"][i] for i in rng.integers(0, 2, size=n)]}
    cur = [None if i % 17 == 0 else int(v) for i, v in enumerate(rng.integers(0, 4, size=n))]
    prev = [None if i % 23 == 0 else int(v) for i, v in enumerate(rng.integers(0, 4, size=n))]
    assert go.cluster(T, distance, stranded, predicate=(cur, op, prev)) == so.cluster_predicate(T, cur, op, prev, distance, stranded)
    names = [["a", "b", "c"][i] for i in rng.integers(0, 3, size=n)]
    assert go.cluster(T, distance, stranded, predicate=(names, op, names)) == so.cluster_predicate(T, names, op, names, distance, stranded)


# ---- the C restatement's NEAREST / DISJOIN / CONTAINS against the pure-Python restatement (pinned above) --------
def _rand_table(rng, n, n_chrom=3, span=2000, wmax=300, strands=False):
    chrom = rng.integers(0, n_chrom, n).astype(np.int32)
    start = rng.integers(0, span, n).astype(np.int32)
    end = (start + rng.integers(1, wmax, n)).astype(np.int32)
    T = {"chrom": [f"c{c}" for c in chrom], "start": start.tolist(), "end": end.tolist()}
    code = {"code": chrom, "start": start, "end": end}
    if strands:
        sc = rng.choice(np.array([0, 1, 2, 255], dtype=np.uint8), n, p=[0.45, 0.45, 0.05, 0.05])
        T["strand"] = [{0: "+", 1: "-", 2: ".", 255: None}[int(v)] for v in sc]
        code["strand"] = sc
    return T, code


@pytest.mark.parametrize("seed,k,maxd,stranded,signed", [(1, 1, None, False, False), (2, 3, None, False, True), (3, 5, 150, False, False),
                                                         (4, 2, None, True, False), (5, 4, None, True, True), (6, 3, 400, True, True)])
def test_c_oracle_nearest_equals_python_oracle(c_oracle, seed, k, maxd, stranded, signed):
    from oracle import giql_oracle as go

    rng = np.random.default_rng(seed)
    Q, qc = _rand_table(rng, 150, strands=stranded)
    T, tc = _rand_table(rng, 120, strands=stranded)
    want = go.nearest(Q, T, k=k, max_distance=maxd, stranded=stranded, signed=signed)
    q, t, d, v = c_oracle.nearest(qc["code"], qc["start"], qc["end"], tc["code"], tc["start"], tc["end"], 3, k=k, max_distance=maxd,
                                  q_strand=qc.get("strand"), t_strand=tc.get("strand"), stranded=stranded, signed=signed)
    got = [(int(a), int(b), int(c) if ok else None) for a, b, c, ok in zip(q, t, d, v)]
    assert got == want


@pytest.mark.parametrize("seed,self_ref", [(1, True), (2, False), (3, False)])
def test_c_oracle_disjoin_equals_python_oracle(c_oracle, seed, self_ref):
    from oracle import giql_oracle as go

    rng = np.random.default_rng(seed)
    T, tc = _rand_table(rng, 200)
    R, rc = _rand_table(rng, 150)
    want = go.disjoin(T, None if self_ref else R)
    args = () if self_ref else (rc["code"], rc["start"], rc["end"])
    row, s, e = c_oracle.disjoin(tc["code"], tc["start"], tc["end"], *args, n_codes=3)
    assert sorted(zip(row.tolist(), s.tolist(), e.tolist())) == want


@pytest.mark.parametrize("op", ["contains", "within"])
def test_c_oracle_spatial_joins_equal_python_oracle(c_oracle, op):
    from oracle import giql_oracle as go

    rng = np.random.default_rng(9)
    L, lc = _rand_table(rng, 300, wmax=600)
    R, rc = _rand_table(rng, 300, wmax=200)
    want = sorted((int(a), int(b)) for a, b in go.join_pairs(L, R, op))
    got = c_oracle.spatial_pairs(lc["code"], lc["start"], lc["end"], rc["code"], rc["start"], rc["end"], 3, op)
    assert [tuple(r) for r in got.tolist()] == want and len(want) > 20

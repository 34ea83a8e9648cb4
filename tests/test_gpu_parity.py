"""GPU parity tests: the CUDA path (Engine -> ctypes -> libgiqlb200.so) against the CPU oracle, the
reference's golden vectors and size-independent properties.  Integer work: every comparison is bit-exact.
"""
import json

import numpy as np
import pytest

from backends import EngineBackend, OracleBackend
from golden_cases import GOLDEN, Unsupported, load_cases, normalise, run_case

pytestmark = pytest.mark.gpu

CASES = load_cases()


def _rand(rng, n, chroms=("chr1", "chr2", "chrX"), span=400, max_len=60, min_len=1, strands=None):
    start = rng.integers(0, span, size=n)
    length = rng.integers(min_len, max_len, size=n)
    t = {"chrom": [chroms[i] for i in rng.integers(0, len(chroms), size=n)], "start": start.tolist(), "end": (start + length).tolist()}
    if strands:
        t["strand"] = [strands[i] for i in rng.integers(0, len(strands), size=n)]
    return t


# ---- the reference's golden vectors through the CUDA path --------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=[f"{i}:{c['test']}" for i, c in enumerate(CASES)])
def test_cross_target_golden_on_gpu(case, engine):
    if "expected" not in case or isinstance(case["expected"], dict):
        pytest.skip("reference test has no literal expected rows")
    try:
        rows = run_case(case, EngineBackend(engine))
    except Unsupported as exc:
        pytest.skip(f"literal-range {exc} filter is not on the engine path yet")
    assert normalise(rows) == normalise(case["expected"]), case["source"]


def test_iejoin_peaks_genes_fixture_on_gpu(engine):
    g = json.loads((GOLDEN / "iejoin_peaks_genes.json").read_text())
    tab = lambda rows: {"chrom": [r[0] for r in rows], "start": [r[1] for r in rows], "end": [r[2] for r in rows]}
    P, G = tab(g["peaks"]), tab(g["genes"])
    pairs = EngineBackend(engine).pairs(P, G)
    got = sorted((P["chrom"][i], P["start"][i], P["end"][i], G["chrom"][j], G["start"][j], G["end"][j]) for i, j in pairs)
    assert got == sorted(tuple(r) for r in g["expected_overlaps"])


def test_distance_known_answers_on_gpu(engine):
    ka = json.loads((GOLDEN / "distance_known_answers.json").read_text())
    be = EngineBackend(engine)
    for k in ka:
        a, b = k["a"], k["b"]
        A = {"chrom": [a["chrom"]], "start": [a["start"]], "end": [a["end"]]}
        B = {"chrom": [b["chrom"]], "start": [b["start"]], "end": [b["end"]]}
        if a["has_strand"]:
            A["strand"] = [a["strand"]]
        if b["has_strand"]:
            B["strand"] = [b["strand"]]
        got = be.distance(A, B, stranded=k["stranded"], signed=k["signed"])
        assert got == [k["expected"]], k["source"]


def test_distance_reference_grid_on_gpu(engine):
    grid = json.loads((GOLDEN / "distance_reference_case_grid.json").read_text())["rows"]
    be = EngineBackend(engine)
    for stranded in (False, True):
        for signed in (False, True):
            rows = [r for r in grid if bool(r[8]) == stranded and bool(r[9]) == signed]
            A = {"chrom": [r[0] for r in rows], "start": [r[1] for r in rows], "end": [r[2] for r in rows], "strand": [r[3] for r in rows]}
            B = {"chrom": [r[4] for r in rows], "start": [r[5] for r in rows], "end": [r[6] for r in rows], "strand": [r[7] for r in rows]}
            got = be.distance(A, B, stranded=stranded, signed=signed)
            assert got == [r[10] for r in rows]


# ---- INTERSECTS family against the oracle ---------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_join_random_vs_oracle(seed, engine):
    rng = np.random.default_rng(seed)
    # reference hypothesis domain: 3 chroms, start 0-200, len 1-50 (tests/test_duckdb_iejoin.py:5926-5945), scaled up
    L, R = _rand(rng, 300 + 50 * seed), _rand(rng, 200 + 70 * seed)
    ob, eb = OracleBackend(), EngineBackend(engine)
    assert eb.pairs(L, R) == ob.pairs(L, R)
    assert eb.exists(L, R) == ob.exists(L, R)
    assert eb.counts(L, R) == ob.counts(L, R)


@pytest.mark.parametrize("seed", range(4))
def test_join_with_degenerate_rows_vs_oracle(seed, engine):
    """zero-width rows on either side take the general (literal-predicate) path"""
    rng = np.random.default_rng(50 + seed)
    L, R = _rand(rng, 250, min_len=0, max_len=8, span=120), _rand(rng, 260, min_len=0, max_len=8, span=120)
    ob, eb = OracleBackend(), EngineBackend(engine)
    assert eb.pairs(L, R) == ob.pairs(L, R)
    assert eb.exists(L, R) == ob.exists(L, R)
    assert eb.counts(L, R) == ob.counts(L, R)


def test_join_edge_cases(engine):
    eb, ob = EngineBackend(engine), OracleBackend()
    empty = {"chrom": [], "start": [], "end": []}
    one = {"chrom": ["chr1"], "start": [10], "end": [20]}
    assert eb.pairs(empty, one) == [] and eb.pairs(one, empty) == [] and eb.pairs(empty, empty) == []
    assert eb.exists(one, empty) == [False] and eb.counts(one, empty) == [0]
    # touching is not overlap; 1-bp overlap is (test_cross_target_oracle.py:1488-1525)
    L = {"chrom": ["chr1", "chr1"], "start": [100, 100], "end": [200, 201]}
    R = {"chrom": ["chr1"], "start": [200], "end": [300]}
    assert eb.pairs(L, R) == [(1, 0)]
    # duplicates keep multiplicity (tests/test_duckdb_iejoin.py:4514-4555)
    L = {"chrom": ["chr1"] * 3, "start": [5, 5, 5], "end": [9, 9, 9]}
    R = {"chrom": ["chr1"] * 2, "start": [7, 7], "end": [8, 8]}
    assert eb.pairs(L, R) == [(i, j) for i in range(3) for j in range(2)]
    # never crosses chromosomes, even when coordinates coincide; quote in a name (test_duckdb_iejoin.py:3780-3810)
    L = {"chrom": ["chr1", "chr'2", "chr3"], "start": [0, 0, 0], "end": [1000, 1000, 1000]}
    R = {"chrom": ["chr'2", "chr4", "chr1"], "start": [10, 10, 999], "end": [20, 20, 5000]}
    assert eb.pairs(L, R) == ob.pairs(L, R) == [(0, 2), (1, 0)]
    # disjoint chromosome sets -> empty (tests/test_duckdb_iejoin.py:3715-3744); ANTI keeps left-only chroms (:5873-5919)
    L = {"chrom": ["chrA", "chrB"], "start": [1, 1], "end": [5, 5]}
    R = {"chrom": ["chrC"], "start": [1], "end": [5]}
    assert eb.pairs(L, R) == []
    assert engine.anti(eb._table(L, None), eb._table(R, None)).cpu().tolist() == [0, 1]
    assert engine.semi(eb._table(L, None), eb._table(R, None)).cpu().tolist() == []


def test_join_coordinate_encodings(engine):
    """4-encoding matrix: the same canonical intervals stored under each convention give the same pairs
    (tests/test_duckdb_iejoin.py:5174-5233; encode() = inverse of canonical.py, coordinate_space/encodings.py:15-66)."""
    from giql_b200.table import Table, canonical_offsets

    rng = np.random.default_rng(3)
    L, R = _rand(rng, 120), _rand(rng, 130)
    eb, ob = EngineBackend(engine), OracleBackend()
    expect = ob.pairs(L, R)
    for cs in ("0based", "1based"):
        for it in ("half_open", "closed"):
            cfg = Table("t", coordinate_system=cs, interval_type=it)
            sadj, eadj = canonical_offsets(cfg)
            enc = lambda T: {"chrom": T["chrom"], "start": [s - sadj for s in T["start"]], "end": [e - eadj for e in T["end"]]}
            assert eb.pairs(enc(L), R, cfg, None) == expect
            assert eb.pairs(L, enc(R), None, cfg) == expect
            assert eb.exists(enc(L), enc(R), cfg, cfg) == ob.exists(L, R)


def test_join_skewed_output(engine):
    """one giant interval overlapping everything + a desert of non-overlapping rows (load-balancing paths)"""
    n = 5000
    L = {"chrom": ["chr1"] * (n + 1), "start": [0] + list(range(0, 10 * n, 10)), "end": [10 * n + 5] + [s + 3 for s in range(0, 10 * n, 10)]}
    R = {"chrom": ["chr1"] * n + ["chr2"] * n, "start": list(range(1, 10 * n, 10)) * 2, "end": [s + 1 for s in range(1, 10 * n, 10)] * 2}
    eb = EngineBackend(engine)
    got = eb.pairs(L, R)
    exp = sorted([(0, j) for j in range(n)] + [(i + 1, i) for i in range(n)])
    assert got == exp


def test_join_c1_size_vs_c_oracle(engine, c_oracle):
    """BASELINE configs[0]: 100k x 100k synthetic hg38 tables; sorted packed pairs must be identical."""
    from giql_b200 import synth

    A, B = synth.peaks(100_000, 20260101), synth.peaks(100_000, 20260102)
    ta = engine.table_from_codes(A["code"], A["start"], A["end"], A["names"])
    tb = engine.table_from_codes(B["code"], B["start"], B["end"], ta.dictionary)
    got = engine.intersects(ta, tb).sorted_packed()
    exp = c_oracle.intersects_pairs(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"], len(A["names"]))
    exp = (exp[:, 0] << 32) | exp[:, 1]
    assert len(got) == len(exp) and np.array_equal(got, exp)
    ex = engine.exists(ta, tb).cpu().numpy().astype(bool)
    assert np.array_equal(ex, c_oracle.intersects_exists(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"], len(A["names"])))
    cnt = engine.count_overlaps(ta, tb).cpu().numpy()
    assert np.array_equal(cnt, c_oracle.intersects_counts(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"], len(A["names"])))


def test_join_string_chroms_match_precoded(engine):
    """device dictionary encoding: string ingest and pre-encoded ingest give identical pairs"""
    from giql_b200 import synth

    A, B = synth.with_strings(synth.peaks(20_000, 11)), synth.with_strings(synth.genes(5_000, 12))
    sa = engine.table({"chrom": A["chrom"], "start": A["start"], "end": A["end"]})
    sb = engine.table({"chrom": B["chrom"], "start": B["start"], "end": B["end"]})
    ca = engine.table_from_codes(A["code"], A["start"], A["end"], A["names"])
    cb = engine.table_from_codes(B["code"], B["start"], B["end"], ca.dictionary)
    assert np.array_equal(engine.intersects(sa, sb).sorted_packed(), engine.intersects(ca, cb).sorted_packed())
    present = sorted(set(A["chrom"]) | set(B["chrom"]), key=lambda s: s.encode())
    assert sa.dictionary.names == present                      # bytewise order: chr1 < chr10 < chr2
    names = np.array(sa.dictionary.names, dtype=object)
    assert (names[sa.codes.cpu().numpy()] == A["chrom"]).all()


def test_join_full_size_properties(engine):
    """10M x 10M (BASELINE configs[1]): size-independent properties — every emitted pair satisfies the literal
    predicate, the total equals sum(count_overlaps), per-left multiplicities equal count_overlaps, and the
    mirrored join (R x L) is the transposed multiset."""
    import torch

    from giql_b200 import synth

    A, B = synth.peaks(10_000_000, 20260103), synth.peaks(10_000_000, 20260104)
    ta = engine.table_from_codes(A["code"], A["start"], A["end"], A["names"])
    tb = engine.table_from_codes(B["code"], B["start"], B["end"], ta.dictionary)
    res = engine.intersects(ta, tb)
    l, r = res.left_rows.long(), res.right_rows.long()
    assert len(res) > 10_000_000
    ok = (ta.codes[l] == tb.codes[r]) & (ta.start[l] < tb.end[r]) & (ta.end[l] > tb.start[r])
    assert bool(ok.all())
    cnt = engine.count_overlaps(ta, tb)
    assert int(cnt.sum()) == len(res)
    assert torch.equal(torch.bincount(l, minlength=ta.n), cnt)
    packed = torch.sort((l << 32) | r).values
    assert bool((packed[1:] != packed[:-1]).all())            # no pair emitted twice
    mirror = engine.intersects(tb, ta)
    mp = torch.sort((mirror.right_rows.long() << 32) | mirror.left_rows.long()).values
    assert torch.equal(packed, mp)
    ex = engine.exists(ta, tb)
    assert torch.equal(ex.bool(), cnt > 0)

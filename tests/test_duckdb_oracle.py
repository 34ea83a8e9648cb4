"""Third oracle at C1 when the box has DuckDB (it does not in this image: the tests skip): the reference's transpiled SQL on
DuckDB — or, without giql, the hand-restated naive predicate — against the C restatement the GPU suite is checked with."""
import numpy as np
import pytest

duckdb = pytest.importorskip("duckdb", reason="duckdb is not installed in this image")


@pytest.mark.parametrize("plan", ["iejoin", "naive"])
def test_c_oracle_equals_duckdb_at_c1(c_oracle, plan):
    from giql_b200 import synth
    from oracle import duckdb_oracle as do

    A, B = synth.peaks(100_000, 20260101), synth.peaks(100_000, 20260102)
    n, _, packed = do.intersects_join_timed(A, B, threads=4, plan=plan, repeats=1)
    exp = c_oracle.intersects_pairs(A["code"], A["start"], A["end"], B["code"], B["start"], B["end"], len(A["names"]))
    assert n == len(exp) and np.array_equal(packed, (exp[:, 0] << 32) | exp[:, 1])

"""Host-side logic that needs no GPU: Table contract, canonical offsets, ingest conversions."""
import numpy as np
import pytest

from giql_b200.table import Table, Tables, canonical_offsets


def test_table_defaults_match_reference():
    # reference: src/giql/table.py:70-77, src/giql/constants.py:7-11
    t = Table("peaks")
    assert (t.genomic_col, t.chrom_col, t.start_col, t.end_col, t.strand_col) == ("interval", "chrom", "start", "end", "strand")
    assert (t.coordinate_system, t.interval_type) == ("0based", "half_open")


def test_table_validation_messages_match_reference():
    # reference: src/giql/table.py:79-93
    with pytest.raises(ValueError, match="coordinate_system must be '0based' or '1based', got 'zero'"):
        Table("t", coordinate_system="zero")
    with pytest.raises(ValueError, match="interval_type must be 'half_open' or 'closed', got 'open'"):
        Table("t", interval_type="open")


def test_tables_container():
    ts = Tables()
    t = Table("a")
    ts.register("a", t)
    assert "a" in ts and ts.get("a") is t and ts.get("b") is None and list(ts) == [t]


@pytest.mark.parametrize("cs,it,exp", [
    ("0based", "half_open", (0, 0)), ("0based", "closed", (0, 1)),
    ("1based", "half_open", (-1, -1)), ("1based", "closed", (-1, 0)),
])
def test_canonical_offsets(cs, it, exp):
    # reference: src/giql/canonical.py:16-52; inverse pinned by tests/integration/coordinate_space/encodings.py:15-66
    assert canonical_offsets(Table("t", coordinate_system=cs, interval_type=it)) == exp
    assert canonical_offsets(None) == (0, 0)


def test_utf8_ingest_layouts():
    pa = pytest.importorskip("pyarrow")
    from giql_b200.engine import _as_utf8

    offs, data = _as_utf8(["chr1", "chr'2", "", "chrX"], "chrom")
    assert offs.tolist() == [0, 4, 9, 9, 13] and bytes(data[:13]) == b"chr1chr'2chrX"
    sliced = pa.array(["a", "bb", "ccc", "dddd"]).slice(1, 2)
    offs, data = _as_utf8(sliced, "chrom")
    assert [bytes(data[offs[i]:offs[i + 1]]) for i in range(2)] == [b"bb", b"ccc"]
    offs, data = _as_utf8(np.array(["x", "yy"]), "chrom")
    assert offs.tolist() == [0, 1, 3]
    offs, _ = _as_utf8([], "chrom")
    assert offs.tolist() == [0]
    with pytest.raises(ValueError, match="NULL chromosome"):
        _as_utf8(["chr1", None], "chrom")


def test_int_ingest_rejects_nulls_and_overflow():
    from giql_b200.engine import _as_i32

    assert _as_i32([1, 2, 3], "start").dtype == np.int32
    assert _as_i32(np.array([1, 2], dtype=np.int64), "start").tolist() == [1, 2]
    with pytest.raises(ValueError, match="int32"):
        _as_i32(np.array([2**31], dtype=np.int64), "start")
    with pytest.raises(ValueError, match="NULL"):
        _as_i32(np.array([1.0, np.nan]), "start")
    with pytest.raises(ValueError, match="integer"):
        _as_i32(np.array(["a"]), "start")


def test_synth_generators_are_valid_bed():
    from giql_b200 import synth

    for gen in (synth.peaks, synth.genes, synth.reads):
        t = gen(5000, 7)
        assert (t["start"] >= 0).all() and (t["end"] > t["start"]).all()
        lens = np.array([dict(synth.HG38)[n] for n in np.array(t["names"], dtype=object)[t["code"]]])
        assert (t["end"] <= lens).all()
        assert t["names"] == sorted(t["names"], key=lambda s: s.encode())
    a, b = synth.peaks(1000, 1), synth.peaks(1000, 1)
    assert np.array_equal(a["start"], b["start"])   # seeded


def test_bind_host_to_gpu_is_safe_without_a_gpu():
    """the affinity helper never widens the allowed core set and degrades to a no-op when NVML / the GPU is absent"""
    import os

    import giql_b200

    before = os.sched_getaffinity(0)
    prev, new = giql_b200.bind_host_to_gpu(0)
    try:
        assert prev == before
        assert new is None or (new and new <= before)
    finally:
        os.sched_setaffinity(0, before)

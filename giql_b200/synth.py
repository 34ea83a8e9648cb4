"""Synthetic hg38-shaped BED6 tables: 24 primary chromosomes at real lengths,
log-normal interval widths, clustered peak-like starts, shuffled row order.  numpy only (host side);
used by bench.py and the tests.  Column set matches the reference's BED6 loader
(abdenlab/giql tests/integration/bedtools/utils/duckdb_loader.py:17-27).
"""
from __future__ import annotations

import numpy as np

HG38 = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
    ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
    ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
    ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415),
]
NAMES = [c for c, _ in HG38]
LENGTHS = np.array([n for _, n in HG38], dtype=np.int64)
CUM = np.concatenate([[0], np.cumsum(LENGTHS)])
GENOME = int(CUM[-1])

# dictionary order = bytewise string order (chr1 < chr10 < ... < chr2 < ...), as the engine's codes are
SORTED_NAMES = sorted(NAMES, key=lambda s: s.encode())
_CODE_OF_GENOME_ORDER = np.array([SORTED_NAMES.index(c) for c in NAMES], dtype=np.int32)


def _place(rng, n, width, hotspot_frac, n_hotspots, sigma):
    """chromosome (genome order index) and start for n intervals of the given widths"""
    n_hot = int(round(n * hotspot_frac))
    pos = np.empty(n, dtype=np.int64)
    centres = rng.integers(0, GENOME, size=max(n_hotspots, 1))
    pick = rng.integers(0, len(centres), size=n_hot)
    pos[:n_hot] = centres[pick] + np.rint(rng.normal(0.0, sigma, size=n_hot)).astype(np.int64)
    pos[n_hot:] = rng.integers(0, GENOME, size=n - n_hot)
    # the chromosome comes from the hotspot centre (or the uniform draw itself)
    anchor = pos.copy()
    anchor[:n_hot] = centres[pick]
    chrom = np.searchsorted(CUM, anchor, side="right") - 1
    chrom = np.clip(chrom, 0, len(LENGTHS) - 1)
    start = pos - CUM[chrom]
    start = np.clip(start, 0, np.maximum(LENGTHS[chrom] - width, 0))
    return chrom.astype(np.int32), start.astype(np.int64)


def _finish(rng, chrom, start, width, shuffle=True):
    n = len(start)
    end = start + width
    strand = rng.integers(0, 2, size=n).astype(np.uint8)       # 0 '+', 1 '-'
    score = rng.integers(0, 1001, size=n).astype(np.int32)
    if shuffle:
        perm = rng.permutation(n)
        chrom, start, end, strand, score = chrom[perm], start[perm], end[perm], strand[perm], score[perm]
    return {
        "code": _CODE_OF_GENOME_ORDER[chrom],                   # int32 code into SORTED_NAMES
        "start": start.astype(np.int32),
        "end": end.astype(np.int32),
        "strand_code": strand,
        "score": score,
        "names": SORTED_NAMES,
    }


def peaks(n: int, seed: int):
    """width ~ LogNormal(ln 500, 0.75) clipped [1, 100000]; 80 % hotspot starts (n/20 centres, sigma 2000)."""
    rng = np.random.default_rng(seed)
    width = np.clip(np.rint(rng.lognormal(np.log(500.0), 0.75, size=n)), 1, 100_000).astype(np.int64)
    chrom, start = _place(rng, n, width, 0.8, max(n // 20, 1), 2000.0)
    return _finish(rng, chrom, start, width)


def genes(n: int, seed: int):
    """width ~ LogNormal(ln 20000, 1.2) clipped [200, 2500000]; uniform starts."""
    rng = np.random.default_rng(seed)
    width = np.clip(np.rint(rng.lognormal(np.log(20000.0), 1.2, size=n)), 200, 2_500_000).astype(np.int64)
    chrom, start = _place(rng, n, width, 0.0, 1, 1.0)
    return _finish(rng, chrom, start, width)


def reads(n: int, seed: int):
    """width 150 * LogNormal(0, 0.1); 70 % hotspot / 30 % uniform; shuffled."""
    rng = np.random.default_rng(seed)
    width = np.clip(np.rint(150.0 * rng.lognormal(0.0, 0.1, size=n)), 1, 10_000).astype(np.int64)
    chrom, start = _place(rng, n, width, 0.7, max(n // 20, 1), 2000.0)
    return _finish(rng, chrom, start, width)


def with_strings(t: dict) -> dict:
    """Adds the VARCHAR columns (chrom, strand, name) of a BED6 table to a generated table."""
    names = np.array(t["names"], dtype=object)
    out = dict(t)
    out["chrom"] = names[t["code"]]
    out["strand"] = np.array(["+", "-"], dtype=object)[t["strand_code"]]
    return out

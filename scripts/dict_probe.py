"""Dictionary-encode time across chromosome naming schemes (the per-CTA name cache must not depend on how 24 particular
names happen to hash): 2 x 10 M rows per scheme, names drawn with hg38-like weights, `Engine.encode` timed with CUDA events.

    python scripts/dict_probe.py            # needs a B200; GIQLB_LIB selects an A/B build
"""
import sys
from pathlib import Path

import numpy as np
import pyarrow as pa
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import giql_b200  # noqa: E402

N = 10_000_000
SCHEMES = {
    "ucsc chr1..chrY (24)": [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"],
    "ensembl 1..MT (25)": [str(i) for i in range(1, 23)] + ["X", "Y", "MT"],
    "refseq NC_0000NN.NN (24)": [f"NC_{i:06d}.{10 + i % 3}" for i in range(1, 25)],
    "mouse chr1..chrY (21)": [f"chr{i}" for i in range(1, 20)] + ["chrX", "chrY"],
    "ucsc + 170 alt contigs (194)": [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"] + [f"chrUn_KI27{i:04d}v1" for i in range(170)],
    "scaffolds (3000)": [f"scaffold_{i}" for i in range(3000)],
}


def column(names, seed):
    rng = np.random.default_rng(seed)
    w = np.linspace(3.0, 1.0, len(names)) ** 2                        # long chromosomes first, like a genome
    idx = rng.choice(len(names), size=N, p=w / w.sum())
    arr = pa.array(names).take(pa.array(idx))
    return arr


def main():
    eng = giql_b200.Engine(0)
    z = np.zeros(N, np.int32)
    for label, names in SCHEMES.items():
        ta = eng.table({"chrom": column(names, 1), "start": z, "end": z + 1})
        tb = eng.table({"chrom": column(names, 2), "start": z, "end": z + 1})
        times = []
        for _ in range(6):
            ta.dictionary = tb.dictionary = None
            ta.codes = torch.empty_like(ta.codes) if ta.codes is not None else None
            tb.codes = torch.empty_like(tb.codes) if tb.codes is not None else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            d = eng.encode(ta, tb)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        assert len(d.names) == len(names)
        print(f"{label:32s} encode 2 x {N // 1_000_000} M rows: median {np.median(times[1:]):.3f} ms  (min {min(times[1:]):.3f})")


if __name__ == "__main__":
    main()

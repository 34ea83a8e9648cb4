"""The synthetic hg38-shaped tables of SURVEY.md §8(d), generated on the device with torch (same distributions as
``synth.py``, a different random stream): 200 M reads appear in ~0.1 s instead of minutes of host numpy, so bench.py can
run the large BASELINE configs inside its time budget.  Deterministic for a given (seed, device type): every rank of a
multi-GPU run generates the identical table.  Returns int32 device tensors ``code`` (into ``synth.SORTED_NAMES``),
``start``, ``end`` and uint8 ``strand_code``; ``names`` as in ``synth``.
"""
from __future__ import annotations

import math

import torch

from . import synth


def _place(gen, n, width, hotspot_frac, n_hotspots, sigma, device):
    lengths = torch.tensor(synth.LENGTHS.tolist(), dtype=torch.int64, device=device)
    cum = torch.tensor(synth.CUM.tolist(), dtype=torch.int64, device=device)
    n_hot = int(round(n * hotspot_frac))
    pos = torch.empty(n, dtype=torch.int64, device=device)
    anchor = torch.empty(n, dtype=torch.int64, device=device)
    if n_hot:
        centres = torch.randint(0, synth.GENOME, (max(n_hotspots, 1),), generator=gen, device=device, dtype=torch.int64)
        pick = torch.randint(0, centres.numel(), (n_hot,), generator=gen, device=device, dtype=torch.int64)
        jitter = torch.round(torch.randn(n_hot, generator=gen, device=device, dtype=torch.float32) * sigma).to(torch.int64)
        anchor[:n_hot] = centres[pick]
        pos[:n_hot] = anchor[:n_hot] + jitter
    if n - n_hot:
        pos[n_hot:] = torch.randint(0, synth.GENOME, (n - n_hot,), generator=gen, device=device, dtype=torch.int64)
        anchor[n_hot:] = pos[n_hot:]
    chrom = (torch.searchsorted(cum, anchor, right=True) - 1).clamp_(0, lengths.numel() - 1)
    start = pos - cum[chrom]
    start = torch.minimum(start.clamp_(min=0), (lengths[chrom] - width).clamp_(min=0))
    return chrom, start


def _finish(gen, chrom, start, width, device):
    n = start.numel()
    code_of = torch.tensor(synth._CODE_OF_GENOME_ORDER.tolist(), dtype=torch.int32, device=device)
    perm = torch.randperm(n, generator=gen, device=device)            # shuffled row order: the inputs are unsorted
    chrom, start, width = chrom[perm], start[perm], width[perm]
    strand = torch.randint(0, 2, (n,), generator=gen, device=device, dtype=torch.int64).to(torch.uint8)
    return {"code": code_of[chrom], "start": start.to(torch.int32), "end": (start + width).to(torch.int32), "strand_code": strand,
            "names": synth.SORTED_NAMES}


def _lognormal(gen, n, mu, sigma, device):
    return torch.exp(torch.randn(n, generator=gen, device=device, dtype=torch.float32) * sigma + mu)


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def peaks(n: int, seed: int, device):
    """width ~ LogNormal(ln 500, 0.75) clipped [1, 100000]; 80 % hotspot starts (n/20 centres, sigma 2000)."""
    g = _gen(seed, device)
    width = torch.round(_lognormal(g, n, math.log(500.0), 0.75, device)).clamp_(1, 100_000).to(torch.int64)
    chrom, start = _place(g, n, width, 0.8, max(n // 20, 1), 2000.0, device)
    return _finish(g, chrom, start, width, device)


def genes(n: int, seed: int, device):
    """width ~ LogNormal(ln 20000, 1.2) clipped [200, 2500000]; uniform starts."""
    g = _gen(seed, device)
    width = torch.round(_lognormal(g, n, math.log(20000.0), 1.2, device)).clamp_(200, 2_500_000).to(torch.int64)
    chrom, start = _place(g, n, width, 0.0, 1, 1.0, device)
    return _finish(g, chrom, start, width, device)


def reads(n: int, seed: int, device):
    """width 150 * LogNormal(0, 0.1); 70 % hotspot / 30 % uniform; shuffled."""
    g = _gen(seed, device)
    width = torch.round(150.0 * _lognormal(g, n, 0.0, 0.1, device)).clamp_(1, 10_000).to(torch.int64)
    chrom, start = _place(g, n, width, 0.7, max(n // 20, 1), 2000.0, device)
    return _finish(g, chrom, start, width, device)

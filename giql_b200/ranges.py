"""Genomic range literals — the host-side mirror of the reference's ``RangeParser``.

``'chr1:1000-2000'`` / ``'chr1:[a,b)'`` / ``'chr1:[a,b]'`` / ``'chr1:1500'`` -> ``(chrom, start, end)`` in canonical
0-based half-open coordinates, with the reference's error messages (src/giql/range_parser.py:75-188; a closed upper
bound becomes ``end + 1`` at :44-62; a bare position is the one-base interval at :104-118).
"""
from __future__ import annotations

import re

_POINT = re.compile(r"^([\w.]+):(\d+)$")
_BRACKET = re.compile(r"^([\w.]+):\[(\d+),(\d+)([\)\]])(?::([+-]))?$")   # no blank after the comma: range_parser.py:84-86
_DASH = re.compile(r"^([\w.]+):(\d+)-(\d+)(?::([+-]))?$")


def parse_range(text: str) -> tuple[str, int, int]:
    text = text.strip().strip("'\"")
    m = _POINT.match(text)
    if m:
        return m.group(1), int(m.group(2)), int(m.group(2)) + 1
    m = _BRACKET.match(text)
    closed = False
    if m:
        closed = m.group(4) == "]"
    else:
        m = _DASH.match(text)
    if not m:
        raise ValueError(f"Invalid genomic range format: {text}")
    start, end = int(m.group(2)), int(m.group(3))
    if start >= end:
        raise ValueError(f"Start must be less than end: {start} >= {end}")
    return m.group(1), start, end + (1 if closed else 0)

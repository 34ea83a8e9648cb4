"""Genomic table configuration — the host-side mirror of the reference's ``giql.Table``.

Same field names, defaults, validation messages and coordinate conventions as abdenlab/giql
``src/giql/table.py:16-93`` (``Table``) and ``:96-136`` (``Tables``), so a ``giql.Table`` can be
passed wherever a ``giql_b200.Table`` is expected (duck-typed on the attributes).  The reference applies
the coordinate convention by wrapping SQL column fragments (``src/giql/canonical.py:16-52``); the engine
applies the same convention as two integer offsets at ingest (``canonical_offsets``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, Optional

DEFAULT_CHROM_COL = "chrom"
DEFAULT_START_COL = "start"
DEFAULT_END_COL = "end"
DEFAULT_STRAND_COL = "strand"
DEFAULT_GENOMIC_COL = "interval"


@dataclass
class Table:
    name: str
    genomic_col: str = DEFAULT_GENOMIC_COL
    chrom_col: str = DEFAULT_CHROM_COL
    start_col: str = DEFAULT_START_COL
    end_col: str = DEFAULT_END_COL
    strand_col: Optional[str] = DEFAULT_STRAND_COL
    coordinate_system: str = "0based"
    interval_type: str = "half_open"

    def __post_init__(self) -> None:
        if self.coordinate_system not in ("0based", "1based"):
            raise ValueError(
                f"coordinate_system must be '0based' or '1based', got {self.coordinate_system!r}"
            )
        if self.interval_type not in ("half_open", "closed"):
            raise ValueError(
                f"interval_type must be 'half_open' or 'closed', got {self.interval_type!r}"
            )


class Tables:
    """Name -> Table lookup (reference: src/giql/table.py:96-136)."""

    def __init__(self) -> None:
        self._tables: dict[str, Table] = {}

    def register(self, name: str, table: Table) -> None:
        self._tables[name] = table

    def get(self, name: str) -> Optional[Table]:
        return self._tables.get(name)

    def __contains__(self, name: str) -> bool:
        return name in self._tables

    def __iter__(self) -> Iterator[Table]:
        return iter(self._tables.values())


def canonical_offsets(table) -> tuple[int, int]:
    """``(start_adj, end_adj)`` with canonical = raw + adj (0-based half-open).

    reference: src/giql/canonical.py:16-29 (start: 1-based -> start - 1) and :32-52
    (end: 0-based closed -> end + 1; 1-based half-open -> end - 1; otherwise identity).
    ``table`` may be None (treated as 0-based half-open, as the reference does).
    """
    if table is None:
        return 0, 0
    s_adj = 0 if table.coordinate_system == "0based" else -1
    key = (table.coordinate_system, table.interval_type)
    if key == ("0based", "closed"):
        e_adj = 1
    elif key == ("1based", "half_open"):
        e_adj = -1
    else:
        e_adj = 0
    return s_adj, e_adj

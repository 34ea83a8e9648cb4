"""giql_b200 — B200-native execution engine for GIQL's genomic-interval operators.

The data path is hand-written sm_100a CUDA behind the C ABI in ``include/giqlb.h``
(``giql_b200/_lib/libgiqlb200.so``); this package is the Python host layer that mirrors the reference's
operator surface (abdenlab/giql).  There is no CPU fallback.
"""
from ._cabi import EngineError, EngineUnavailable
from .ranges import parse_range
from .table import Table, Tables, canonical_offsets

__all__ = ["Engine", "DeviceTable", "Dictionary", "PairResult", "Table", "Tables", "canonical_offsets",
           "EngineError", "EngineUnavailable", "bind_host_to_gpu", "Session", "UnsupportedQuery", "parse_range"]


def __getattr__(name):
    # torch is imported lazily so that `import giql_b200` (Table, canonical_offsets) stays light
    if name in ("Engine", "DeviceTable", "Dictionary", "PairResult", "bind_host_to_gpu"):
        from . import engine

        return getattr(engine, name)
    if name in ("Session", "UnsupportedQuery", "EngineBackend"):
        from . import frontend

        return getattr(frontend, name)
    raise AttributeError(name)

"""The C-ABI library loads and exports every symbol include/giqlb.h declares (no GPU, no compute)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "giqlb.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(giqlb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from giql_b200 import _cabi

    if not _cabi.LIB_PATH.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "giql_b200" / "csrc"), "-s", "-j8"])
    return _cabi.LIB_PATH


def test_header_declares_the_operator_surface():
    names = _declared()
    for required in ("giqlb_dict_encode", "giqlb_index_build", "giqlb_intersects_plan", "giqlb_intersects_emit",
                     "giqlb_intersects_join", "giqlb_intersects_exists", "giqlb_intersects_count_rows",
                     "giqlb_distance"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in giqlb.h but not exported: {missing}"


def test_ctypes_prototypes_cover_the_header(lib_path):
    from giql_b200 import _cabi

    assert sorted(_cabi.PROTOTYPES) == _declared()
    assert _cabi.load().giqlb_abi_version() == 1


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "--list-elf", str(lib_path)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_engine_fails_loudly_without_gpu():
    import torch

    import giql_b200

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(giql_b200.EngineUnavailable):
        giql_b200.Engine()


def test_product_never_imports_oracle():
    for py in (ROOT / "giql_b200").rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), py

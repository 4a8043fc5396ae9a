"""The C-ABI library loads and exports every symbol include/svsb200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "svsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svsb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for name in ("svsb200_index_create", "svsb200_search", "svsb200_search_device", "svsb200_merge_topk_device",
                 "svsb200_get_counters", "svsb200_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    from scalablevectorsearch_b200 import _lib
    lib = _lib.lib()
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"libsvsb200.so lacks {missing}"
    assert sorted(_lib.SYMBOLS) == header_symbols(), "python binding list out of sync with the header"
    assert lib.svsb200_version() == 201


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device index creation must fail loudly (never silently compute on the CPU)."""
    import numpy as np
    from scalablevectorsearch_b200 import DistanceType, Svsb200Error, Vamana, _lib
    if _lib.lib().svsb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    data = np.zeros((4, 8), dtype=np.float32)
    graph = np.zeros((4, 3), dtype=np.uint32)
    with pytest.raises(Svsb200Error, match="no CUDA device"):
        Vamana.from_arrays(data, graph, 0, DistanceType.L2)


def test_product_never_touches_the_oracle():
    """The product path must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "scalablevectorsearch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f in ("synthetic.py",) and False, \
                    f"{os.path.join(dirpath, f)} mentions the oracle"
    out = os.popen(f"ldd {os.path.join(pkg, 'libsvsb200.so')}").read()
    assert "oracle" not in out and "svsref" not in out


def test_oracle_library_exports_every_declared_symbol():
    """The checker's own header and shared object must agree (a declared-but-undefined function only shows up when a
    GPU test first calls it)."""
    from oracle.bindings import OracleLib
    OracleLib()   # builds oracle/liboracle.so on demand
    header = open(os.path.join(ROOT, "oracle", "vamana_oracle.h")).read()
    names = set(re.findall(r"\b(oracle_[a-z0-9_]+)\s*\(", header))
    assert len(names) > 10
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing

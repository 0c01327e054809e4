"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rsb.h
declares, the ctypes table covers them, and argument validation works without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rsb.h")).read()
    return sorted(set(re.findall(r"RSB_API[^;(]*?\b(rsb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G

    G.build()
    from recsys_benchmark_b200 import _lib

    return _lib.load()


def test_header_declares_symbols():
    syms = _header_symbols()
    assert "rsb_lookup_fwd" in syms and "rsb_segment_reduce_apply" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(lib):
    from recsys_benchmark_b200 import _lib

    for s in _header_symbols():
        assert hasattr(lib, s), f"librsb.so does not export {s}"
        assert s in _lib.PROTOTYPES, f"ctypes table misses {s}"
    assert sorted(_lib.PROTOTYPES) == _header_symbols()


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.rsb_version()
    assert b"bad argument" in lib.rsb_error_string(10001)
    assert b"unsupported" in lib.rsb_error_string(10002)
    assert b"workspace" in lib.rsb_error_string(10003)


def test_row_width_support(lib):
    for w in [1, 3, 4, 7, 8, 12, 16, 32, 64, 128]:
        assert lib.rsb_row_width_supported(w) == 1
    for w in [0, -1, 33, 130, 256]:
        assert lib.rsb_row_width_supported(w) == 0


def test_workspace_queries(lib):
    assert lib.rsb_sort_workspace_bytes(0) > 0
    assert lib.rsb_sort_workspace_bytes(1 << 20) >= 16 * (1 << 20)
    assert lib.rsb_segment_workspace_bytes(1 << 20, 16) >= 2 * ((1 << 20) // 32) * 16 * 4
    assert lib.rsb_small_table_workspace_bytes(1 << 20, 16) == -1  # not a small table


def test_bad_arguments_are_rejected_before_any_launch(lib):
    BAD = 10001
    assert lib.rsb_lookup_fwd(0, None, 0, None, 4, 3, 16, None, 10, 10, None, 0, None, 0, None, None, None, None,
                              None, None, None, None, None, None) == BAD
    assert lib.rsb_lookup_bwd_rows(0, None, 4, 3, 16, None, 10, None, 0, None, 0, None, None, None, None, None,
                                   None, None, None, None) == BAD
    assert lib.rsb_sort_rows(None, -1, 10, 0, 0, None, None, None, 0, None) == BAD
    assert lib.rsb_sort_rows(None, 0, 10, 0, 0, None, None, None, 0, None) == 0  # empty input is fine
    # row ids are 32-bit sort keys and 0xffffffff is the segmented reduction's sentinel: a key space of 2^32 rows (or
    # 2^32 lookups in one call) is refused before any launch instead of wrapping around
    assert lib.rsb_sort_rows(None, 4, (1 << 32), 0, 0, None, None, None, 0, None) == BAD
    assert lib.rsb_sort_rows(None, 4, (1 << 40), 0, 0, None, None, None, 0, None) == BAD
    assert lib.rsb_sort_rows(None, (1 << 32), 10, 0, 0, None, None, None, 0, None) == BAD
    assert lib.rsb_sort_rows(None, 4, (1 << 32) - 1, 0, 0, None, None, None, 0, None) == BAD  # (NULL buffers; size ok)
    assert lib.rsb_segment_reduce_apply(7, None, None, 4, None, 16, None, None, None, 0.0, 0.9, 0.999, 1e-8, 1,
                                        None, 0, None) == BAD
    assert lib.rsb_pep_threshold_table(None, None, 0, 4, 4, None, None, None) == BAD
    assert lib.rsb_mask_table(None, None, 4, None, None) == BAD


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from recsys_benchmark_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()

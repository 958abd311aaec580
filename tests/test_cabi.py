"""The C-ABI shared library: builds for sm_100a, loads, and exports exactly what include/*.h
declares.  No compute calls here (no GPU needed)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
from gcgcn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gcgcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcgcn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gcgcn_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_counters(lib):
    assert b"sm_100a" in lib.gcgcn_version()
    assert lib.gcgcn_launch_count() == 0 or lib.gcgcn_launch_count() > 0
    assert lib.gcgcn_workspace_bytes(243, 6207, 8) > 3 * 243 * 1024 * 4
    assert lib.gcgcn_block_saved_bytes(243, 6207, 8) > 0


def test_invalid_arguments_are_errors_not_fallbacks(lib):
    rc = lib.gcgcn_edge_mean_fwd(None, None, 0, None, None)
    assert rc == -1 and b"NULL" in lib.gcgcn_last_error()
    rc = lib.gcgcn_pair_gather_fwd(None, None, 4, None, 0, None, None, None, None, None, None, None)
    assert rc == -1
    with pytest.raises(_lib.GcgcnError):
        _lib.call("gcgcn_gemm", 0, 0, -1, 1, 1, 1.0, None, 1, None, 1, 0.0, None, 1, None, None, 0, None)


def test_library_is_built_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_tile_blocks_switch_and_batch_struct_layout(lib):
    """gcgcn_set_tile_blocks returns the previous setting (no device work), and the ctypes mirror of gcgcn_batch has the
    C struct's layout: three 4-byte-aligned tail fields after class_end (the packing hint of the tile kernel)."""
    first = _lib.set_tile_blocks(False)
    assert _lib.set_tile_blocks(True) is False
    assert _lib.set_tile_blocks(first) is True
    assert lib.gcgcn_set_tile_blocks(1 if first else 0) in (0, 1)
    B = _lib.Batch
    assert B.tile_doc.offset == B.class_end.offset + 16 and B.num_tiles.offset == B.tile_doc.offset + 8
    assert B.tile_rows.offset == B.num_tiles.offset + 4 and ctypes.sizeof(B) == B.tile_rows.offset + 4
    text = open(os.path.join(ROOT, "include", "gcgcn_b200.h")).read()
    rows = int(re.search(r"#define GCGCN_TILE_ROWS (\d+)", text).group(1))
    from gcgcn_b200.batch import TILE_ROWS
    assert rows == TILE_ROWS == 96

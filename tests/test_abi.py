"""The C-ABI library loads and exports every symbol include/gg_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT
from graphgym_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gg_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gg_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(handle, name), f'{name} is declared in gg_b200.h but not exported'


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_pure_host_entry_points():
    L = _lib.lib()
    assert L.gg_version() >= 100
    assert L.gg_layout_capacity(10, 5, 0) == 10   # KEEP
    assert L.gg_layout_capacity(10, 5, 1) == 15   # ADD_REMAINING appends N loops
    assert L.gg_layout_capacity(10, 5, 3) == 10   # REMOVE
    assert L.gg_layout_build_workspace_bytes(1000, 100, 1) > 0
    assert L.gg_sort_pairs_workspace_bytes(1000) >= 8000
    assert L.gg_gemm_tn_workspace_bytes(10000, 128, 128) > 0


def test_no_symbol_leaks():
    """Only gg_* is exported: the kernels and helpers stay hidden."""
    import subprocess
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if ' T ' in l]
    assert exported and all(s.startswith('gg_') for s in exported), exported

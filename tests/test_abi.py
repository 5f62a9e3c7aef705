"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every
symbol include/nrgbd.h declares; the ctypes table mirrors the header one to one."""
import ctypes
import os
import re

from neuralrgbd_b200 import _lib
from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'nrgbd.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nrgbd_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 10
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), 'libnrgbd.so does not export %s' % n


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES.keys()) == _declared()


def test_version_and_error_string_without_gpu():
    L = _lib.lib()
    assert L.nrgbd_abi_version() == 1
    assert L.nrgbd_last_error() is not None
    L.nrgbd_reset_launch_count()
    assert L.nrgbd_launch_count() == 0


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libnrgbd.so')
    monkeypatch.setattr(_lib, '_lib', None)
    try:
        _lib.lib()
        assert False, 'expected NrgbdError'
    except _lib.NrgbdError as e:
        assert 'no fallback' in str(e)

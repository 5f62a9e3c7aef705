"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every
symbol include/nrgbd.h declares; the ctypes table mirrors the header one to one."""
import ctypes
import os
import re

from neuralrgbd_b200 import _lib
from tests.conftest import ROOT


def _declared(header='nrgbd.h'):
    src = open(os.path.join(ROOT, 'include', header)).read()
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


def test_dev_knobs_live_in_their_own_header():
    """Development probes / A-B switches are not part of the product ABI: separate header, separate
    ctypes table, exported by the library, and no environment variable is read by the native code."""
    dev = _declared('nrgbd_dev.h')
    assert sorted(_lib.DEV_SIGNATURES.keys()) == dev and len(dev) >= 4
    assert not set(dev) & set(_declared())
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in dev:
        assert hasattr(L, n)
    import glob
    for f in glob.glob(os.path.join(ROOT, 'neuralrgbd_b200', 'csrc', '*')):
        assert 'getenv' not in open(f).read(), '%s reads the environment' % f


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


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    """include/nrgbd.h must be consumable from C (the boundary is a C ABI, not C++), and the one struct that crosses it
    must have the layout the ctypes mirror assumes."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        import pytest
        pytest.skip('no C compiler')
    hdr_dir = os.path.join(ROOT, 'include')
    r = subprocess.run([gcc, '-fsyntax-only', '-x', 'c', '-std=c99', '-Wall', '-Werror', os.path.join(hdr_dir, 'nrgbd.h')],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nrgbd.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(nrgbd_bn_input),'
                   'offsetof(nrgbd_bn_input,stats),offsetof(nrgbd_bn_input,count),offsetof(nrgbd_bn_input,gamma),'
                   'offsetof(nrgbd_bn_input,beta),offsetof(nrgbd_bn_input,running_mean),offsetof(nrgbd_bn_input,running_var),'
                   'offsetof(nrgbd_bn_input,eps),offsetof(nrgbd_bn_input,momentum),offsetof(nrgbd_bn_input,relu),'
                   'offsetof(nrgbd_bn_input,C));return 0;}\n')
    exe = tmp_path / 'layout'
    r = subprocess.run([gcc, '-I', hdr_dir, str(src), '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    B = _lib.BnInput
    want = [ctypes.sizeof(B)] + [getattr(B, f).offset for f in ('stats', 'count', 'gamma', 'beta', 'running_mean', 'running_var',
                                                                 'eps', 'momentum', 'relu', 'C')]
    assert got == want


def test_c_program_links_and_uses_the_library(tmp_path):
    """examples/write_depth_pgm.c - a plain C host (no Python, torch or CUDA headers) - compiles against include/nrgbd.h,
    links libnrgbd.so and produces the reference's 16-bit PGM format."""
    import shutil
    import subprocess
    import numpy as np
    from oracle import export_oracle as E
    gcc = shutil.which('gcc')
    if gcc is None:
        import pytest
        pytest.skip('no C compiler')
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / 'write_depth_pgm')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'examples', 'write_depth_pgm.c'),
                        '-L', libdir, '-lnrgbd', '-Wl,-rpath,' + libdir, '-o', exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / 'd.pgm')
    r = subprocess.run([exe, out, '64', '48'], capture_output=True, text=True)
    assert r.returncode == 0 and 'ABI 1' in r.stdout, r.stdout + r.stderr
    yy, xx = np.mgrid[0:48, 0:64]
    assert open(out, 'rb').read() == E.pgm16_bytes((500 + 7 * xx + 3 * yy).astype(np.uint16))

"""The C-ABI shared library loads on a host without a GPU, exports every symbol include/dnmf.h declares,
and rejects bad arguments before touching the device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'dnmf.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dnmf_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported_and_bound():
    from pydnmfk_b200 import _lib as L
    names = _declared()
    assert len(names) >= 35
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), 'libdnmf.so does not export %s' % n
        assert n in L.SIGNATURES, 'pydnmfk_b200/_lib.py does not bind %s' % n
    assert sorted(L.SIGNATURES) == names


def test_version_and_workspace_queries_need_no_gpu():
    from pydnmfk_b200 import _lib as L
    assert 'sm_100a' in L.version()
    for op in (L.OP_AH, L.OP_WTA, L.OP_KL_UHT, L.OP_KL_WTU, L.OP_GRAM, L.OP_RESIDUAL, L.OP_SUMS, L.OP_NNZ):
        for dt in (L.F32, L.F64):
            b = L.workspace_bytes(op, 65536, 65536, 32, dt)
            assert b >= 0 and b < (8 << 30)
    # workspace grows with the problem, and is a pure function of the shape
    assert L.workspace_bytes(L.OP_WTA, 4096, 4096, 32, L.F32) == L.workspace_bytes(L.OP_WTA, 4096, 4096, 32, L.F32)


def test_argument_errors_are_reported_without_a_device():
    from pydnmfk_b200 import _lib as L
    with pytest.raises(L.DnmfError) as e:
        L.call('dnmf_ah', 1, 8, 1, 8, 1, 8, 8, 8, L.MAX_K + 1, L.F32, 0, None, 0, None)
    assert e.value.status == -2 and 'DNMF_MAX_K' in str(e.value)
    with pytest.raises(L.DnmfError) as e:
        L.call('dnmf_ah', None, 8, None, 8, None, 8, 8, 8, 4, L.F32, 0, None, 0, None)
    assert e.value.status == -1
    with pytest.raises(L.DnmfError) as e:
        L.call('dnmf_wta', 1, 8, 1, 8, 1, 8, 8, 8, 4, 0, 7, 0, None, 0, None)   # bad dtype
    assert e.value.status == -1
    with pytest.raises(L.DnmfError):
        L.workspace_bytes(99, 8, 8, 4, L.F32)


def test_no_cpu_fallback():
    """The product refuses to run without a CUDA device instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pydnmfk_b200 import device as D
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        D.default_ops()
    # and nothing in the product imports the oracle
    pkg = os.path.join(ROOT, 'pydnmfk_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, fn)).read().replace('no oracle', ''), fn


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/dnmf.h compiles as C99 and a C program linked against libdnmf.so can call it (no GPU needed for the
    version / error / workspace queries)."""
    import shutil
    import subprocess
    from pydnmfk_b200 import _lib as L
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    src = tmp_path / 'cabi.c'
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "dnmf.h"\n'
        'int main(void) {\n'
        '  const char* v = dnmf_version();\n'
        '  int64_t ws = dnmf_workspace_bytes(DNMF_OP_WTA, 4096, 4096, 32, DNMF_F32);\n'
        '  int rc = dnmf_ah(0, 8, 0, 8, 0, 8, 8, 8, 4, DNMF_F32, DNMF_MATH_ACCURATE, 0, 0, 0);\n'
        '  printf("%s|%lld|%d|%s\\n", v, (long long)ws, rc, dnmf_last_error());\n'
        '  return strstr(v, "sm_100a") && ws > 0 && rc == DNMF_E_ARG ? 0 : 1;\n}\n')
    exe = tmp_path / 'cabi'
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                    '-L', libdir, '-l:libdnmf.so', '-Wl,-rpath,' + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert 'sm_100a' in out and 'bad argument' in out


def test_ctypes_signatures_match_the_header_arity():
    """Every entry point's ctypes argument list has exactly as many entries as the C declaration has parameters (a
    signature that drifts from the header corrupts the call silently: ctypes does not know the C prototype)."""
    from pydnmfk_b200 import _lib as L
    src = open(os.path.join(ROOT, 'include', 'dnmf.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    decls = re.findall(r'\b(dnmf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S)
    assert len(decls) >= 60
    for name, params in decls:
        params = ' '.join(params.split())
        n = 0 if params in ('', 'void') else params.count(',') + 1
        res, args = L.SIGNATURES[name]
        assert len(args) == n, '%s: header has %d parameters, _lib.py binds %d' % (name, n, len(args))

"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports
exactly the functions include/gpslim_b200.h declares, the ctypes structs match the header's
constants, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gpslim_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gps_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from gpflowSlim._backend import lib
    cdll = lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(cdll, n), 'missing export ' + n
    assert sorted(lib.SIGNATURES) == names, set(lib.SIGNATURES) ^ set(names)
    assert cdll.gps_version() == 100


def test_header_constants_match_ctypes():
    from gpflowSlim._backend import lib
    src = open(HEADER).read()
    for name in ('GPS_MAX_PRIMS', 'GPS_MAX_DIMS', 'GPS_MAX_OPS', 'GPS_MAX_SLOTS', 'GPS_MAX_THETA'):
        val = int(re.search(r'#define %s (\d+)' % name, src).group(1))
        assert getattr(lib, name) == val
    assert ctypes.sizeof(lib.gps_prim) == 4 * (4 + lib.GPS_MAX_DIMS)
    assert ctypes.sizeof(lib.gps_op) == 32
    assert ctypes.sizeof(lib.gps_kernel_desc) == 16 + lib.GPS_MAX_PRIMS * ctypes.sizeof(lib.gps_prim) \
        + lib.GPS_MAX_OPS * 32
    assert ctypes.sizeof(lib.DLTensor) == 48


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    import gpflowSlim as gpf
    from gpflowSlim._backend import lib
    h = ctypes.c_void_p()
    assert lib.load().gps_create(0, ctypes.byref(h)) != 0
    k = gpf.kernels.RBF(2)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        k.K(torch.zeros(3, 2, dtype=torch.float64))


def test_first_handle_request_does_not_deadlock():
    """Regression: the very first handle_for() call loads the library under the module lock."""
    import subprocess
    import sys
    code = ('import sys; sys.path.insert(0, %r); import torch\n'
            'from gpflowSlim._backend import lib\n'
            'try:\n'
            '    lib.handle_for(torch.device("cuda", 0)); print("HANDLE")\n'
            'except RuntimeError as e:\n'
            '    print("RAISED")\n') % os.path.join(ROOT, 'gpflow-slim_b200')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert ('RAISED' in out.stdout) or ('HANDLE' in out.stdout), out.stderr[-2000:]

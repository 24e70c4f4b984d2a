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


def test_packed_view_equals_fieldwise_dltensor():
    """lib.view() fills the DLTensor + shape/stride words with one struct.pack_into; the result
    must be byte-identical to a DLTensor built field by field with ctypes (what the C side reads:
    include/gpslim_b200.h, DLPack v0.8 layout)."""
    from gpflowSlim._backend import lib
    base = torch.arange(12 * 9, dtype=torch.float64).reshape(12, 9)
    cases_ = [base, base[2:7, 1:6], base[:, :1], base[3], base[0:1], torch.zeros(0, 4, dtype=torch.float64),
              torch.tensor(3.5, dtype=torch.float64), torch.arange(5, dtype=torch.int64)]
    for t in cases_:
        v = lib.view(t)
        tt = v.tensor
        nd = tt.dim()
        shape = (ctypes.c_int64 * max(nd, 1))(*tt.shape)
        strides = (ctypes.c_int64 * max(nd, 1))(*tt.stride())
        want = lib.DLTensor(ctypes.c_void_p(tt.data_ptr()), lib.DLDevice(1, 0), nd,
                            lib.DLDataType(2 if tt.dtype == torch.float64 else 0, 64, 1), shape, strides, 0)
        got = v.dl
        assert (got.data or 0) == (want.data or 0)          # None for an empty tensor's null pointer
        assert (got.device.device_type, got.device.device_id) == (1, 0)
        assert got.ndim == want.ndim == nd
        assert (got.dtype.code, got.dtype.bits, got.dtype.lanes) == (want.dtype.code, 64, 1)
        assert got.byte_offset == 0
        assert [got.shape[i] for i in range(nd)] == list(tt.shape)
        assert [got.strides[i] for i in range(nd)] == list(tt.stride())
        # the pointers point INTO the view's own buffer, which the view keeps alive
        a = ctypes.addressof(v.buf)
        assert ctypes.cast(got.shape, ctypes.c_void_p).value == a + 48
        assert ctypes.cast(got.strides, ctypes.c_void_p).value == a + 64
        raw = bytes(v.buf)[:24] + bytes(v.buf)[40:48]
        ref_raw = bytes(want)[:24] + bytes(want)[40:48]
        assert raw == ref_raw
    with pytest.raises(TypeError):
        lib.view(torch.zeros(2, 2, dtype=torch.float32))
    with pytest.raises(ValueError):
        lib.view(torch.zeros(4, 4, dtype=torch.float64).t())

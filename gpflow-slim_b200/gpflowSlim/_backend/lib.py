"""ctypes binding of libgpslim_b200.so (the C ABI declared in include/gpslim_b200.h).

There is NO CPU fallback: if the shared library is missing or no sm_100 GPU is visible, every
compute entry point raises.  torch is used only to own device memory and streams; tensors are
handed to the library as borrowed DLPack `DLTensor` views (data_ptr / shape / strides).
"""
import ctypes
import os
import struct
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), '_lib', 'libgpslim_b200.so')

GPS_MAX_PRIMS = 16
GPS_MAX_DIMS = 32
GPS_MAX_OPS = 48
GPS_MAX_SLOTS = 96
GPS_MAX_THETA = 512

(GPS_RBF, GPS_EXPONENTIAL, GPS_MATERN12, GPS_MATERN32, GPS_MATERN52, GPS_LINEAR,
 GPS_PERIODIC) = range(7)
(GPS_OP_CONST, GPS_OP_ADD, GPS_OP_MUL, GPS_OP_COPY, GPS_OP_LINEAR, GPS_OP_PRODUCT) = range(6)
TRI_NONE, TRI_LOWER, TRI_UPPER = 0, 1, 2


class CholeskyError(ArithmeticError):
    """Raised when a matrix is not positive definite (tf.cholesky raises InvalidArgumentError)."""


class DLDevice(ctypes.Structure):
    _fields_ = [('device_type', ctypes.c_int32), ('device_id', ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [('code', ctypes.c_uint8), ('bits', ctypes.c_uint8), ('lanes', ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [('data', ctypes.c_void_p), ('device', DLDevice), ('ndim', ctypes.c_int32),
                ('dtype', DLDataType), ('shape', ctypes.POINTER(ctypes.c_int64)),
                ('strides', ctypes.POINTER(ctypes.c_int64)), ('byte_offset', ctypes.c_uint64)]


class gps_prim(ctypes.Structure):
    _fields_ = [('type', ctypes.c_int32), ('ndims', ctypes.c_int32), ('ard', ctypes.c_int32),
                ('theta_off', ctypes.c_int32), ('dims', ctypes.c_int32 * GPS_MAX_DIMS)]


class gps_op(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('op', 'dst', 'a', 'b', 'c', 'd', 'n', 'pad')]


class gps_kernel_desc(ctypes.Structure):
    _fields_ = [('n_prims', ctypes.c_int32), ('n_ops', ctypes.c_int32),
                ('n_theta', ctypes.c_int32), ('out_slot', ctypes.c_int32),
                ('prims', gps_prim * GPS_MAX_PRIMS), ('ops', gps_op * GPS_MAX_OPS)]


_P = ctypes.POINTER
_T = _P(DLTensor)
_H = ctypes.c_void_p
_D = _P(gps_kernel_desc)

# name -> argtypes; mirrors include/gpslim_b200.h one to one (tests check the export list)
SIGNATURES = {
    'gps_create': [ctypes.c_int, _P(_H)],
    'gps_destroy': [_H],
    'gps_set_stream': [_H, ctypes.c_void_p],
    'gps_last_error': [_H],
    'gps_version': [],
    'gps_set_option': [_H, ctypes.c_char_p, ctypes.c_int64],
    'gps_profile_read': [_H, _P(ctypes.c_double), _P(ctypes.c_double), _P(ctypes.c_int64),
                         ctypes.c_int],
    'gps_gram_fwd': [_H, _D, _T, _T, _T, ctypes.c_double, ctypes.c_int, _T],
    'gps_gram_bwd': [_H, _D, _T, _T, _T, _T, _T, _T],
    'gps_kdiag_fwd': [_H, _D, _T, _T, _T],
    'gps_kdiag_bwd': [_H, _D, _T, _T, _T, _T, _T],
    'gps_potrf': [_H, _T, ctypes.c_int, _P(ctypes.c_int)],
    'gps_potri': [_H, _T, _T],
    'gps_chol_bwd': [_H, _T, _T, _T, _T],
    'gps_trsm_bwd': [_H, _T, _T, _T, _T, _T, _T],
    'gps_trsm_rlt': [_H, _T, _T],
    'gps_tri_inv_t': [_H, _T, _T],
    'gps_gemm_nt': [_H, ctypes.c_double, _T, _T, ctypes.c_double, _T, ctypes.c_int, ctypes.c_int,
                    ctypes.c_int],
    'gps_transpose': [_H, _T, _T],
    'gps_gemm_nt_rowmap': [_H, ctypes.c_double, _T, _T, ctypes.c_double, _T, _T, ctypes.c_int64,
                           ctypes.c_double],
    'gps_trsm_rlt_prefix': [_H, _T, _T, _P(ctypes.c_int64)],
    'gps_trsm_rln_prefix': [_H, _T, _T, _T, _P(ctypes.c_int64)],
    'gps_gpr_weight_rows': [_H, _T, _T, _T, ctypes.c_int64],
    'gps_sum_log_diag': [_H, _T, _T],
    'gps_row_sumsq': [_H, ctypes.c_double, _T, ctypes.c_double, _T],
    'gps_gpr_nlml_fwd_bwd': [_H, _D, _T, _T, _T, ctypes.c_double, ctypes.c_int, _T, _T, _T,
                             _P(ctypes.c_int)],
    'gps_gpr_predict': [_H, _D, _T, _T, _T, ctypes.c_double, _T, ctypes.c_int, _T, _T,
                        _P(ctypes.c_int)],
}

_lib = None
_lock = threading.RLock()
_handles = {}


def load():
    """Load the shared library (works without a GPU; compute calls do not)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    'libgpslim_b200.so not found at %s -- build it with '
                    '`python -c "import __graft_entry__ as g; g.build()"` or '
                    '`make -C gpflow-slim_b200/csrc`.  There is no CPU fallback.' % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, argtypes in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = ctypes.c_char_p if name == 'gps_last_error' else ctypes.c_int
            _lib = lib
    return _lib


class Handle(object):
    """One library handle per (process, device)."""

    def __init__(self, device_index):
        self.lib = load()
        self.device_index = device_index
        self.ptr = _H()
        self._stream = None
        rc = self.lib.gps_create(device_index, ctypes.byref(self.ptr))
        if rc != 0:
            raise RuntimeError('gps_create(device=%d) failed (rc=%d): an sm_100 (B200) GPU is '
                               'required; there is no CPU fallback' % (device_index, rc))

    def check(self, rc):
        if rc == 0:
            return
        msg = self.lib.gps_last_error(self.ptr)
        msg = msg.decode() if msg else ''
        if rc > 0:
            raise CholeskyError(msg or 'matrix is not positive definite (info=%d)' % rc)
        raise ValueError('libgpslim_b200: %s (rc=%d)' % (msg, rc))

    def sync_stream(self):
        """Make the library launch on torch's current stream (only this method sets the handle's
        stream, so the call is skipped while the stream has not changed)."""
        s = torch.cuda.current_stream(self.device_index).cuda_stream
        if s != self._stream:
            self.lib.gps_set_stream(self.ptr, ctypes.c_void_p(s))
            self._stream = s

    def set_option(self, name, value):
        self.check(self.lib.gps_set_option(self.ptr, name.encode(), int(value)))

    def profile_read(self, reset=True):
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        self.check(self.lib.gps_profile_read(self.ptr, ctypes.byref(ms), ctypes.byref(fl),
                                             ctypes.byref(n), int(reset)))
        return ms.value, fl.value, n.value


def handle_for(tensor_or_device):
    dev = tensor_or_device.device if isinstance(tensor_or_device, torch.Tensor) \
        else torch.device(tensor_or_device)
    if dev.type != 'cuda':
        raise RuntimeError('gpflowSlim (B200) computes on CUDA tensors only -- got a %s tensor; '
                           'there is no CPU fallback' % dev.type)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    h = _handles.get(idx)
    if h is None:
        with _lock:
            h = _handles.get(idx)
            if h is None:
                h = _handles[idx] = Handle(idx)
    h.sync_stream()
    return h


# DLTensor (48 bytes) followed by shape[2] and strides[2]: one 80-byte buffer per view, filled by a
# single struct.pack_into -- a third of the host time of building the ctypes objects field by
# field; it is paid on every library call (37 per SVGP step, ~25 per panel of the distributed
# factorisation).
_DL_PACK = struct.Struct('<QiiiBBHQQQ4q')
assert _DL_PACK.size == 80 and ctypes.sizeof(DLTensor) == 48


class _View(object):
    """Keeps the shape/stride words alive next to the DLTensor that points at them."""
    __slots__ = ('buf', 'dl', 'tensor', 'ref')

    def __init__(self, t):
        if t.dtype is torch.float64:
            code = 2
        elif t.dtype is torch.int64:
            code = 0
        else:
            raise TypeError('float64 (or int64 index) tensor required, got %s' % t.dtype)
        nd = t.dim()
        self.tensor = t
        self.buf = buf = ctypes.create_string_buffer(80)
        addr = ctypes.addressof(buf)
        dev = t.device
        idx = dev.index if dev.index is not None else 0
        sh, st = t.shape, t.stride()
        if nd == 2:
            s0, s1, t0, t1 = sh[0], sh[1], st[0], st[1]
        elif nd == 1:
            s0, s1, t0, t1 = sh[0], 0, st[0], 0
        elif nd == 0:
            s0 = s1 = t0 = t1 = 0
        else:
            raise ValueError('at most 2 dimensions, got %d' % nd)
        _DL_PACK.pack_into(buf, 0, t.data_ptr(), 2 if dev.type == 'cuda' else 1, idx, nd, code, 64, 1,
                           addr + 48, addr + 64, 0, s0, s1, t0, t1)
        self.dl = DLTensor.from_buffer(buf)
        self.ref = ctypes.byref(self.dl)


def view(t):
    """Borrowed DLTensor view of a torch tensor (row-major with unit inner stride required)."""
    if t is None:
        return None
    if t.dim() == 2 and t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError('tensor must have unit innermost stride')
    if t.dim() == 0:
        t = t.reshape(1)
    return _View(t)


def ref(v):
    return v.ref if v is not None else None

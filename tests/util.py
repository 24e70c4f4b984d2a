import numpy as np
import torch


def dev():
    return torch.device('cuda', 0)


def conv(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64, device=dev())


def relerr(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def assert_close(a, b, rtol, what=''):
    e = relerr(a, b)
    assert e < rtol, '%s: relative error %.3e exceeds %.1e' % (what, e, rtol)

"""Control-path coefficient construction with torch ops (runs on any device).

Host-side data preparation either side of the hot path: the reference builds these once per
dataset (/root/reference/benchmark_classification/datasets/common.py:82-84 via torchcde;
benchmark_forecasting/datasets/common.py:79-81 via the in-tree natural spline) and feeds the
packed ``[B, K-1, 4C]`` tensor to ``set_X``.  Packing: ``cat(a, b, two_c, three_d)``.
"""
import ctypes

import torch


def hermite_backward_difference_coeffs(x, t):
    """x ``[..., K, C]`` (no NaNs), t ``[K]`` -> ``[..., K-1, 4C]``.  Cubic Hermite pieces whose
    start slope is the previous interval's secant (the first piece reuses its own)."""
    h = (t[1:] - t[:-1]).unsqueeze(-1)
    dx = x[..., 1:, :] - x[..., :-1, :]
    m1 = dx / h
    m0 = torch.cat((m1[..., :1, :], m1[..., :-1, :]), dim=-2)
    two_c = 2 * (3 * (dx / h - m0) - m1 + m0) / h
    three_d = (1 / h ** 2) * (m1 - m0) - two_c / h
    return torch.cat((x[..., :-1, :], m0, two_c, three_d), dim=-1)


def natural_cubic_coeffs(x, t):
    """x ``[..., K, C]`` (no NaNs), t ``[K]`` -> ``[..., K-1, 4C]`` natural cubic spline
    (second derivative zero at both ends); knot slopes from a Thomas sweep along K."""
    K = x.shape[-2]
    h = t[1:] - t[:-1]
    r = h.reciprocal()
    dx = x[..., 1:, :] - x[..., :-1, :]
    if K == 2:
        z = torch.zeros_like(dx)
        return torch.cat((x[..., :-1, :], dx * r[:, None], z, z), dim=-1)
    s = 3 * dx * (r ** 2)[:, None]
    rhs = torch.zeros_like(x)
    rhs[..., :-1, :] += s
    rhs[..., 1:, :] += s
    diag = torch.zeros(K, dtype=x.dtype, device=x.device)
    diag[:-1] += 2 * r
    diag[1:] += 2 * r
    cp = torch.empty(K, dtype=x.dtype, device=x.device)      # modified diagonal
    d = [rhs[..., 0, :]]
    cp[0] = diag[0]
    for i in range(1, K):
        w = r[i - 1] / cp[i - 1]
        cp[i] = diag[i] - w * r[i - 1]
        d.append(rhs[..., i, :] - w * d[i - 1])
    k = [None] * K
    k[K - 1] = d[K - 1] / cp[K - 1]
    for i in range(K - 2, -1, -1):
        k[i] = (d[i] - r[i] * k[i + 1]) / cp[i]
    kd = torch.stack(k, dim=-2)
    k0, k1 = kd[..., :-1, :], kd[..., 1:, :]
    rr = r[:, None]
    two_c = (6 * dx * rr - 4 * k0 - 2 * k1) * rr
    three_d = (-6 * dx * rr + 3 * (k0 + k1)) * rr ** 2
    return torch.cat((x[..., :-1, :], k0, two_c, three_d), dim=-1)


def _cuda_args(x, t, who):
    if not x.is_cuda or x.dim() != 3:
        raise ValueError(f"snsde: {who} needs a CUDA tensor [B, K, C]")
    x = x.detach().to(torch.float32).contiguous()
    t = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if t.shape != (x.shape[1],):
        raise ValueError("snsde: knots must be [K]")
    return x, t


def fill_missing_cuda(x, t):
    """NaN fill torchcde applies before the Hermite builder (linear in t between observed neighbours, first
    observed value at the head, forward fill at the tail) on device: ``snsde_fill_missing``."""
    from . import _lib
    x, t = _cuda_args(x, t, "fill_missing_cuda")
    B, K, C = x.shape
    out = torch.empty_like(x)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.load().snsde_fill_missing(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(t.data_ptr()), B, K, C,
                                              ctypes.c_void_p(out.data_ptr()), x.device.index or 0, ctypes.c_void_p(stream)))
    return out


def natural_coeffs_cuda(x, t, out=None, missing=False):
    """CUDA version of :func:`natural_cubic_coeffs` (bit-identical to that torch-op chain) for a NaN-free CUDA tensor
    ``x [B, K, C]``: knot-only Thomas factors once, then one thread per (row, channel) series (``snsde_natural_coeffs``).
    ``missing=True``: ``x`` may hold NaNs (missing observations); the reference's per-series missing-value branch
    (controldiffeq/interpolate.py:56-153) runs on device (``snsde_natural_coeffs_missing``)."""
    from . import _lib
    x, t = _cuda_args(x, t, "natural_coeffs_cuda")
    B, K, C = x.shape
    if out is None:
        out = torch.empty((B, K - 1, 4 * C), device=x.device, dtype=torch.float32)
    if missing:
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().snsde_natural_coeffs_missing(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(t.data_ptr()), B, K, C,
                                                            ctypes.c_void_p(out.data_ptr()), x.device.index or 0,
                                                            ctypes.c_void_p(stream)))
        return out
    scratch = torch.empty(3 * K, device=x.device, dtype=torch.float32)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.load().snsde_natural_coeffs(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(t.data_ptr()), B, K, C,
                                                ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(scratch.data_ptr()),
                                                x.device.index or 0, ctypes.c_void_p(stream)))
    return out


def hermite_coeffs_cuda(x, t, out=None, fill_missing=False):
    """Fused CUDA version of :func:`hermite_backward_difference_coeffs` for a CUDA tensor ``x [B, K, C]`` (one HBM
    pass through the C ABI, ``snsde_hermite_coeffs``).  ``fill_missing=True`` first fills NaNs on device the way
    torchcde does (:func:`fill_missing_cuda`)."""
    from . import _lib
    x, t = _cuda_args(x, t, "hermite_coeffs_cuda")
    if fill_missing:
        x = fill_missing_cuda(x, t)
    B, K, C = x.shape
    if out is None:
        out = torch.empty((B, K - 1, 4 * C), device=x.device, dtype=torch.float32)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.load().snsde_hermite_coeffs(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(t.data_ptr()), B, K, C,
                                                ctypes.c_void_p(out.data_ptr()), x.device.index or 0,
                                                ctypes.c_void_p(stream)))
    return out

"""Oracle (test infrastructure): piecewise-cubic control path X(t).

Restates
* torchcde 0.2.5 ``CubicSpline._interpret_t/evaluate`` (third-party, absent from
  /root/reference; call sites: benchmark_classification/models_sde/neuralsde.py:184,296).
  The same read formula exists in-tree at
  benchmark_classification/controldiffeq/interpolate.py:261-274.
* torchcde 0.2.5 ``hermite_cubic_coefficients_with_backward_differences``
  (call sites: benchmark_classification/datasets/common.py:82-84,
  tests/test_neuralsde_core_alignment.py:64).
* the in-tree natural cubic spline coefficient builder
  benchmark_classification/controldiffeq/interpolate.py:7-53 (tridiagonal system
  solved by misc.py:13-66), restated as a vectorised Thomas sweep.

Coefficient packing everywhere: last dim = cat(a, b, two_c, three_d), each C wide.
"""
import torch


def _fill_missing_linear(x, t):
    """NaN fill as torchcde's linear_interpolation_coeffs does for the Hermite
    builder: linear interpolation between observed neighbours, forward fill at the
    tail, first valid value at the head.  Data prep only (host side)."""
    if not torch.isnan(x).any():
        return x
    x = x.clone()
    flat = x.reshape(-1, x.shape[-2], x.shape[-1])
    for b in range(flat.shape[0]):
        for c in range(flat.shape[-1]):
            v = flat[b, :, c]
            ok = ~torch.isnan(v)
            if ok.all() or not ok.any():
                continue
            idx = torch.nonzero(ok).flatten()
            first, last = idx[0].item(), idx[-1].item()
            v[:first] = v[first]
            v[last + 1:] = v[last]
            for lo, hi in zip(idx[:-1].tolist(), idx[1:].tolist()):
                if hi - lo > 1:
                    w = (t[lo + 1:hi] - t[lo]) / (t[hi] - t[lo])
                    v[lo + 1:hi] = v[lo] + w * (v[hi] - v[lo])
    return x


def hermite_cubic_coefficients_with_backward_differences(x, t=None):
    """x: [..., K, C] -> coeffs [..., K-1, 4C].

    Per interval k with width h: slope m1 = (x[k+1]-x[k])/h, m0 = slope of the
    previous interval (the first interval reuses its own slope);
    a = x[k], b = m0, two_c = 2(3(dx/h - m0) - m1 + m0)/h,
    three_d = (m1 - m0)/h^2 - two_c/h.
    """
    if t is None:
        t = torch.linspace(0, x.size(-2) - 1, x.size(-2), dtype=x.dtype, device=x.device)
    x = _fill_missing_linear(x, t)
    h = (t[1:] - t[:-1]).unsqueeze(-1)
    x_prev = x[..., :-1, :]
    x_next = x[..., 1:, :]
    slope = (x_next - x_prev) / h
    slope_prev = torch.cat((slope[..., :1, :], slope[..., :-1, :]), dim=-2)
    a = x_prev
    b = slope_prev
    two_c = 2 * (3 * ((x_next - x_prev) / h - b) - slope + slope_prev) / h
    three_d = (1 / h ** 2) * (slope - b) - two_c / h
    return torch.cat([a, b, two_c, three_d], dim=-1)


def natural_cubic_spline_coeffs(t, x):
    """x: [..., K, C] (no NaNs) -> (a, b, two_c, three_d), each [..., K-1, C].

    Natural cubic spline: knot derivatives k_i solve the tridiagonal system
      (2/h_0) k_0 + (1/h_0) k_1                         = 3 dx_0/h_0^2
      (1/h_{i-1}) k_{i-1} + 2(1/h_{i-1}+1/h_i) k_i + (1/h_i) k_{i+1}
                                                        = 3(dx_{i-1}/h_{i-1}^2 + dx_i/h_i^2)
      (1/h_{K-2}) k_{K-2} + (2/h_{K-2}) k_{K-1}         = 3 dx_{K-2}/h_{K-2}^2
    (interpolate.py:22-42).  Operation order of the forward/backward sweep follows
    misc.py:52-64 so fp32 results match the reference to rounding.
    """
    if torch.isnan(x).any():
        return _natural_with_missing_values(t, x)
    K = x.size(-2)
    path = x.transpose(-1, -2)                       # [..., C, K]
    if K == 2:
        a = path[..., :1]
        b = (path[..., 1:] - path[..., :1]) / (t[1:] - t[:1])
        z = torch.zeros_like(a)
        out = (a, b, z, z.clone())
    else:
        h = t[1:] - t[:-1]
        r = h.reciprocal()
        r2 = r ** 2
        three_dx = 3 * (path[..., 1:] - path[..., :-1])
        six_dx = 2 * three_dx
        scaled = three_dx * r2
        diag = torch.empty(K, dtype=x.dtype, device=x.device)
        diag[:-1] = r
        diag[-1] = 0
        diag[1:] += r
        diag *= 2
        rhs = torch.empty_like(path)
        rhs[..., :-1] = scaled
        rhs[..., -1] = 0
        rhs[..., 1:] += scaled
        # Thomas algorithm, upper = lower = r
        nd = [diag[0]]
        nb = [rhs[..., 0]]
        for i in range(1, K):
            w = r[i - 1] / nd[i - 1]
            nd.append(diag[i] - w * r[i - 1])
            nb.append(rhs[..., i] - w * nb[i - 1])
        ks = [None] * K
        ks[K - 1] = nb[K - 1] / nd[K - 1]
        for i in range(K - 2, -1, -1):
            ks[i] = (nb[i] - r[i] * ks[i + 1]) / nd[i]
        kd = torch.stack(ks, dim=-1)
        a = path[..., :-1]
        b = kd[..., :-1]
        two_c = (six_dx * r - 4 * kd[..., :-1] - 2 * kd[..., 1:]) * r
        three_d = (-six_dx * r + 3 * (kd[..., :-1] + kd[..., 1:])) * r2
        out = (a, b, two_c, three_d)
    return tuple(o.transpose(-1, -2) for o in out)


def _natural_with_missing_values(t, x):
    """Missing-value branch of the in-tree builder (interpolate.py:56-153), one scalar series at a time as the reference
    does: an all-NaN series gives zero coefficients; a NaN at either end is imputed with the nearest observation; the
    natural spline is built on the OBSERVED knots only and every original interval [t_i, t_{i+1}) then receives the
    piece it lies in, re-centred at t_i:  with offset = t_obs - t_i <= 0
        a_i = a + ((two_c/2 - three_d offset/3) offset - b) offset,   b_i = b + (three_d offset - two_c) offset,
        two_c_i = two_c - 2 three_d offset,                           three_d_i = three_d."""
    lead = x.shape[:-2]
    K, C = x.shape[-2:]
    flat = x.reshape(-1, K, C)
    outs = [torch.zeros(flat.shape[0], K - 1, C, dtype=x.dtype) for _ in range(4)]
    for n in range(flat.shape[0]):
        for c in range(C):
            path = flat[n, :, c].clone()
            obs = ~torch.isnan(path)
            if not obs.any():
                continue
            vals = path[obs]
            if torch.isnan(path[0]):
                path[0] = vals[0]
            if torch.isnan(path[-1]):
                path[-1] = vals[-1]
            obs = ~torch.isnan(path)
            t_o, p_o = t[obs], path[obs]
            a, b, c2, d3 = natural_cubic_spline_coeffs(t_o, p_o.unsqueeze(-1))
            a, b, c2, d3 = a[:, 0], b[:, 0], c2[:, 0], d3[:, 0]
            piece = torch.bucketize(t[:-1], t_o, right=True) - 1          # observed interval containing t_i
            off = t_o[piece] - t[:-1]
            a_inner = (0.5 * c2[piece] - d3[piece] * off / 3) * off
            outs[0][n, :, c] = a[piece] + (a_inner - b[piece]) * off
            outs[1][n, :, c] = b[piece] + (d3[piece] * off - c2[piece]) * off
            outs[2][n, :, c] = c2[piece] - 2 * d3[piece] * off
            outs[3][n, :, c] = d3[piece]
    return tuple(o.reshape(*lead, K - 1, C) for o in outs)


class CubicSpline:
    """X(t) from packed coefficients ``[..., K-1, 4C]`` and knots ``[K]``.

    ``index = clamp(bucketize(t, knots) - 1, 0, K-2)`` (right=False: at an interior
    knot the LEFT interval is used with frac = its full width; t <= knots[0] -> 0);
    ``frac = t - knots[index]``;
    ``X = a + (b + (two_c/2 + three_d*frac/3) * frac) * frac``.
    """

    def __init__(self, coeffs, t):
        if isinstance(coeffs, (tuple, list)):
            coeffs = torch.cat(list(coeffs), dim=-1)
        C = coeffs.size(-1) // 4
        self._t = t
        self._a, self._b, self._two_c, self._three_d = (
            coeffs[..., :C], coeffs[..., C:2 * C], coeffs[..., 2 * C:3 * C], coeffs[..., 3 * C:])

    def interpret_t(self, t):
        t = torch.as_tensor(t, dtype=self._b.dtype, device=self._b.device)
        maxlen = self._b.size(-2) - 1
        index = torch.bucketize(t.detach(), self._t.detach()).sub(1).clamp(0, maxlen)
        frac = t - self._t[index]
        return frac, index

    def evaluate(self, t):
        frac, index = self.interpret_t(t)
        frac = frac.unsqueeze(-1)
        inner = 0.5 * self._two_c[..., index, :] + self._three_d[..., index, :] * frac / 3
        inner = self._b[..., index, :] + inner * frac
        return self._a[..., index, :] + inner * frac

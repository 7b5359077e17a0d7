"""Oracle (test infrastructure): torchsde 0.2.5 fixed-step solve, restated.

torchsde is an un-vendored pip dependency of the reference (environment.yml:20,
``torchsde==0.2.5``) and is absent here; **parity unpinned** (see oracle/__init__.py).
Anchors: reference call sites benchmark_classification/models_sde/neuralsde.py:78-82
(``torchsde.sdeint(sde=func, y0=z0, ts=ts, dt=dt, method=...)``) and
tutorial notebooks cell 7 (``dt=0.05, method='euler'``).

Restated semantics (torchsde/_core/base_solver.py ``integrate``,
methods/euler.py, methods/milstein.py, _core/interp.py):

* time scalars are 0-d tensors of ``ts.dtype`` (float32): ``curr_t = ts[0]``;
  ``next_t = min(curr_t + dt, ts[-1])``; loop ``while curr_t < out_t``;
* output at ``out_t`` = ``(t1-t)/(t1-t0)*y0 + (t-t0)/(t1-t0)*y1`` of the two states
  bracketing it (== y1 exactly when a step lands on ``out_t``);
* Euler (Ito):      ``y1 = y0 + f*dt + g*dW``        (diagonal noise: elementwise)
* Milstein (Ito, diagonal, derivative-based):
                    ``y1 = y0 + f*dt + g*dW + 0.5*vjp_y(g; g*(dW^2-dt))``
* SRK (Ito, diagonal noise -> ``SRK.diagonal_or_scalar_step``, strong order 1.5, tableau
  ``tableaus/srid2.py`` = Roessler 2010 SRI2 for diagonal noise): the torch-ists wrapper's
  default method (torch-ists/torch_ists/diff_module/NSDE/nsde_model.py:63-74, ``default_method='srk'``).
  It queries ``bm(t0, t1, return_U=True)`` -> ``(I_k, I_k0)`` where ``I_k0 = U`` is the space-time
  Levy integral ``h (W/2 + Hst)``, ``Hst ~ N(0, h/12)`` (brownian_interval.py ``_H_to_U``).
* before the loop ``sdeint`` evaluates ``f`` and ``g`` once at ``ts[0]`` (shape checks).
"""
import torch


class BrownianTable:
    """Explicit-increment Brownian source: ``bm(t0, t1)`` returns row k of ``dW[S,B,H]`` on
    the k-th call.  This is the ``bm=`` object the reference could receive through its
    ``**kwargs`` pass-through (neuralsde.py:84,105,82); parity is defined on identical
    increments because ``BrownianInterval`` streams are not reproducible on device."""

    def __init__(self, dW, check_times=None, dU=None):
        self.dW = dW
        self.dU = dU                        # space-time Levy integrals U[S,B,H] (method 'srk')
        self.k = 0
        self.check_times = check_times      # optional [(t0,t1)] list to assert the call pattern

    def __call__(self, t0, t1, return_U=False):
        if self.check_times is not None:
            e0, e1 = self.check_times[self.k]
            assert float(t0) == e0 and float(t1) == e1, (self.k, float(t0), float(t1), e0, e1)
        w = self.dW[self.k]
        u = self.dU[self.k] if return_U else None
        self.k += 1
        return (w, u) if return_U else w


def _lerp(t0, y0, t1, y1, t):
    assert t0 <= t <= t1
    return (t1 - t) / (t1 - t0) * y0 + (t - t0) / (t1 - t0) * y1


def euler_step(sde, bm, t0, t1, y0):
    dt = t1 - t0
    dW = bm(t0, t1)
    f = sde.f(t0, y0)
    g = sde.g(t0, y0)
    return y0 + f * dt + g * dW


def milstein_step(sde, bm, t0, t1, y0):
    """torchsde 0.2.5 ``Milstein.step`` (Ito, derivative-based) on ``ForwardSDE.g_prod_and_gdg_prod_diagonal``: the vjp is
    taken with ``create_graph=torch.is_grad_enabled()`` and ``grad_outputs = g * v`` inside the graph, so under autograd
    the gradient flows through both factors of ``0.5 * (g v) dg/dy`` (second derivative of g)."""
    dt = t1 - t0
    dW = bm(t0, t1)
    v = dW ** 2 - dt
    f = sde.f(t0, y0)
    requires_grad = torch.is_grad_enabled()
    with torch.enable_grad():
        y = y0 if y0.requires_grad else y0.detach().requires_grad_(True)
        g = sde.g(t0, y)
        # torchsde misc.vjp: an output outside the graph (a g built from buffers only, e.g. LatentSDE.g_aug) is made a
        # leaf first, so the vjp is "unused" -> zeros, not an error
        gdg = None
        if g.requires_grad:
            (gdg,) = torch.autograd.grad(g, y, grad_outputs=g * v, create_graph=requires_grad, allow_unused=True)
    if gdg is None:
        gdg = torch.zeros_like(y0)
    if not requires_grad:
        g, gdg = g.detach(), gdg.detach()
    return y0 + f * dt + g * dW + 0.5 * gdg


class SRID2:
    """torchsde/_core/methods/tableaus/srid2.py (Roessler 2010, SRI2 for diagonal noise)."""
    STAGES = 4
    C0 = (0, 1, 1 / 2, 0)
    C1 = (0, 1 / 4, 1, 1 / 4)
    A0 = ((), (1,), (1 / 4, 1 / 4), (0, 0, 0))
    A1 = ((), (1 / 4,), (1, 0), (0, 0, 1 / 4))
    B0 = ((), (0,), (1, 1 / 2), (0, 0, 0))
    B1 = ((), (-1 / 2,), (1, 0), (2, -1, 1 / 2))
    alpha = (1 / 6, 1 / 6, 2 / 3, 0)
    beta1 = (-1, 4 / 3, 2 / 3, 0)
    beta2 = (1, -4 / 3, 1 / 3, 0)
    beta3 = (2, -4 / 3, -2 / 3, 0)
    beta4 = (-2, 5 / 3, -2 / 3, 1)


def srk_step(sde, bm, t0, t1, y0):
    """torchsde 0.2.5 ``SRK.diagonal_or_scalar_step`` (methods/srk.py), statement for statement - including
    its re-evaluation of f and g of the earlier stages inside every stage."""
    tab = SRID2
    dt = t1 - t0
    rdt = 1 / dt
    sqrt_dt = dt.sqrt() if torch.is_tensor(dt) else dt ** 0.5
    I_k, I_k0 = bm(t0, t1, return_U=True)
    I_kk = (I_k ** 2 - dt) * (1 / 2)
    I_kkk = (I_k ** 3 - 3 * dt * I_k) * (1 / 6)
    y1 = y0
    H0, H1 = [], []
    for s in range(tab.STAGES):
        H0s, H1s = y0, y0
        for j in range(s):
            f = sde.f(t0 + tab.C0[j] * dt, H0[j])
            g = sde.g(t0 + tab.C1[j] * dt, H1[j])
            H0s = H0s + tab.A0[s][j] * f * dt + tab.B0[s][j] * g * I_k0 * rdt
            H1s = H1s + tab.A1[s][j] * f * dt + tab.B1[s][j] * g * sqrt_dt
        H0.append(H0s)
        H1.append(H1s)
        f = sde.f(t0 + tab.C0[s] * dt, H0s)
        g_weight = (tab.beta1[s] * I_k + tab.beta2[s] * I_kk / sqrt_dt
                    + tab.beta3[s] * I_k0 * rdt + tab.beta4[s] * I_kkk * rdt)
        g = sde.g(t0 + tab.C1[s] * dt, H1s)
        y1 = y1 + tab.alpha[s] * f * dt + g_weight * g
    return y1


_STEPPERS = {"euler": euler_step, "milstein": milstein_step, "srk": srk_step}


def step_times(ts, dt):
    """The (t0, t1) float pairs the fixed-step loop visits, in ``ts.dtype`` arithmetic."""
    out = []
    curr_t = ts[0]
    for out_t in ts[1:]:
        while curr_t < out_t:
            next_t = min(curr_t + dt, ts[-1])
            out.append((float(curr_t), float(next_t)))
            curr_t = next_t
    return out


class _Renamed:
    """torchsde's ``names={'drift': ..., 'diffusion': ...}`` (``sdeint`` wraps the SDE so that ``f``/``g`` resolve to the
    named methods; used by the reference at torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:139)."""

    def __init__(self, sde, names):
        self.f = getattr(sde, names.get("drift", "f"))
        self.g = getattr(sde, names.get("diffusion", "g"))
        self.noise_type = getattr(sde, "noise_type", "diagonal")
        self.sde_type = getattr(sde, "sde_type", "ito")


def sdeint_with_grad(sde, y0, ts, dt, bm, method="euler", options=None, names=None, **unused):
    """``sdeint`` with autograd left on: what the reference does when it trains through ``torchsde.sdeint``
    (benchmark_classification/common_sde.py:156-162) - the oracle for the engine's backward pass."""
    return _integrate(_Renamed(sde, names) if names else sde, y0, ts, dt, bm, method)


@torch.no_grad()
def sdeint(sde, y0, ts, dt, bm, method="euler", options=None, names=None, **unused):
    """Returns ``[len(ts), B, H]`` like ``torchsde.sdeint``.  ``bm`` is mandatory here."""
    return _integrate(_Renamed(sde, names) if names else sde, y0, ts, dt, bm, method)


def _integrate(sde, y0, ts, dt, bm, method):
    if getattr(sde, "noise_type", "diagonal") != "diagonal" or getattr(sde, "sde_type", "ito") != "ito":
        raise ValueError("oracle supports Ito SDEs with diagonal noise only")
    if method not in _STEPPERS:
        raise ValueError(f"oracle: unsupported method {method!r}")
    step = _STEPPERS[method]
    sde.f(ts[0], y0), sde.g(ts[0], y0)            # shape-check evaluation torchsde performs
    prev_t = curr_t = ts[0]
    prev_y = curr_y = y0
    ys = [y0]
    for out_t in ts[1:]:
        while curr_t < out_t:
            next_t = min(curr_t + dt, ts[-1])
            prev_t, prev_y = curr_t, curr_y
            curr_y = step(sde, bm, curr_t, next_t, curr_y)
            curr_t = next_t
        ys.append(_lerp(prev_t, prev_y, curr_t, curr_y, out_t))
    return torch.stack(ys, dim=0)


def solver_dt(times):
    """``dt = max(min(diff(times)), 1e-3)`` - neuralsde.py:32-33."""
    return max((times[1:] - times[:-1]).min().item(), 1e-3)

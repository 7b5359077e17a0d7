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
* before the loop ``sdeint`` evaluates ``f`` and ``g`` once at ``ts[0]`` (shape checks).
"""
import torch


class BrownianTable:
    """Explicit-increment Brownian source: ``bm(t0, t1)`` returns row k of ``dW[S,B,H]`` on
    the k-th call.  This is the ``bm=`` object the reference could receive through its
    ``**kwargs`` pass-through (neuralsde.py:84,105,82); parity is defined on identical
    increments because ``BrownianInterval`` streams are not reproducible on device."""

    def __init__(self, dW, check_times=None):
        self.dW = dW
        self.k = 0
        self.check_times = check_times      # optional [(t0,t1)] list to assert the call pattern

    def __call__(self, t0, t1):
        if self.check_times is not None:
            e0, e1 = self.check_times[self.k]
            assert float(t0) == e0 and float(t1) == e1, (self.k, float(t0), float(t1), e0, e1)
        w = self.dW[self.k]
        self.k += 1
        return w


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
    dt = t1 - t0
    dW = bm(t0, t1)
    v = dW ** 2 - dt
    f = sde.f(t0, y0)
    g_prod = sde.g(t0, y0) * dW
    with torch.enable_grad():
        yr = y0.detach().requires_grad_(True)
        g = sde.g(t0, yr)
        (gdg,) = torch.autograd.grad(g, yr, grad_outputs=g.detach() * v, allow_unused=True)
    if gdg is None:
        gdg = torch.zeros_like(y0)
    return y0 + f * dt + g_prod + 0.5 * gdg


_STEPPERS = {"euler": euler_step, "milstein": milstein_step}


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


@torch.no_grad()
def sdeint(sde, y0, ts, dt, bm, method="euler", options=None, **unused):
    """Returns ``[len(ts), B, H]`` like ``torchsde.sdeint``.  ``bm`` is mandatory here."""
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

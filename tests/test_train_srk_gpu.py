"""GPU parity tests of round 2: the backward pass, method 'srk' (8 f2), the three reference
wrappers through patch() against goldens minted from the reference's own classes, teacher-forced one-step
parity at the BASELINE shapes, and the boundary hygiene (device guard, range-flag polling).

Tolerances: forward parity 1e-4 of the tensor scale (north star); gradients 1e-4 of each gradient tensor's own
max norm (floor 1e-6) against autograd through the oracle on identical increments.
"""
import copy

import numpy as np
import pytest
import torch

import snsde_b200
from oracle import solver, spline, vector_field, wrapper

from test_engine_gpu import close, make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def grad_close(got, want, name, rtol=1e-4):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    scale = max(float(want.abs().max()), 1e-6)
    err = float((got - want).abs().max())
    assert err <= rtol * scale, f"{name}: grad max abs err {err:.3e} > {rtol:g} * {scale:.3g}"


def euler_with_grad(m, y0, ts, dt, dW):
    """The oracle's fixed-step Euler loop with autograd enabled (oracle.solver.sdeint itself is @no_grad)."""
    return solver.sdeint_with_grad(m, y0, ts, dt, solver.BrownianTable(dW))


BWD_CASES = [
    # io, no, H, C, L, B, K      the five named models first (common_sde.py:303-342), then option coverage
    (4, 17, 32, 5, 1, 9, 7), (6, 17, 32, 5, 2, 8, 6), (2, 16, 32, 4, 1, 8, 6), (1, 18, 32, 3, 2, 7, 6), (1, 0, 32, 3, 1, 8, 5),
    (3, 18, 64, 3, 1, 12, 6), (4, 17, 128, 35, 1, 16, 9), (0, 5, 16, 4, 2, 8, 5), (5, 6, 32, 3, 1, 8, 5), (3, 2, 32, 3, 1, 8, 5),
    (2, 13, 32, 4, 1, 8, 5), (4, 15, 32, 4, 1, 8, 5), (1, 14, 32, 3, 1, 8, 5), (3, 19, 32, 3, 1, 8, 5), (1, 8, 16, 3, 1, 8, 5),
    (5, 9, 16, 3, 1, 8, 5), (3, 10, 16, 3, 1, 8, 5), (1, 11, 16, 3, 1, 8, 5), (2, 12, 16, 3, 1, 8, 5), (6, 3, 16, 3, 1, 8, 5),
]


@pytest.mark.parametrize("io,no,H,C,L,B,K", BWD_CASES)
def test_backward_matches_autograd_through_the_oracle(io, no, H, C, L, B, K, dev):
    """dL/dz0 and dL/d(every parameter) of L = sum(w * sdeint(...)) on identical increments."""
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=100 + io * 20 + no, spacing=0.5)
    if no == 7:
        y0 = y0.abs() + 0.5
    dt = 0.5
    S = K - 1
    g = torch.Generator().manual_seed(7)
    dW = torch.randn(S, B, H, generator=g) * dt ** 0.5
    ts = torch.cat([times[:1], times[2:3], (times[2:3] + times[3:4]) / 2, times[-1:]])     # a knot, a mid-step lerp, the end
    w = torch.randn(len(ts), B, H, generator=g)
    # oracle (fp64 autograd)
    mo = copy.deepcopy(m).double()
    mo.set_X(coeffs.double(), times.double())
    y0o = y0.double().requires_grad_(True)
    zo = euler_with_grad(mo, y0o, ts.double(), dt, dW.double())
    (zo * w.double()).sum().backward()
    # engine
    mg = copy.deepcopy(m).to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0g = y0.to(dev).requires_grad_(True)
    zg = snsde_b200.sdeint(mg, y0g, ts.to(dev), dt=dt, method="euler", bm=snsde_b200.BrownianIncrements(dW.to(dev)),
                           precision="fp32")
    assert zg.requires_grad
    close(zg, zo.float())
    (zg * w.to(dev)).sum().backward()
    grad_close(y0g.grad, y0o.grad, "y0")
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        want = named_o[name].grad
        if want is None:                      # parameter does not influence the solve (dead initial_network, theta at noise 0)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        grad_close(p.grad, want, name)


MILSTEIN_BWD_CASES = [
    # io, no, H, C, L, B, K     state-dependent elementwise diffusions: the Milstein term carries second derivatives of g
    (6, 17, 32, 5, 1, 9, 7), (6, 17, 64, 6, 2, 8, 6), (4, 13, 32, 4, 1, 8, 6), (5, 6, 32, 3, 1, 8, 5), (6, 3, 16, 3, 1, 8, 5),
    (1, 8, 16, 3, 1, 8, 5), (5, 9, 16, 3, 1, 8, 5), (3, 10, 16, 3, 1, 8, 5), (1, 11, 16, 3, 1, 8, 5), (1, 7, 16, 3, 1, 8, 5),
    # state-independent diffusions: the Milstein term vanishes, the sweep equals Euler's
    (2, 16, 32, 4, 1, 8, 6), (4, 17, 128, 10, 1, 8, 5), (1, 0, 16, 3, 1, 8, 5), (0, 4, 16, 4, 1, 8, 5),
]


@pytest.mark.parametrize("io,no,H,C,L,B,K", MILSTEIN_BWD_CASES)
def test_milstein_backward_matches_autograd_through_the_oracle(io, no, H, C, L, B, K, dev):
    """Method 'milstein' under autograd: torchsde differentiates 0.5 * vjp(g; g (dW^2 - h)) through both factors
    (create_graph); the engine's closed form carries the second derivative of every elementwise diffusion."""
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=300 + io * 20 + no, spacing=0.5)
    if no == 7:
        y0 = y0.abs() + 1.5                    # sqrt(y): keep the short trajectory in the domain
    dt = 0.5
    S = K - 1
    g = torch.Generator().manual_seed(11)
    dW = torch.randn(S, B, H, generator=g) * dt ** 0.5
    if no == 7:
        dW = dW * 0.3
    ts = torch.cat([times[:1], times[2:3], (times[2:3] + times[3:4]) / 2, times[-1:]])
    w = torch.randn(len(ts), B, H, generator=g)
    mo = copy.deepcopy(m).double()
    mo.set_X(coeffs.double(), times.double())
    y0o = y0.double().requires_grad_(True)
    zo = solver.sdeint_with_grad(mo, y0o, ts.double(), dt, solver.BrownianTable(dW.double()), method="milstein")
    (zo * w.double()).sum().backward()
    mg = copy.deepcopy(m).to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0g = y0.to(dev).requires_grad_(True)
    zg = snsde_b200.sdeint(mg, y0g, ts.to(dev), dt=dt, method="milstein", bm=snsde_b200.BrownianIncrements(dW.to(dev)),
                           precision="fp32")
    close(zg, zo.float())
    (zg * w.to(dev)).sum().backward()
    grad_close(y0g.grad, y0o.grad, "y0")
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        want = named_o[name].grad
        if want is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        # theta / sigma: one number summed from S*B*H signed terms that largely cancel - fp32 accumulation error scales with
        # the sum of magnitudes, not with the total (observed 1.0e-4 of the total on (4,13)); every tensor-valued gradient
        # keeps the 1e-4 bar
        grad_close(p.grad, want, name, rtol=5e-4 if want.numel() == 1 else 1e-4)


def test_milstein_backward_refuses_noise_networks(dev):
    m, times, coeffs, y0 = make_problem(3, 18, 4, 16, 3, 1, 5, seed=1)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with pytest.raises(RuntimeError, match="backward"):          # never drops gradients silently
        snsde_b200.sdeint(mg, y0.to(dev).requires_grad_(True), times.to(dev), dt=1.0, method="milstein", seed=1)
    with torch.no_grad():                      # inference is unaffected
        snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, method="milstein", seed=1)


SRK_BWD_CASES = [
    # family, io, no, H, C, L, B, K     the torch-ists default method under autograd: three drift and four diffusion sites per step
    ("benchmark", 4, 17, 32, 5, 1, 9, 7), ("benchmark", 6, 17, 32, 5, 2, 8, 6), ("benchmark", 2, 16, 32, 4, 1, 8, 6),
    ("benchmark", 3, 18, 32, 3, 1, 7, 6), ("benchmark", 1, 19, 32, 3, 2, 7, 6), ("benchmark", 1, 14, 16, 3, 1, 8, 5),
    ("benchmark", 4, 15, 16, 4, 1, 8, 5), ("benchmark", 0, 5, 16, 4, 2, 8, 5), ("benchmark", 5, 6, 32, 3, 1, 8, 5),
    ("benchmark", 6, 3, 16, 3, 1, 8, 5), ("benchmark", 1, 9, 16, 3, 1, 8, 5), ("benchmark", 1, 0, 16, 3, 1, 8, 5),
    ("benchmark", 4, 17, 128, 10, 1, 8, 5), ("benchmark", 3, 18, 64, 3, 1, 6, 5), ("tutorial", 0, 0, 32, 2, 1, 8, 6),
    # enough rows for 4-row groups (the launch shape of real batches), ragged last group
    ("benchmark", 4, 17, 32, 4, 1, 322, 5), ("benchmark", 3, 18, 32, 3, 1, 301, 5), ("benchmark", 6, 17, 128, 6, 1, 90, 4),
]


@pytest.mark.parametrize("family,io,no,H,C,L,B,K", SRK_BWD_CASES)
def test_srk_backward_matches_autograd_through_the_oracle(family, io, no, H, C, L, B, K, dev):
    """Method 'srk' under autograd on the torch-ists grid (linspace knots, sliver last step): dL/dz0 and dL/d(every
    parameter) against fp64 autograd through the oracle's srk_step on identical (dW, U)."""
    m, _, _, y0 = make_problem(io, no, B, H, C, L, K, seed=400 + io * 20 + no, family=family)
    times = torch.linspace(0, 1, K)
    x = (torch.randn(B, K, C, generator=torch.Generator().manual_seed(1)) * 0.2).cumsum(1)
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    dt = solver.solver_dt(times)
    steps = solver.step_times(times, dt)
    S = len(steps)
    g = torch.Generator().manual_seed(13)
    h = torch.tensor([b - a for a, b in steps]).view(S, 1, 1)
    dW = torch.randn(S, B, H, generator=g) * h.sqrt()
    dU = h * (0.5 * dW + torch.randn(S, B, H, generator=g) * (h / 12).sqrt())
    ts = torch.cat([times[:1], times[2:3], (times[2:3] + times[3:4]) / 2, times[-1:]])
    w = torch.randn(len(ts), B, H, generator=g)
    mo = copy.deepcopy(m).double()
    mo.set_X(coeffs.double(), times.double())
    y0o = y0.double().requires_grad_(True)
    zo = solver.sdeint_with_grad(mo, y0o, ts.double(), dt, solver.BrownianTable(dW.double(), dU=dU.double()), method="srk")
    (zo * w.double()).sum().backward()
    mg = copy.deepcopy(m).to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0g = y0.to(dev).requires_grad_(True)
    zg = snsde_b200.sdeint(mg, y0g, ts.to(dev), dt=dt, method="srk", bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev)),
                           precision="fp32")
    assert zg.requires_grad
    close(zg, zo.float())
    (zg * w.to(dev)).sum().backward()
    grad_close(y0g.grad, y0o.grad, "y0")
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        want = named_o[name].grad
        if want is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        grad_close(p.grad, want, name, rtol=5e-4 if want.numel() == 1 else 1e-4)


def test_srk_backward_philox_replay_equals_the_table(dev):
    """In-kernel increments: the reverse sweep regenerates dW and the Levy integrals from the two Philox streams."""
    m, times, coeffs, y0 = make_problem(4, 17, 10, 32, 4, 1, 9, seed=21, spacing=0.25)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    grads = []
    for mode in ("philox", "table"):
        mg.zero_grad(set_to_none=True)
        y0g = y0.to(dev).requires_grad_(True)
        if mode == "philox":
            z = snsde_b200.sdeint(mg, y0g, times.to(dev), dt=0.25, method="srk", seed=5, precision="fp32")
        else:
            plan = snsde_b200.plans_of(mg)[("srk", "fp32", str(dev))]
            dW, dU = snsde_b200.philox_increments(5, plan.step_plan(times, 0.25, times), 10, 32, dev, with_U=True)
            z = snsde_b200.sdeint(mg, y0g, times.to(dev), dt=0.25, method="srk", bm=snsde_b200.BrownianIncrements(dW, dU),
                                  precision="fp32")
        z[-1].pow(2).sum().backward()
        grads.append((y0g.grad.clone(), mg.linear_out.weight.grad.clone(), mg.noise_t[0].weight.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0])                      # same increments, same arithmetic
    for a, b in zip(*grads):                                           # (atomics order the coefficient-network sums)
        grad_close(a, b, "philox replay vs table", rtol=1e-5)


def test_backward_through_fused_final_index_and_philox_replay(dev):
    """Training path of the classification wrapper: per-row final_index capture + in-kernel Philox increments
    (the backward regenerates the same stream) vs the oracle fed the materialised increments."""
    io, no, B, H, C, L, K = 4, 17, 21, 64, 6, 1, 12
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=5)
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(2))
    mg = copy.deepcopy(m).to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0g = y0.to(dev).requires_grad_(True)
    z = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0g, seed=99, precision="fp32")
    loss = (z ** 2).sum()
    loss.backward()
    plan = snsde_b200.plans_of(mg)[("euler", "fp32", str(dev))]
    ts, _ = snsde_b200.final_index_slots(times, fi)
    dW = snsde_b200.philox_increments(99, plan.step_plan(ts, 1.0, times), B, H, dev).cpu()
    mo = copy.deepcopy(m).double()
    mo.set_X(coeffs.double(), times.double())
    y0o = y0.double().requires_grad_(True)
    z_all = euler_with_grad(mo, y0o, times.double(), 1.0, dW.double())           # every knot
    zo = z_all[fi, torch.arange(B)]
    close(z, zo.float())
    (zo ** 2).sum().backward()
    grad_close(y0g.grad, y0o.grad, "y0")
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        if named_o[name].grad is not None:
            grad_close(p.grad, named_o[name].grad, name)


def test_backward_c2_shape_slice_and_tensor_core_forward(dev):
    """c2 model/shape at B=64: forward states from the tcgen05 kernel, reverse sweep in fp32."""
    io, no, B, H, C, L, K = 4, 17, 64, 128, 35, 1, 41
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=17)
    y0 = y0 * 0.2
    fi = torch.randint(2, K, (B,), generator=torch.Generator().manual_seed(3))
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(4))
    mg = copy.deepcopy(m).to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0g = y0.to(dev).requires_grad_(True)
    z = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0g, bm=snsde_b200.BrownianIncrements(dW.to(dev)))
    assert snsde_b200.plans_of(mg)[("euler", "auto", str(dev))].kernel == "tcgen05"
    head = torch.randn(H, generator=torch.Generator().manual_seed(5))
    (z @ head.to(dev)).sum().backward()
    mo = copy.deepcopy(m).double()
    mo.set_X(coeffs.double(), times.double())
    y0o = y0.double().requires_grad_(True)
    zo = euler_with_grad(mo, y0o, times.double(), 1.0, dW.double())[fi, torch.arange(B)]
    close(z, zo.float())
    (zo @ head.double()).sum().backward()
    grad_close(y0g.grad, y0o.grad, "y0")
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        if named_o[name].grad is not None:
            grad_close(p.grad, named_o[name].grad, name)


def test_backward_tutorial_family_lipswish(dev):
    B, H, C, L, K = 10, 32, 2, 1, 8
    m, times, coeffs, y0 = make_problem(0, 0, B, H, C, L, K, seed=9, family="tutorial", spacing=0.125)
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(1)) * 0.125 ** 0.5
    mg = copy.deepcopy(m).to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0g = y0.to(dev).requires_grad_(True)
    z = snsde_b200.sdeint(mg, y0g, times.to(dev), dt=0.125, bm=snsde_b200.BrownianIncrements(dW.to(dev)))
    z[-1].pow(2).sum().backward()
    mo = copy.deepcopy(m).double()
    mo.set_X(coeffs.double(), times.double())
    y0o = y0.double().requires_grad_(True)
    zo = euler_with_grad(mo, y0o, times.double(), 0.125, dW.double())
    close(z, zo.float())
    zo[-1].pow(2).sum().backward()
    grad_close(y0g.grad, y0o.grad, "y0")
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        grad_close(p.grad, named_o[name].grad, name)


def test_training_step_through_the_patched_reference_style_wrapper(dev):
    """`pred = model(...); loss.backward(); optimizer.step()` as the reference harness does (common_sde.py:156-162):
    gradients reach the head, initial_network (through z0) and func; the loss goes down; a deep copy still works."""
    class RefStyleNeuralSDE(torch.nn.Module):      # shape of reference NeuralSDE (neuralsde.py:51-120), real torchcde absent
        def __init__(self, func, C, H, out):
            super().__init__()
            self.func, self.initial = func, True
            self.initial_network = torch.nn.Linear(C, H)
            self.linear = torch.nn.Linear(H, out)

        def _prepare_initial_state(self, times, z0):
            return self.initial_network(self.func.X.evaluate(times[0])) if z0 is None else z0

        def _solve_sde_path(self, times, ts, z0, kwargs):
            raise AssertionError("must be replaced by patch()")

        def forward(self, times, coeffs, final_index, z0=None, stream=False, **kwargs):
            raise AssertionError("must be replaced by patch()")

    B, H, C, L, K = 32, 32, 4, 1, 10
    func, times, coeffs, _ = make_problem(4, 17, B, H, C, L, K, seed=3)
    model = snsde_b200.patch(RefStyleNeuralSDE(func, C, H, 1).to(dev))
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(1)).to(dev)
    target = torch.randn(B, 1, generator=torch.Generator().manual_seed(2)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)      # the reference's lr (sepsis-sde.py:38)
    losses = []
    for it in range(12):
        opt.zero_grad()
        pred = model(times.to(dev), [coeffs.to(dev)], fi, seed=5)
        loss = torch.nn.functional.mse_loss(pred, target)
        loss.backward()
        if it == 0:
            for name, p in model.named_parameters():
                if name.startswith("func.initial_network") or p.grad is None:
                    continue
                assert torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, name
            assert model.initial_network.weight.grad is not None and model.func.linear_out.weight.grad is not None
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    best = copy.deepcopy(model)                        # common_sde.py:181
    with torch.no_grad():
        a = model(times.to(dev), [coeffs.to(dev)], fi, seed=5)
        b = best(times.to(dev), [coeffs.to(dev)], fi, seed=5)
    assert torch.equal(a, b)


# ---- method 'srk' --------------------------------------------------------------------------------
SRK_CASES = [(4, 17, 32, 5, 1, 9), (6, 17, 64, 6, 2, 8), (2, 16, 32, 4, 1, 8), (3, 18, 32, 3, 1, 7), (1, 19, 32, 3, 2, 8),
             (0, 5, 16, 4, 1, 8), (5, 2, 16, 3, 1, 8), (1, 11, 16, 3, 1, 8), (4, 13, 128, 10, 1, 8), (1, 14, 32, 3, 1, 8)]


@pytest.mark.parametrize("io,no,H,C,L,B", SRK_CASES)
def test_srk_matches_the_oracle_on_the_torch_ists_grid(io, no, H, C, L, B, dev):
    """torch-ists grid: times = linspace(0, 1, K), dt = min diff -> a sliver last step; outputs at every knot."""
    K = 23
    m, _, _, y0 = make_problem(io, no, B, H, C, L, K, seed=200 + io + no)
    times = torch.linspace(0, 1, K)
    x = (torch.randn(B, K, C, generator=torch.Generator().manual_seed(1)) * 0.2).cumsum(1)
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    dt = solver.solver_dt(times)
    steps = solver.step_times(times, dt)
    S = len(steps)
    g = torch.Generator().manual_seed(2)
    h = torch.tensor([b - a for a, b in steps]).view(S, 1, 1)
    dW = torch.randn(S, B, H, generator=g) * h.sqrt()
    dU = h * (0.5 * dW + torch.randn(S, B, H, generator=g) * (h / 12).sqrt())
    m.set_X(coeffs, times)
    want = solver.sdeint(m, y0, times, dt, solver.BrownianTable(dW, dU=dU), method="srk")
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        got = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt, method="srk",
                                bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev)))
    assert snsde_b200.plans_of(mg)[("srk", "auto", str(dev))].kernel == "fma_fp32"
    close(got, want)


def test_srk_in_kernel_levy_areas_equal_their_replay_and_have_the_right_moments(dev):
    io, no, B, H, C, L, K = 4, 17, 64, 32, 4, 1, 33
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=8, spacing=0.25)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        a = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=0.25, method="srk", seed=31)
        plan = snsde_b200.plans_of(mg)[("srk", "auto", str(dev))]
        sp = plan.step_plan(times, 0.25, times)
        dW, dU = snsde_b200.philox_increments(31, sp, B, H, dev, with_U=True)
        b = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=0.25, method="srk", bm=snsde_b200.BrownianIncrements(dW, dU))
    assert torch.equal(a, b)
    # the W stream is the Euler kernels' stream; U = h (W/2 + Hst): Var U = h^3/3, Cov(U, W) = h^2/2
    dW1 = snsde_b200.philox_increments(31, sp, B, H, dev)
    assert torch.equal(dW, dW1)
    h = 0.25
    W, U = dW.double().flatten().cpu(), dU.double().flatten().cpu()
    assert abs(float(U.var()) / (h ** 3 / 3) - 1) < 0.03 and abs(float((U * W).mean()) / (h ** 2 / 2) - 1) < 0.03
    Hst = U / h - 0.5 * W
    assert abs(float(Hst.var()) / (h / 12) - 1) < 0.03 and abs(float((Hst * W).mean())) < 0.02 * h


# ---- the three reference wrappers through patch(), against goldens minted from the reference's own classes ----
def _oracle_func(c, dev):
    B, K, C, H, HH, L = c["dims"]
    m = vector_field.DiffusionModel(C, H, HH, L, input_option=c["input_option"], noise_option=c["noise_option"])
    m.load_state_dict(c["state_dict"])
    return m.to(dev)


def test_forecasting_wrapper_golden_through_patch(golden_dir, dev):
    """NeuralSDE_forecasting.forward (benchmark_forecasting/models_sde/neuralsde.py:158-186): natural-spline 4-tuple
    coeffs, z0 = initial_network(X(t0)), every knot solved, the last output_time knots headed (head = Identity)."""
    cases = [c for c in torch.load(golden_dir / "forward_golden.pt") if c["kind"] == "forecasting"]
    assert cases
    for c in cases:
        B, K, C, H, HH, L = c["dims"]

        class Forecasting(torch.nn.Module):          # the reference class's forward signature and attributes
            def __init__(self, func):
                super().__init__()
                self.func, self.initial, self.output_time = func, True, c["output_time"]
                self.initial_network = torch.nn.Linear(C, H)
                self.linear = torch.nn.Identity()

            def _prepare_initial_state(self, times, z0):
                return self.initial_network(self.func.X.evaluate(times[0])) if z0 is None else z0

            def _solve_sde_path(self, times, ts, z0, kwargs):
                raise AssertionError("replaced by patch()")

            def forward(self, times, coeffs, final_index, z0=None, stream=False, **kwargs):
                raise AssertionError("replaced by patch()")

        model = Forecasting(_oracle_func(c, dev))
        model.initial_network.load_state_dict(c["initial_network"])
        model = snsde_b200.patch(model.to(dev))
        assert snsde_b200.wrapper_kind(model) == "forecasting"
        coeffs4 = tuple(t.to(dev) for t in c["coeffs"].chunk(4, dim=-1))
        with torch.no_grad():
            z = model(c["times"].to(dev), coeffs4, None, bm=snsde_b200.BrownianIncrements(c["dW"].to(dev)), precision="fp32")
        assert z.shape == c["z"].shape == (B, c["output_time"], H)
        close(z, c["z"], rtol=1e-5)


def test_torch_ists_wrapper_golden_through_patch(golden_dir, dev):
    """torch-ists NeuralSDE.forward(coeffs, times) -> (head(z), z) (nsde_model.py:76-84), default method 'srk'."""
    cases = [c for c in torch.load(golden_dir / "forward_golden.pt") if c["kind"] == "torch_ists"]
    assert len(cases) == 3
    for c in cases:
        B, K, C, H, HH, L = c["dims"]

        class TorchIsts(torch.nn.Module):            # the reference class's own methods, minus the torchsde call
            def __init__(self, func):
                super().__init__()
                self.func, self.initial = func, True
                self.initial_network = torch.nn.Linear(C, H)
                self.linear = torch.nn.Sequential(torch.nn.Tanh(), torch.nn.Linear(H, H), torch.nn.ReLU(), torch.nn.Linear(H, 2))

            def _prepare_initial_state(self, times):
                return self.initial_network(self.func.X.evaluate(times[0]))

            def _solve_sde_path(self, times, y0, kwargs):
                raise AssertionError("replaced by patch()")

            def forward(self, coeffs, times, **kwargs):
                self.func.set_X(coeffs, times)
                y0 = self._prepare_initial_state(times)
                z = self._solve_sde_path(times, y0, kwargs).permute(1, 0, 2)
                return self.linear(z), z

        model = TorchIsts(_oracle_func(c, dev))
        model.load_state_dict(c["model_state"])
        model = snsde_b200.patch(model.to(dev))
        assert snsde_b200.wrapper_kind(model) == "torch_ists" and "forward" not in model.__dict__
        kw = {} if c["method"] is None else {"method": c["method"]}
        with torch.no_grad():
            pred, z = model(c["coeffs"].to(dev), c["times"].to(dev),
                            bm=snsde_b200.BrownianIncrements(c["dW"].to(dev), c["dU"].to(dev)), precision="fp32", **kw)
        close(z, c["z"], rtol=1e-5)
        close(pred, c["pred"], rtol=1e-5)


# ---- teacher-forced one-step parity at the BASELINE shapes --------------------------
TF_CASES = [  # name, io, no, method, B, H, C, S, precision
    ("c3", 6, 17, "milstein", 2048, 64, 35, 24, "tc"),
    ("c4", 3, 18, "euler", 1024, 128, 21, 16, "tc"),
    ("c5", 4, 17, "euler", 256, 256, 14, 16, "tc"),
]


@pytest.mark.parametrize("name,io,no,method,B,H,C,S,precision", TF_CASES)
def test_teacher_forced_single_steps_on_the_tensor_core_kernels(name, io, no, method, B, H, C, S, precision, dev):
    """Chaos-free strict gate for the c3/c4/c5 models on the tcgen05 kernels: every solver step is restarted from the
    ORACLE's state (one launch per step, full batch shape), so the error is the one-step error of the kernel, not the
    growth of a 200-step ill-conditioned trajectory.  Gate: 1e-5 of the state scale per step."""
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, 1, S + 1, seed=300 + io)
    dW = torch.randn(S, B, H, generator=torch.Generator().manual_seed(6))
    m.set_X(coeffs, times)
    traj = solver.sdeint(m, y0, times, 1.0, solver.BrownianTable(dW), method=method)      # oracle trajectory [S+1,B,H]
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    worst = 0.0
    with torch.no_grad():
        for s in range(S):
            ts = times[s:s + 2].to(dev)
            got = snsde_b200.sdeint(mg, traj[s].to(dev), ts, dt=1.0, method=method, precision=precision,
                                    bm=snsde_b200.BrownianIncrements(dW[s:s + 1].to(dev)))
            worst = max(worst, close(got[1], traj[s + 1], rtol=1e-5))
    kinds = {p.kernel for p in snsde_b200.plans_of(mg).values()}
    assert kinds <= {"tcgen05", "tcgen05_general"}, kinds
    m.to("cpu")
    print(f"{name}: worst one-step rel err {worst:.2e}")


# ---- boundary hygiene ---------------------------------------------------------------------------------
def test_engine_calls_leave_the_current_device_alone(dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    other = torch.device("cuda", 1)
    m, times, coeffs, y0 = make_problem(4, 17, 8, 32, 4, 1, 6, seed=1)
    mg = m.to(other)
    mg.set_X(coeffs.to(other), times.to(other))
    torch.cuda.set_device(0)
    with torch.no_grad():
        snsde_b200.sdeint(mg, y0.to(other), times.to(other), dt=1.0, seed=1)
    assert torch.cuda.current_device() == 0
    import ctypes
    cur = ctypes.c_int(-1)
    ctypes.CDLL("libcudart.so.12").cudaGetDevice(ctypes.byref(cur))
    assert cur.value == 0


def test_saturated_tensor_core_solve_is_reported_on_the_next_call_or_rerun(dev):
    B, H, C, L, K = 16, 64, 4, 1, 6
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=2)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    big = y0.clone()
    big[3, 5] = 1.0e5                                   # beyond fp16 range
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(1))
    bm = lambda: snsde_b200.BrownianIncrements(dW.to(dev))          # noqa: E731
    with torch.no_grad():
        snsde_b200.sdeint(mg, big.to(dev), times.to(dev), dt=1.0, bm=bm())                 # silent at this point...
        torch.cuda.synchronize()
        with pytest.raises(snsde_b200.EngineError, match="fp16 range"):                     # ...raised by the next call
            snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=bm())
        got = snsde_b200.sdeint(mg, big.to(dev), times.to(dev), dt=1.0, bm=bm(), check_range=True)    # rerun on fp32
    m.to("cpu"); m.set_X(coeffs, times)
    want = solver.sdeint(m, big, times, 1.0, solver.BrownianTable(dW))
    close(got, want)
    with torch.no_grad():                               # flag consumed: the plan is usable again
        mg = m.to(dev); mg.set_X(coeffs.to(dev), times.to(dev))
        snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=bm())


# ---- the seam's neighbours as engine kernels ---------------------------------------------
class _RefStyleClassification(torch.nn.Module):      # attributes / signatures of reference NeuralSDE (neuralsde.py:51-120)
    def __init__(self, func, C, H, O):
        super().__init__()
        self.func, self.initial = func, True
        self.initial_network = torch.nn.Linear(C, H)
        self.linear = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.BatchNorm1d(H), torch.nn.ReLU(), torch.nn.Dropout(0.1),
                                          torch.nn.Linear(H, O))

    def _prepare_initial_state(self, times, z0):
        return self.initial_network(self.func.X.evaluate(times[0])) if z0 is None else z0

    def _solve_sde_path(self, times, ts, z0, kwargs):
        raise AssertionError("replaced by patch()")

    def forward(self, times, coeffs, final_index, z0=None, stream=False, **kwargs):
        raise AssertionError("replaced by patch()")


class _RefStyleForecasting(_RefStyleClassification):   # NeuralSDE_forecasting (benchmark_forecasting/...:123-186)
    def __init__(self, func, C, H, O, output_time):
        super().__init__(func, C, H, O)
        self.output_time = output_time
        self.linear = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.ReLU(), torch.nn.Linear(H, O))


def test_full_forward_goldens_with_fused_neighbours(golden_dir, dev, monkeypatch):
    """Eval-mode forward of the classification / forecasting wrappers INCLUDING z0 = initial_network(X(t0)) and the real
    read-out heads (BatchNorm1d running statistics, Dropout), against predictions minted from the reference's own classes.
    The PyTorch modules of the neighbours must not run: they are replaced by raising stubs after patch()."""
    cases = {c["kind"]: c for c in torch.load(golden_dir / "forward_golden.pt") if c["kind"].endswith("_full")}
    assert set(cases) == {"classification_full", "forecasting_full"}
    for kind, c in cases.items():
        B, K, C, H, HH, L = c["dims"]
        func = _oracle_func(c, dev)
        if kind == "classification_full":
            model = _RefStyleClassification(func, C, H, c["out_channels"])
        else:
            model = _RefStyleForecasting(func, C, H, c["out_channels"], c["output_time"])
        model.load_state_dict(c["model_state"])
        model = snsde_b200.patch(model.to(dev)).eval()
        called = []
        model._prepare_initial_state = lambda *a, **k: called.append("z0") or (_ for _ in ()).throw(AssertionError("PyTorch z0 producer ran"))
        for mod in model.linear:
            mod.register_forward_hook(lambda *a: called.append("head"))
        coeffs = [c["coeffs"].to(dev)] if kind == "classification_full" else tuple(t.to(dev) for t in c["coeffs"].chunk(4, -1))
        fi = c["final_index"].to(dev) if kind == "classification_full" else None
        with torch.no_grad():
            pred = model(c["times"].to(dev), coeffs, fi, bm=snsde_b200.BrownianIncrements(c["dW"].to(dev)), precision="fp32")
        assert not called, called
        assert pred.shape == c["pred"].shape
        close(pred, c["pred"], rtol=1e-5)
        # training mode: BatchNorm batch statistics / dropout stay in PyTorch (and gradients flow)
        model.train()
        del model._prepare_initial_state
        pred_t = model(c["times"].to(dev), coeffs, fi, seed=3)
        assert "head" in called and pred_t.requires_grad


def test_neighbour_kernels_alone(dev):
    torch.manual_seed(0)
    B, K, C, H, O = 37, 6, 5, 96, 4
    times = torch.linspace(0.5, 3.0, K)
    x = torch.randn(B, K, C).cumsum(1)
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    lin = torch.nn.Linear(C, H)
    want = lin(spline.CubicSpline(coeffs, times).evaluate(times[0]))
    got = snsde_b200.initial_state(lin.to(dev), coeffs.to(dev), times.to(dev))
    close(got, want, rtol=1e-6)
    head = torch.nn.Sequential(torch.nn.Tanh(), torch.nn.Linear(H, 64), torch.nn.ReLU(), torch.nn.Linear(64, O)).eval()
    z = torch.randn(B, 7, H)
    want = head(z)
    got = snsde_b200.readout_head(head.to(dev), z.to(dev))
    assert got.shape == (B, 7, O)
    close(got, want, rtol=1e-5)
    assert snsde_b200.readout_head(torch.nn.Linear(H, O).to(dev), z.to(dev)) is None          # not a recognised stack
    assert snsde_b200.readout_head(head.train(), z.to(dev)) is None                           # training mode stays in PyTorch

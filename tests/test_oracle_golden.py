"""Oracle vs fixtures minted from the reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import philox, solver, spline, vector_field, wrapper

TOL = dict(atol=1e-6, rtol=1e-6)        # the reference test's own tolerance (tests/...alignment.py:127-128)


@pytest.fixture(scope="module")
def fg_cases(golden_dir):
    return torch.load(golden_dir / "fg_golden.pt")


def _build(case):
    B, K, C, H, HH, L = case["dims"]
    m = vector_field.DiffusionModel(C, H, HH, L, input_option=case["input_option"],
                                    noise_option=case["noise_option"])
    ref_keys = {k: tuple(v.shape) for k, v in case["state_dict"].items()}
    own_keys = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_keys == own_keys                      # same pin as the reference test :102-104
    m.load_state_dict(case["state_dict"])
    return m


def test_fg_all_options_match_reference(fg_cases):
    assert len(fg_cases) == 143
    seen = set()
    with torch.no_grad():
        for case in fg_cases:
            m = _build(case)
            m.set_X(case["coeffs"], case["times"])
            for i, t in enumerate(case["t"]):
                f, g = m.f(t, case["y"]), m.g(t, case["y"])
                assert torch.allclose(f, case["f"][i], **TOL), (case["input_option"], case["noise_option"])
                assert torch.allclose(g, case["g"][i], **TOL, equal_nan=True), (case["input_option"], case["noise_option"])
            seen.add((case["input_option"], case["noise_option"]))
    assert len(seen) == 140


def test_reference_test_fixture_recipe(fg_cases):
    named = [c for c in fg_cases if "name" in c]
    assert sorted(c["name"] for c in named) == ["gsde", "lnsde", "lsde"]
    for c in named:
        assert torch.isfinite(c["f"]).all() and torch.isfinite(c["g"]).all()
        assert c["f"].shape == (1, 2, 4)


def test_natural_spline_matches_in_tree_controldiffeq(golden_dir):
    for c in torch.load(golden_dir / "spline_golden.pt"):
        a, b, c2, d3 = spline.natural_cubic_spline_coeffs(c["times"], c["x"])
        for own, ref in ((a, c["a"]), (b, c["b"]), (c2, c["two_c"]), (d3, c["three_d"])):
            assert torch.allclose(own, ref, atol=1e-5, rtol=1e-5)
        if c.get("missing"):                   # missing-value branch: coefficients only
            continue
        sp = spline.CubicSpline(torch.cat([c["a"], c["b"], c["two_c"], c["three_d"]], -1), c["times"])
        ev = torch.stack([sp.evaluate(t) for t in c["tq"]])
        assert torch.allclose(ev, c["evaluate"], **TOL)


def test_interval_choice_equals_in_tree_rule():
    # bucketize(t)-1 (torchcde) == (t > times).sum()-1 (controldiffeq/interpolate.py:263), both clamped
    times = torch.tensor([0.0, 0.5, 1.25, 2.0, 3.5])
    sp = spline.CubicSpline(torch.zeros(1, 4, 4), times)
    for t in [-1.0, 0.0, 0.25, 0.5, 0.5001, 1.25, 3.4, 3.5, 9.0]:
        _, idx = sp.interpret_t(torch.tensor(t))
        want = int(((torch.tensor(t) > times).sum() - 1).clamp(0, 3))
        assert int(idx) == want


def test_hermite_interpolates_knots_and_is_c1_inside():
    torch.manual_seed(0)
    times = torch.tensor([0.0, 0.3, 1.0, 1.5, 2.7])
    x = torch.randn(2, 5, 3, dtype=torch.float64)
    co = spline.hermite_cubic_coefficients_with_backward_differences(x, times.double())
    sp = spline.CubicSpline(co, times.double())
    for k in range(5):
        assert torch.allclose(sp.evaluate(times[k].double()), x[:, k], atol=1e-12)
    # right end of interval k reproduces x[k+1] (left interval evaluated at full width)
    C = 3
    h = (times[1:] - times[:-1]).double()[None, :, None]
    a, b, c2, d3 = co[..., :C], co[..., C:2 * C], co[..., 2 * C:3 * C], co[..., 3 * C:]
    end = a + (b + (c2 / 2 + d3 * h / 3) * h) * h
    assert torch.allclose(end, x[:, 1:], atol=1e-12)
    # slope at the right end equals this interval's secant slope (backward difference), so the
    # next interval starts with the same derivative
    slope_end = b + (c2 + d3 * h) * h
    sec = (x[:, 1:] - x[:, :-1]) / h
    assert torch.allclose(slope_end, sec, atol=1e-10)
    assert torch.allclose(b[:, 1:], sec[:, :-1], atol=1e-12)


def test_hermite_nan_fill():
    times = torch.arange(5.0)
    x = torch.tensor([[float("nan"), 1.0, float("nan"), 3.0, float("nan")]]).T.unsqueeze(0)
    co = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    assert torch.allclose(co[0, :, 0], torch.tensor([1.0, 1.0, 2.0, 3.0]))


def test_forward_wrappers_match_reference(golden_dir):
    for c in torch.load(golden_dir / "forward_golden.pt"):
        if c["kind"].endswith("_full"):         # wrapper + its own neighbours: consumed by the GPU tests through patch()
            continue
        m = _build(c)
        bm = solver.BrownianTable(c["dW"])
        if c["kind"] == "classification":
            z = wrapper.classification_latent(m, c["times"], c["coeffs"], c["final_index"], c["z0"], bm)
        elif c["kind"] == "torch_ists":          # nsde_model.py:76-84: every knot, default method 'srk'
            m.set_X(c["coeffs"], c["times"])
            init = torch.nn.Linear(c["dims"][2], c["dims"][3])
            init.load_state_dict({k.split(".", 1)[1]: v for k, v in c["model_state"].items() if k.startswith("initial_network.")})
            with torch.no_grad():
                z0 = init(m.X.evaluate(c["times"][0]))
            bm = solver.BrownianTable(c["dW"], dU=c["dU"])
            z = wrapper.streamed_latent(m, c["times"], c["coeffs"], z0, bm, method=c["method"] or "srk")
        else:
            m.set_X(c["coeffs"], c["times"])
            init = torch.nn.Linear(c["dims"][2], c["dims"][3])
            init.load_state_dict(c["initial_network"])
            with torch.no_grad():
                z0 = init(m.X.evaluate(c["times"][0]))
            z = wrapper.streamed_latent(m, c["times"], c["coeffs"], z0, bm)[:, -c["output_time"]:, :]
        assert z.shape == c["z"].shape
        assert torch.allclose(z, c["z"], **TOL)


# ---- solver restatement: closed-form / structural checks (torchsde parity is unpinned) ----

class _OU(torch.nn.Module):
    sde_type, noise_type = "ito", "diagonal"

    def __init__(self, th, mu, sg):
        super().__init__()
        self.th, self.mu, self.sg = th, mu, sg

    def f(self, t, y):
        return self.th * (self.mu - y)

    def g(self, t, y):
        return torch.full_like(y, self.sg)


def test_euler_on_ou_matches_hand_recursion_and_lerp():
    ou = _OU(0.7, 0.2, 0.3)
    ts = torch.tensor([0.0, 0.13, 0.25, 0.5])
    dt = 0.1
    steps = solver.step_times(ts, dt)
    assert [round(b - a, 6) for a, b in steps] == [0.1] * 5
    dW = torch.randn(len(steps), 4, 2, dtype=torch.float64) * dt ** 0.5
    y0 = torch.ones(4, 2, dtype=torch.float64)
    out = solver.sdeint(ou, y0, ts.double(), dt, solver.BrownianTable(dW))
    ys, y = [y0], y0
    for k in range(5):
        y = y + 0.7 * (0.2 - y) * dt + 0.3 * dW[k]
        ys.append(y)
    assert torch.allclose(out[0], y0)
    assert torch.allclose(out[1], ys[1] + (ys[2] - ys[1]) * 0.3, atol=1e-12)     # t=0.13 in (0.1,0.2)
    assert torch.allclose(out[2], ys[2] + (ys[3] - ys[2]) * 0.5, atol=1e-12)     # t=0.25
    assert torch.allclose(out[3], ys[5], atol=1e-12)                             # lands on the grid


def test_step_plan_float32_sliver():
    # linspace(0,1,64) with dt=min diff -> 64 steps, the last one clamped to ts[-1]
    ts = torch.linspace(0, 1, 64)
    dt = solver.solver_dt(ts)
    st = solver.step_times(ts, dt)
    assert len(st) == 64 and st[-1][1] == 1.0 and (st[-1][1] - st[-1][0]) < 0.5 * dt
    ts = torch.linspace(0, 1, 20)
    assert len(solver.step_times(ts, 0.05)) == 20


class _GBM(torch.nn.Module):
    sde_type, noise_type = "ito", "diagonal"

    def f(self, t, y):
        return 0.1 * y

    def g(self, t, y):
        return 0.4 * y


def test_milstein_matches_closed_form_for_gbm():
    ts = torch.tensor([0.0, 0.5], dtype=torch.float64)
    dW = torch.randn(5, 3, 2, dtype=torch.float64) * 0.1 ** 0.5
    y0 = torch.rand(3, 2, dtype=torch.float64) + 0.5
    out = solver.sdeint(_GBM(), y0, ts, 0.1, solver.BrownianTable(dW), method="milstein")
    y = y0
    for k in range(5):
        y = y + 0.1 * y * 0.1 + 0.4 * y * dW[k] + 0.5 * 0.4 * 0.4 * y * (dW[k] ** 2 - 0.1)
    assert torch.allclose(out[-1], y, atol=1e-12)


class _Lin(torch.nn.Module):
    sde_type, noise_type = "ito", "diagonal"

    def f(self, t, y):
        return -0.8 * y

    def g(self, t, y):
        return torch.zeros_like(y)


def test_srk_without_noise_is_the_third_order_taylor_step():
    # SRID2 drift part: H0_1 = y + f0 h, H0_2 = y + (f0 + f1) h/4, y1 = y + h (f0 + f1 + 4 f2)/6; for f = a y this is
    # y (1 + z + z^2/2 + z^3/6), z = a h
    ts = torch.tensor([0.0, 0.3], dtype=torch.float64)
    y0 = torch.rand(3, 2, dtype=torch.float64) + 0.5
    z = torch.zeros(1, 3, 2, dtype=torch.float64)
    out = solver.sdeint(_Lin(), y0, ts, 0.3, solver.BrownianTable(z, dU=z), method="srk")
    zz = -0.8 * 0.3
    assert torch.allclose(out[-1], y0 * (1 + zz + zz ** 2 / 2 + zz ** 3 / 6), atol=1e-14)


def _coarsen(Wf, hf, n):
    """(dW, U) of n coarse steps from a fine Brownian path Wf [n_fine+1, P, 1]: U_k = int (W_s - W_t0) ds over the
    coarse step (trapezoid on the fine path)."""
    n_fine, P = Wf.shape[0] - 1, Wf.shape[1]
    m = n_fine // n
    Wc = Wf[::m]
    lo, hi = Wf[:-1].reshape(n, m, P, 1), Wf[1:].reshape(n, m, P, 1)
    U = (0.5 * (lo + hi) - Wc[:-1].unsqueeze(1)).sum(1) * hf
    return Wc[1:] - Wc[:-1], U


def test_srk_has_strong_order_1p5_and_needs_the_space_time_levy_integral():
    """Pins the SRID2 tableau and the U convention (I_k0 = int_t0^t1 (W_s - W_t0) ds) structurally.
    GBM (exact solution known): strong error falls ~2^1.5 per halving of h, Milstein ~2^1.  GBM cannot see U
    (L0 b = L1 a there), so an OU process (additive noise, L1 a = -theta sigma, L0 b = 0) checks it: with the true
    U the error keeps order 1.5, with Hst = 0 (U = h W/2) it degrades to order 1."""
    torch.manual_seed(3)
    n_fine, P = 2 ** 13, 600
    hf = 1.0 / n_fine
    dWf = torch.randn(n_fine, P, 1, dtype=torch.float64) * hf ** 0.5
    Wf = torch.cat([torch.zeros(1, P, 1, dtype=torch.float64), dWf.cumsum(0)])
    y0 = torch.ones(P, 1, dtype=torch.float64)
    ts = torch.tensor([0.0, 1.0], dtype=torch.float64)
    exact = torch.exp((0.1 - 0.5 * 0.4 ** 2) * 1.0 + 0.4 * Wf[-1])
    errs, errs_mil = [], []
    for n in (8, 32):
        dW, U = _coarsen(Wf, hf, n)
        y = solver.sdeint(_GBM(), y0, ts, 1.0 / n, solver.BrownianTable(dW, dU=U), method="srk")[-1]
        errs.append(float((y - exact).abs().mean()))
        y = solver.sdeint(_GBM(), y0, ts, 1.0 / n, solver.BrownianTable(dW), method="milstein")[-1]
        errs_mil.append(float((y - exact).abs().mean()))
    order, order_mil = np.log2(errs[0] / errs[1]) / 2, np.log2(errs_mil[0] / errs_mil[1]) / 2
    assert order > 1.3 and 0.8 < order_mil < 1.2 and errs[1] < 0.2 * errs_mil[1], (errs, errs_mil)

    ou = _OU(1.5, 0.0, 0.5)
    dW, U = _coarsen(Wf, hf, 1024)
    ref = solver.sdeint(ou, y0, ts, 1.0 / 1024, solver.BrownianTable(dW, dU=U), method="srk")[-1]
    e_true, e_wrong = [], []
    for n in (8, 32):
        dW, U = _coarsen(Wf, hf, n)
        y = solver.sdeint(ou, y0, ts, 1.0 / n, solver.BrownianTable(dW, dU=U), method="srk")[-1]
        e_true.append(float((y - ref).abs().mean()))
        y = solver.sdeint(ou, y0, ts, 1.0 / n, solver.BrownianTable(dW, dU=0.5 * dW / n), method="srk")[-1]
        e_wrong.append(float((y - ref).abs().mean()))
    o_true, o_wrong = np.log2(e_true[0] / e_true[1]) / 2, np.log2(e_wrong[0] / e_wrong[1]) / 2
    assert o_true > 1.3 and o_wrong < 1.15 and e_true[1] < 0.2 * e_wrong[1], (e_true, e_wrong, o_true, o_wrong)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        r = philox.philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        assert " ".join("%08x" % x for x in r) == want


def test_philox_normal_reference_moments():
    n = np.concatenate([philox.normals_reference(7, s, np.arange(64), 32).ravel() for s in range(40)])
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1) < 0.02
    assert abs((n ** 3).mean()) < 0.08 and abs((n ** 4).mean() - 3) < 0.15


def test_committed_conditioning_evidence_is_consistent():
    """tests/golden/conditioning.json (minted by make_conditioning.py with the CPU oracle) is the evidence behind the relaxed
    long-horizon tolerances: the fp32 reference path is itself 0.2 of the scale away from its fp64 evaluation on the
    GSDE model (c3), while all four are within 4e-7 over the first 24 steps - where the strict 1e-4 gates apply."""
    import json
    import pathlib
    ev = json.loads((pathlib.Path(__file__).parent / "golden" / "conditioning.json").read_text())
    assert set(ev) == {"c2", "c3", "c4", "c5"}
    assert ev["c3"]["max_rel"] > 1e-3 and ev["c3"]["median_rel"] < 1e-6          # ill-conditioned in rare elements only
    assert all(v["max_rel_first_24_steps"] < 1e-5 for v in ev.values())
    assert ev["c2"]["max_rel"] < 1e-4


def test_srk_reverse_sweep_site_order_and_coefficients():
    """The reverse of one SRID2 step as csrc/snsde_bwd_srk.cu performs it - stage states recomputed, the six evaluation
    sites visited in the order g3, g2, f2, g1, f1, (f0, g0), J^T products scattered with the transposed tableau
    coefficients - against autograd through the oracle's srk_step (fp64, a state-dependent drift and diffusion)."""
    torch.manual_seed(5)
    B, H = 3, 4
    A, Bm = torch.randn(H, H, dtype=torch.float64) * 0.5, torch.randn(H, H, dtype=torch.float64) * 0.3

    class Sde:
        sde_type, noise_type = "ito", "diagonal"
        def f(self, t, y): return torch.tanh(y @ A.T + torch.sin(t))
        def g(self, t, y): return torch.tanh(0.7 * (y @ Bm.T) * torch.cos(t))
    sde = Sde()
    t0, t1 = torch.tensor(0.3, dtype=torch.float64), torch.tensor(0.55, dtype=torch.float64)
    h = t1 - t0
    sq = h.sqrt()
    W = torch.randn(B, H, dtype=torch.float64) * sq
    U = h * (W / 2 + torch.randn(B, H, dtype=torch.float64) * (h / 12).sqrt())
    y0 = torch.randn(B, H, dtype=torch.float64, requires_grad=True)
    lam = torch.randn(B, H, dtype=torch.float64)
    y1 = solver.srk_step(sde, solver.BrownianTable(W[None], dU=U[None]), t0, t1, y0)
    (want,) = torch.autograd.grad(y1, y0, grad_outputs=lam)

    def vjp(fn, t, z, cot):                      # J(z)^T cot of one evaluation site
        z = z.detach().requires_grad_(True)
        (out,) = torch.autograd.grad(fn(t, z), z, grad_outputs=cot)
        return out
    with torch.no_grad():
        y = y0.detach()
        tq, th = t0 + 0.25 * h, t0 + 0.5 * h
        f0, g0 = sde.f(t0, y), sde.g(t0, y)
        H01 = y + f0 * h
        H11 = y + 0.25 * f0 * h - 0.5 * g0 * sq
        f1, g1 = sde.f(t1, H01), sde.g(tq, H11)
        H02 = y + 0.25 * f0 * h + g0 * U / h + 0.25 * f1 * h + 0.5 * g1 * U / h
        H12 = y + f0 * h + g0 * sq
        f2, g2 = sde.f(th, H02), sde.g(t1, H12)
        H13 = y + 2 * g0 * sq - g1 * sq + 0.25 * f2 * h + 0.5 * g2 * sq
        Ikk, Ikkk = (W ** 2 - h) / 2, (W ** 3 - 3 * h * W) / 6
        c0, c1, c2 = Ikk / sq, U / h, Ikkk / h
        gw = [-W + c0 + 2 * c1 - 2 * c2, 4 / 3 * W - 4 / 3 * c0 - 4 / 3 * c1 + 5 / 3 * c2, 2 / 3 * W + c0 / 3 - 2 / 3 * c1 - 2 / 3 * c2, c2]
    yb = lam.clone()
    fb = [lam * h / 6, lam * h / 6, lam * 2 * h / 3]
    gb = [lam * gw[0], lam * gw[1], lam * gw[2], lam * gw[3]]
    z = vjp(sde.g, tq, H13, gb[3]); yb += z; gb[0] = gb[0] + 2 * sq * z; gb[1] = gb[1] - sq * z; fb[2] = fb[2] + 0.25 * h * z; gb[2] = gb[2] + 0.5 * sq * z
    z = vjp(sde.g, t1, H12, gb[2]); yb += z; fb[0] = fb[0] + h * z; gb[0] = gb[0] + sq * z
    z = vjp(sde.f, th, H02, fb[2]); yb += z; fb[0] = fb[0] + 0.25 * h * z; gb[0] = gb[0] + U / h * z; fb[1] = fb[1] + 0.25 * h * z; gb[1] = gb[1] + 0.5 * U / h * z
    z = vjp(sde.g, tq, H11, gb[1]); yb += z; fb[0] = fb[0] + 0.25 * h * z; gb[0] = gb[0] - 0.5 * sq * z
    z = vjp(sde.f, t1, H01, fb[1]); yb += z; fb[0] = fb[0] + h * z
    yb += vjp(sde.f, t0, y, fb[0]) + vjp(sde.g, t0, y, gb[0])
    assert torch.allclose(yb, want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("kind", ["MU_Y", "MU_TY", "SQRT", "CUBE", "SIGMOID", "RELU"])
@pytest.mark.parametrize("bounded", [True, False])
def test_milstein_term_closed_form_cotangents(kind, bounded):
    """The closed forms of csrc/snsde_math.cuh `milstein_backward` (cotangents of T = 0.5 v g dg/dy w.r.t. y, the
    coefficient and sigmoid(theta), through BOTH factors as torchsde's create_graph vjp differentiates them) against
    double autograd, for every state-dependent elementwise diffusion of neuralsde.py:233-307."""
    torch.manual_seed(0)
    y = (torch.rand(6, dtype=torch.float64) + 0.3).requires_grad_(True)
    coef = torch.randn(6, dtype=torch.float64).requires_grad_(True)
    s = torch.tensor(0.7, dtype=torch.float64, requires_grad=True)
    tt, v = 1.3, torch.randn(6, dtype=torch.float64)
    raw = {"MU_Y": lambda: coef * y, "MU_TY": lambda: tt * y, "SQRT": lambda: y.sqrt(), "CUBE": lambda: y ** 3,
           "SIGMOID": lambda: torch.sigmoid(y), "RELU": lambda: torch.relu(y)}[kind]()
    g = torch.tanh(s * torch.nan_to_num(raw)) if bounded else raw
    (gdg,) = torch.autograd.grad(g, y, grad_outputs=g * v, create_graph=True)
    ay, ac, as_ = torch.autograd.grad((0.5 * gdg).sum(), (y, coef, s), allow_unused=True)
    # ---- the device function, transcribed ----
    yd, cd, sd = y.detach(), coef.detach(), s.detach()
    r2 = rc = r1c = torch.zeros_like(yd)
    if kind == "MU_Y": rw, r1, rc, r1c = cd * yd, cd, yd, torch.ones_like(yd)
    elif kind == "MU_TY": rw, r1 = tt * yd, torch.full_like(yd, tt)
    elif kind == "SQRT": rw = yd.sqrt(); r1 = 0.5 / rw; r2 = -0.25 / (rw * yd)
    elif kind == "CUBE": rw, r1, r2 = yd ** 3, 3 * yd * yd, 6 * yd
    elif kind == "SIGMOID": rw = 1 / (1 + torch.exp(-yd)); r1 = rw * (1 - rw); r2 = r1 * (1 - 2 * rw)
    else: rw, r1 = torch.relu(yd), (yd > 0).double()
    hv = 0.5 * v
    if bounded:
        gd = torch.tanh(sd * rw); q = 1 - gd * gd
        g1 = q * sd * r1; g2 = -2 * gd * g1 * sd * r1 + q * sd * r2
        gc = q * sd * rc; g1c = -2 * gd * gc * sd * r1 + q * sd * r1c
        gs = q * rw; g1s = -2 * gd * gs * sd * r1 + q * r1
        my, mc, ms = hv * (g1 * g1 + gd * g2), hv * (gc * g1 + gd * g1c), (hv * (gs * g1 + gd * g1s)).sum()
    else:
        my, mc, ms = hv * (r1 * r1 + rw * r2), hv * (rc * r1 + rw * r1c), torch.tensor(0.0, dtype=torch.float64)
    assert torch.allclose(ay, my, rtol=1e-10, atol=1e-12)
    assert torch.allclose(ac if ac is not None else torch.zeros_like(mc), mc, rtol=1e-10, atol=1e-12)
    assert torch.allclose(as_ if as_ is not None else torch.tensor(0.0, dtype=torch.float64), ms, rtol=1e-10, atol=1e-12)

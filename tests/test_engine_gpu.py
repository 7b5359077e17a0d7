"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
identical inputs and identical Brownian increments.

Tolerance (BASELINE.json north_star: "within 1e-4 relative fp32"):
    max |z_gpu - z_oracle| <= 1e-4 * max(|z_oracle|_inf, 1)      per output tensor
i.e. relative to the scale of the latent.  fp32 re-association alone (the oracle itself run
in fp64 vs fp32) moves results by ~1e-6 on these trajectories.
"""
import numpy as np
import pytest
import torch

import snsde_b200
from oracle import philox, solver, spline, vector_field, wrapper

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def close(got, want, rtol=RTOL):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = max(float(want.abs().max()), 1.0)
    err = float((got - want).abs().max())
    assert err <= rtol * scale, f"max abs err {err:.3e} > {rtol:g} * {scale:.3g}"
    return err / scale


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def make_problem(io, no, B, H, C, L, K, seed, HH=None, family="benchmark", spacing=1.0):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    HH = H if HH is None else HH
    if family == "tutorial":
        m = vector_field.TutorialLSDEFunc(C, H, HH, L)
    else:
        m = vector_field.DiffusionModel(C, H, HH, L, theta=0.8, sigma=-0.5, input_option=io, noise_option=no)
    times = torch.arange(K, dtype=torch.float32) * spacing
    x = torch.cat([times[None, :, None].expand(B, K, 1),
                   (torch.randn(B, K, C - 1, generator=g) * 0.1).cumsum(1)], dim=-1) if C > 1 else \
        (torch.randn(B, K, 1, generator=g) * 0.1).cumsum(1)
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    y0 = torch.randn(B, H, generator=g) * 0.5
    return m, times, coeffs, y0


def run_both(m, times, coeffs, y0, ts, dt, dW, method, dev, precision="fp32"):
    m.set_X(coeffs, times)
    want = solver.sdeint(m, y0, ts, dt, solver.BrownianTable(dW), method=method)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        got = snsde_b200.sdeint(mg, y0.to(dev), ts.to(dev), dt=dt, method=method,
                                bm=snsde_b200.BrownianIncrements(dW.to(dev)), precision=precision)
    m.to("cpu")
    return got, want


# ---- every (input_option, noise_option): one Euler step against the REFERENCE's golden f/g ----
def test_single_step_all_140_option_pairs_vs_reference_golden(golden_dir, dev):
    cases = torch.load(golden_dir / "fg_golden.pt")
    h = 0.25
    worst = 0.0
    for c in cases:
        B, K, C, H, HH, L = c["dims"]
        m = vector_field.DiffusionModel(C, H, HH, L, input_option=c["input_option"], noise_option=c["noise_option"])
        m.load_state_dict(c["state_dict"])
        m = m.to(dev)
        m.set_X(c["coeffs"].to(dev), c["times"].to(dev))
        g = torch.Generator().manual_seed(5)
        dW = torch.randn(1, B, H, generator=g) * h ** 0.5
        for i, t in enumerate(c["t"]):
            ts = torch.stack([t, t + h])
            with torch.no_grad():
                got = snsde_b200.sdeint(m, c["y"].to(dev), ts, dt=1.0, bm=snsde_b200.BrownianIncrements(dW),
                                        precision="fp32")
            hh = (ts[1] - ts[0])
            want = c["y"] + c["f"][i] * hh + c["g"][i] * dW[0]
            assert torch.equal(got[0].cpu(), c["y"])
            worst = max(worst, close(got[1], want, rtol=2e-6))
    print("worst single-step rel err", worst)


NAMED = [("lnsde", 4, 17), ("gsde", 6, 17), ("lsde", 2, 16), ("neuralsde_3_18", 3, 18), ("naivesde", 1, 18),
         ("staticsde", 1, 0)]


@pytest.mark.parametrize("name,io,no", NAMED)
@pytest.mark.parametrize("H,C,L,B", [(32, 5, 1, 19), (64, 35, 2, 40), (128, 21, 1, 33)])
def test_euler_trajectories_named_models(name, io, no, H, C, L, B, dev):
    K = 41
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=H + io)
    dt = solver.solver_dt(times)
    ts = times[[0, 3, 17, 40]]
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(1)) * dt ** 0.5
    got, want = run_both(m, times, coeffs, y0, ts, dt, dW, "euler", dev)
    close(got, want)


@pytest.mark.parametrize("io,no", [(6, 17), (4, 17), (2, 16), (1, 3), (3, 6), (5, 8), (1, 9), (3, 10), (5, 11), (4, 13),
                                   (1, 7), (0, 12), (1, 0), (2, 5),
                                   # state-dependent noise NETWORKS: full vjp through noise_y (torchsde gdg_prod)
                                   (3, 18), (1, 19), (4, 14), (6, 15), (0, 19), (2, 18)])
def test_milstein_vs_autograd_oracle(io, no, dev):
    B, H, C, L, K = 11, 16, 4, 2, 13
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=no, HH=(24 if io in (1, 3, 5) else None),
                                        spacing=0.25)
    if no == 7:
        y0 = y0.abs() + 0.5          # sqrt(y): stay positive so the autograd path is finite
    dt = solver.solver_dt(times)
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(2)) * dt ** 0.5 * (0.2 if no in (7, 8) else 1.0)
    got, want = run_both(m, times, coeffs, y0, times, dt, dW, "milstein", dev)
    if no == 7:
        ok = torch.isfinite(want)
        assert torch.equal(torch.isfinite(got.cpu()), ok)
        got, want = torch.where(ok, got.cpu(), torch.zeros(())), torch.where(ok, want, torch.zeros(()))
    close(got, want)


def test_unaligned_grid_tutorial_lsde_c1(dev):
    # BASELINE config c1 (scaled-up tutorial): knots linspace(0,1,20), dt=0.05 -> outputs are lerps
    B, H, C, L = 64, 32, 2, 1
    m, _, _, y0 = make_problem(0, 0, B, H, C, L, 20, seed=3, family="tutorial")
    times = torch.linspace(0, 1, 20)
    x = torch.randn(B, 20, C, generator=torch.Generator().manual_seed(4)).cumsum(1) * 0.2
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    S = len(solver.step_times(times, 0.05))
    dW = torch.randn(S, B, H, generator=torch.Generator().manual_seed(5)) * 0.05 ** 0.5
    got, want = run_both(m, times, coeffs, y0, times, 0.05, dW, "euler", dev)
    close(got, want)
    m2, _, _, _ = make_problem(0, 0, B, H, C, 3, 20, seed=6, family="tutorial", HH=48)
    got, want = run_both(m2, times, coeffs, y0, times, 0.05, dW, "euler", dev)
    close(got, want)


def test_sliver_step_grid_and_stream_all_knots(dev):
    B, H, C, L, K = 9, 32, 3, 1, 64
    m, _, _, y0 = make_problem(4, 17, B, H, C, L, K, seed=8)
    times = torch.linspace(0, 1, K)               # torch-ists grid: dt = min diff < 1/63 -> 64 steps
    x = torch.randn(B, K, C, generator=torch.Generator().manual_seed(9)).cumsum(1) * 0.1
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
    dt = solver.solver_dt(times)
    S = len(solver.step_times(times, dt))
    assert S == 64
    dW = torch.randn(S, B, H, generator=torch.Generator().manual_seed(10)) * dt ** 0.5
    got, want = run_both(m, times, coeffs, y0, times, dt, dW, "euler", dev)
    close(got, want)


def test_fused_final_index_gather_vs_reference_golden_and_oracle(golden_dir, dev):
    for c in torch.load(golden_dir / "forward_golden.pt"):
        if c["kind"] != "classification":
            continue
        B, K, C, H, HH, L = c["dims"]
        m = vector_field.DiffusionModel(C, H, HH, L, input_option=c["input_option"], noise_option=c["noise_option"])
        m.load_state_dict(c["state_dict"])
        m = m.to(dev)
        m.set_X(c["coeffs"].to(dev), c["times"].to(dev))
        with torch.no_grad():
            z = snsde_b200.solve_final(m, c["times"].to(dev), c["final_index"].to(dev), c["z0"].to(dev),
                                       bm=snsde_b200.BrownianIncrements(c["dW"].to(dev)), precision="fp32")
        close(z, c["z"], rtol=1e-5)
    # larger ragged case vs the oracle wrapper
    B, H, C, L, K = 37, 64, 6, 1, 30
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=21)
    fi = torch.randint(2, K, (B,), generator=torch.Generator().manual_seed(3))
    fi[0], fi[1] = 0, K - 1
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(4))
    want = wrapper.classification_latent(m, times, coeffs, fi, y0, solver.BrownianTable(dW))
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        got = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev),
                                     bm=snsde_b200.BrownianIncrements(dW.to(dev)), precision="fp32")
    close(got, want)
    assert torch.equal(got[0].cpu(), y0[0])          # final_index 0 -> z0 itself


def test_patch_drop_in_on_reference_style_wrapper(dev):
    class RefStyleNeuralSDE(torch.nn.Module):      # shape of reference NeuralSDE (neuralsde.py:51-120)
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

    B, H, C, L, K = 12, 32, 4, 1, 17
    func, times, coeffs, _ = make_problem(4, 17, B, H, C, L, K, seed=31)
    model = RefStyleNeuralSDE(func, C, H, 3)
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(1))
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        func.set_X(coeffs, times)
        z0 = model.initial_network(func.X.evaluate(times[0]))
        want = model.linear(wrapper.classification_latent(func, times, coeffs, fi, z0, solver.BrownianTable(dW)))
        model = snsde_b200.patch(model.to(dev))
        got = model(times.to(dev), [coeffs.to(dev)], fi.to(dev), bm=snsde_b200.BrownianIncrements(dW.to(dev)),
                    precision="fp32")
        close(got, want)
        zt = model._solve_sde_path(times.to(dev), times.to(dev), z0.to(dev),
                                   {"bm": snsde_b200.BrownianIncrements(dW.to(dev)), "precision": "fp32"})
        func.to("cpu"); func.set_X(coeffs, times)
        close(zt, solver.sdeint(func, z0, times, 1.0, solver.BrownianTable(dW)))


def test_in_kernel_philox_equals_table_replay_and_matches_numpy_reference(dev):
    B, H, C, L, K = 24, 32, 3, 1, 12
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=41, spacing=0.5)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    dt = solver.solver_dt(times)
    with torch.no_grad():
        a = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt, seed=1234, precision="fp32")
        plan = snsde_b200.plans_of(mg)[("euler", "fp32", str(dev))]
        sp = plan.step_plan(times, dt, times)
        dW = snsde_b200.philox_increments(1234, sp, B, H, dev)
        b = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt, bm=snsde_b200.BrownianIncrements(dW), precision="fp32")
        c = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt, seed=1235, precision="fp32")
    assert torch.equal(a, b)                       # identical increments -> bit-identical trajectories
    assert not torch.equal(a, c)
    m.to("cpu"); m.set_X(coeffs, times)
    close(a, solver.sdeint(m, y0, times, dt, solver.BrownianTable(dW.cpu())))
    # the integer stream + normal map against the float64 numpy restatement
    for s in (0, 5):
        ref = philox.normals_reference(1234, s, np.arange(B), H) * float(sp.steps["sqrt_h"][s])
        assert np.abs(dW[s].cpu().numpy() - ref).max() < 2e-5
    big = snsde_b200.philox_increments(7, sp, 4096, 64, dev).cpu().numpy() / np.sqrt(dt)
    assert abs(big.mean()) < 3e-3 and abs(big.std() - 1) < 3e-3
    assert abs(np.corrcoef(big[0].ravel(), big[1].ravel())[0, 1]) < 0.01      # steps independent
    assert abs(np.corrcoef(big[0, :-1].ravel(), big[0, 1:].ravel())[0, 1]) < 0.01   # rows independent


def test_batch_sharding_is_bit_invariant(dev):
    B, H, C, L, K = 50, 32, 3, 1, 9
    m, times, coeffs, y0 = make_problem(6, 17, B, H, C, L, K, seed=51)
    mg = m.to(dev)
    cg, tg, yg = coeffs.to(dev), times.to(dev), y0.to(dev)
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        mg.set_X(cg, tg)
        full = snsde_b200.solve_final(mg, tg, fi, yg, seed=99, precision="fp32")
        parts = []
        for lo, hi in ((0, 13), (13, 26), (26, 50)):        # shard starts not multiples of 4 on purpose
            mg.set_X(cg[lo:hi], tg)
            # every shard must use the GLOBAL output-time set so slots agree
            ts, slots = snsde_b200.final_index_slots(tg, fi)
            plan = snsde_b200.plans_of(mg)[("euler", "fp32", str(dev))]
            sp = plan.step_plan(ts, 1.0, tg)
            parts.append(plan.forward(yg[lo:hi], sp, coeffs=cg[lo:hi], row_slot=slots[lo:hi], seed=99, row_offset=lo))
    assert torch.equal(full, torch.cat(parts))


def test_edge_shapes_and_errors(dev):
    # B=1, H=4 (reference-test size), single output time (no steps), non-contiguous coeffs
    m, times, coeffs, y0 = make_problem(4, 17, 1, 4, 3, 2, 5, seed=61)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        z = snsde_b200.sdeint(mg, y0.to(dev), times[:1].to(dev), dt=1.0, seed=1, precision="fp32")
        assert z.shape == (1, 1, 4) and torch.equal(z[0].cpu(), y0)
        wide = torch.zeros(1, 4, 24, device=dev)
        wide[..., :12] = coeffs.to(dev)
        mg.set_X(wide[..., :12], times.to(dev))
        dW = torch.randn(4, 1, 4)
        got = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=snsde_b200.BrownianIncrements(dW), precision="fp32")
    m.to("cpu"); m.set_X(coeffs, times)
    close(got, solver.sdeint(m, y0, times, 1.0, solver.BrownianTable(dW)))
    mg = m.to(dev); mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        with pytest.raises(ValueError):
            snsde_b200.sdeint(mg, torch.zeros(1, 5, device=dev), times.to(dev), dt=1.0)
        with pytest.raises(ValueError):
            snsde_b200.sdeint(mg, y0.to(dev), times.flip(0).to(dev), dt=1.0)
        with pytest.raises(ValueError):
            snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, method="heun")
        with pytest.raises(ValueError, match="dU"):
            snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, method="srk",
                              bm=snsde_b200.BrownianIncrements(torch.zeros(4, 1, 4)))
        with pytest.raises(ValueError):
            snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=snsde_b200.BrownianIncrements(torch.zeros(3, 1, 4)))


def test_large_hidden_and_nan_semantics(dev):
    # H=256 (c5 width): weights exceed shared memory -> L2 streaming branch of the FMA kernel
    m, times, coeffs, y0 = make_problem(4, 17, 10, 256, 14, 1, 9, seed=71)
    dW = torch.randn(8, 10, 256, generator=torch.Generator().manual_seed(1))
    got, want = run_both(m, times, coeffs, y0, times, 1.0, dW, "euler", dev)
    close(got, want)
    # noise option 7 (sqrt y) with negative state: NaN -> nan_to_num -> 0, exactly like the reference
    m, times, coeffs, y0 = make_problem(1, 7, 6, 8, 2, 1, 5, seed=72)
    dW = torch.randn(4, 6, 8, generator=torch.Generator().manual_seed(2))
    got, want = run_both(m, times, coeffs, y0, times, 1.0, dW, "euler", dev)
    assert torch.isfinite(got).all()
    close(got, want)


# ================================ tcgen05 tensor-core kernel ======================================
TC_CASES = [
    # io, no, H, C, L, B, method
    (4, 17, 128, 35, 1, 40, "euler"),        # c2 model (collapsed emb o linear_in, X(t) operand ring)
    (6, 17, 64, 35, 1, 70, "milstein"),      # c3 model
    (2, 16, 64, 5, 2, 19, "euler"),          # LSDE, additive noise, one hidden layer more
    (4, 17, 32, 3, 3, 9, "euler"),
    (3, 6, 64, 4, 1, 33, "milstein"),        # no control read, diagonal sigma * y
    (1, 3, 128, 4, 1, 12, "euler"),
    (5, 13, 96, 7, 2, 27, "euler"),          # geometric drift, Linear(2,H) noise * y
    (2, 0, 32, 2, 1, 5, "euler"),
    (6, 9, 64, 6, 1, 300, "euler"),          # B > 148*... multiple rows per CTA (NR=16)
    (4, 17, 128, 35, 2, 30, "euler"),        # H=128 with a hidden layer: 448 TMEM columns of weights -> last images SS-form from smem
    (2, 16, 128, 9, 4, 8, "euler"),          # 5 drift layers: most images stay in shared memory
    (4, 17, 128, 35, 1, 600, "euler"),       # NR=32 rows per CTA: wider accumulators leave 320 TMEM columns for weights
]


@pytest.mark.parametrize("io,no,H,C,L,B,method", TC_CASES)
def test_tc_kernel_matches_oracle(io, no, H, C, L, B, method, dev):
    K = 25
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=7 * H + io + no)
    dt = solver.solver_dt(times)
    ts = times[[0, 1, 2, 11, 24]]
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(1)) * dt ** 0.5
    got, want = run_both(m, times, coeffs, y0, ts, dt, dW, method, dev, precision="tc")
    assert snsde_b200.plans_of(m)[(method, "tc", str(dev))].kernel == "tcgen05"
    close(got, want)


def test_tc_kernel_long_trajectory_c2_shape_and_philox(dev):
    # c2 shape at reduced batch: 200 steps, in-kernel Philox; FMA and tcgen05 kernels see the same increments
    B, H, C, L, K = 64, 128, 35, 1, 201
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=5)
    fi = torch.randint(2, K, (B,), generator=torch.Generator().manual_seed(3))
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        z_tc = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=77, precision="tc")
        z_fma = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=77, precision="fp32")
        plan = snsde_b200.plans_of(mg)[("euler", "tc", str(dev))]
        assert plan.kernel == "tcgen05"
        sp = plan.step_plan(times, 1.0, times)
        dW = snsde_b200.philox_increments(77, sp, B, H, dev).cpu()
    m.to("cpu")
    want = wrapper.classification_latent(m, times, coeffs, fi, y0, solver.BrownianTable(dW))
    e1 = close(z_fma, want)
    e2 = close(z_tc, want)
    print(f"200-step c2-shape rel err: fma {e1:.2e}  tcgen05 {e2:.2e}")


def test_tc_auto_falls_back_to_fma_when_unsupported(dev):
    m, times, coeffs, y0 = make_problem(0, 17, 4, 32, 3, 1, 5, seed=1)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=1)
        assert snsde_b200.plans_of(mg)[("euler", "auto", str(dev))].kernel == "fma_fp32"
        with pytest.raises(ValueError, match="tensor-core"):
            snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=1, precision="tc")


# ======================= general tcgen05 kernel (streamed weights, 2 M tiles, noise networks) =======================
TCG_CASES = [
    # io, no, H, C, L, B, method      what it exercises
    (3, 18, 128, 21, 1, 24, "euler"),      # c4 model: state-dependent 2-layer noise net beside the drift; 4 H x H matrices
    (1, 19, 64, 3, 1, 40, "euler"),        # 2-layer noise net * y
    (5, 15, 32, 4, 2, 9, "euler"),         # 1-layer noise net * y, geometric drift, deeper drift than noise net
    (4, 14, 64, 6, 1, 17, "euler"),        # 1-layer noise net with control
    (4, 18, 96, 5, 3, 11, "euler"),        # noise net finishes two phases before the drift
    (4, 17, 256, 14, 1, 20, "euler"),      # c5 model: two M tiles, weights (557 KB) streamed through the ring
    (6, 17, 192, 7, 1, 13, "milstein"),    # two M tiles with a partial second tile
    (4, 17, 128, 35, 2, 30, "euler"),      # H=128 with a hidden layer: partly streamed
    (2, 16, 128, 9, 4, 8, "euler"),        # 5 drift phases
    (3, 13, 256, 3, 2, 150, "euler"),      # H=256, no control, NR=8 with >148... single wave, Linear(2,H) noise table
]


@pytest.mark.parametrize("io,no,H,C,L,B,method", TCG_CASES)
def test_general_tc_kernel_matches_oracle(io, no, H, C, L, B, method, dev, monkeypatch):
    monkeypatch.setenv("SNSDE_FORCE_TCG", "1")       # some of these shapes also fit the resident kernel (weights in TMEM)
    K = 21
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=3 * H + io + no)
    dt = solver.solver_dt(times)
    ts = times[[0, 1, 2, 9, 20]]
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(1)) * dt ** 0.5
    got, want = run_both(m, times, coeffs, y0, ts, dt, dW, method, dev, precision="tc")
    assert snsde_b200.plans_of(m)[(method, "tc", str(dev))].kernel == "tcgen05_general"
    close(got, want)


def test_general_kernel_agrees_with_resident_kernel(dev, monkeypatch):
    """Same model through both tensor-core kernels (SNSDE_FORCE_TCG selects the general one)."""
    B, H, C, L, K = 70, 128, 35, 1, 41
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=9)
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(3))
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        a = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=21, precision="tc")
        assert snsde_b200.plans_of(mg)[("euler", "tc", str(dev))].kernel == "tcgen05"
        snsde_b200.engine._PLANS.pop(mg, None)
        monkeypatch.setenv("SNSDE_FORCE_TCG", "1")
        b = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=21, precision="tc")
        assert snsde_b200.plans_of(mg)[("euler", "tc", str(dev))].kernel == "tcgen05_general"
    close(b, a, rtol=1e-5)


def test_weights_in_tmem_agree_with_weights_in_smem(dev, monkeypatch):
    """TS-form MMAs (weight images resident in TMEM) against the SS-form path (SNSDE_TC_NO_TMEM) and two chains."""
    B, H, C, L, K = 70, 128, 35, 1, 41
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=9)
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(3))
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        a = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=21, precision="tc")
        assert snsde_b200.plans_of(mg)[("euler", "tc", str(dev))].kernel == "tcgen05"
        monkeypatch.setenv("SNSDE_TC_NO_TMEM", "1")
        b = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=21, precision="tc")
        monkeypatch.delenv("SNSDE_TC_NO_TMEM")
        monkeypatch.setenv("SNSDE_TC_CH", "2")
        c = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=21, precision="tc")
    assert torch.equal(a, b)               # same products, same accumulation order: only the operand source differs
    close(c, a, rtol=1e-5)


def test_fp16_range_overflow_is_flagged_not_silent(dev):
    m, times, coeffs, y0 = make_problem(4, 17, 6, 64, 3, 1, 6, seed=2)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    with torch.no_grad():
        snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=1, precision="tc")
        plan = snsde_b200.plans_of(mg)[("euler", "tc", str(dev))]
        assert plan.status() == 0
        big = y0.clone()
        big[2, 5] = 1.0e5                                   # beyond the split-fp16 operand range
        z_tc = snsde_b200.sdeint(mg, big.to(dev), times.to(dev), dt=1.0, seed=1, precision="tc")
        assert plan.status() == 1 and plan.status() == 0    # sticky until read, then cleared
        z32 = snsde_b200.sdeint(mg, big.to(dev), times.to(dev), dt=1.0, seed=1, precision="fp32")
    # the fp32 kernel has no such limit; unaffected rows of the tensor-core result are still right
    ok = [0, 1, 3, 4, 5]
    close(z_tc[:, ok], z32[:, ok])


def test_hermite_coefficients_on_device_match_the_oracle_builder(dev):
    from snsde_b200 import data
    g = torch.Generator().manual_seed(0)
    for B, K, C in ((3, 2, 1), (5, 7, 3), (64, 201, 35)):
        times = torch.cat([torch.zeros(1), torch.rand(K - 1, generator=g) + 0.3]).cumsum(0)
        x = torch.randn(B, K, C, generator=g).cumsum(1)
        want = spline.hermite_cubic_coefficients_with_backward_differences(x, times)
        got = data.hermite_coeffs_cuda(x.to(dev), times.to(dev))
        assert got.shape == want.shape
        assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6)
        # and the solve accepts them directly
    m, times, coeffs, y0 = make_problem(4, 17, 8, 32, 3, 1, 9, seed=5)
    x = torch.cat([coeffs[:, :, :3], (coeffs[:, -1:, :3] + coeffs[:, -1:, 3:6] * 1.0)], dim=1)   # knot values back from a,b (h = 1)
    with torch.no_grad():
        mg = m.to(dev)
        c_dev = data.hermite_coeffs_cuda(x.to(dev), times.to(dev))
        mg.set_X(c_dev, times.to(dev))
        z = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=3)
    assert torch.isfinite(z).all()


def test_natural_spline_coefficients_on_device_match_the_pinned_oracle(dev):
    """snsde_natural_coeffs against the oracle builder (itself pinned to the reference's in-tree controldiffeq by
    tests/golden/spline_golden.pt) and, bit for bit, against the torch-op chain the package ships for data prep."""
    from snsde_b200 import data
    g = torch.Generator().manual_seed(1)
    for B, K, C in ((3, 2, 1), (4, 3, 2), (5, 7, 3), (64, 501, 14), (300, 33, 5)):
        times = torch.cat([torch.zeros(1), torch.rand(K - 1, generator=g) + 0.3]).cumsum(0)
        x = torch.randn(B, K, C, generator=g).cumsum(1)
        want = torch.cat(spline.natural_cubic_spline_coeffs(times, x), dim=-1)
        got = data.natural_coeffs_cuda(x.to(dev), times.to(dev)).cpu()
        assert got.shape == want.shape == (B, K - 1, 4 * C)
        scale = want.abs().amax(dim=(0, 1), keepdim=True).clamp_min(1.0)
        assert float(((got - want).abs() / scale).max()) <= 2e-5
        chain = data.natural_cubic_coeffs(x.to(dev), times.to(dev)).cpu()
        assert torch.equal(got, chain)
        # the spline interpolates the knots and is C2: value / first / second derivative continuous at interior knots
        a, b, c2, d3 = got.double().chunk(4, dim=-1)
        h = (times[1:] - times[:-1]).double()[None, :, None]
        end = a + (b + (c2 / 2 + d3 * h / 3) * h) * h
        assert torch.allclose(end[:, :-1], a[:, 1:], atol=1e-4 * float(x.abs().max()))
        assert torch.allclose(end[:, -1], x[:, -1].double(), atol=1e-4 * float(x.abs().max()))


def test_missing_value_fill_on_device_matches_the_oracle_and_feeds_the_hermite_builder(dev):
    from snsde_b200 import data
    g = torch.Generator().manual_seed(2)
    B, K, C = 40, 23, 6
    times = torch.cat([torch.zeros(1), torch.rand(K - 1, generator=g) + 0.2]).cumsum(0)
    x = torch.randn(B, K, C, generator=g).cumsum(1)
    holes = torch.rand(B, K, C, generator=g) < 0.35
    holes[0] = False                       # a complete series
    holes[1, :, 0] = True                  # nothing observed: stays NaN
    holes[2, :5, 1] = True                 # leading gap
    holes[3, -6:, 2] = True                # trailing gap
    xm = x.masked_fill(holes, float("nan"))
    want = spline._fill_missing_linear(xm, times)
    got = data.fill_missing_cuda(xm.to(dev), times.to(dev)).cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.isnan(got[1, :, 0]).all() and not torch.isnan(got[0]).any()
    ok = ~torch.isnan(want)
    assert torch.allclose(got[ok], want[ok], rtol=1e-6, atol=1e-6)
    keep = torch.ones(B, dtype=torch.bool); keep[1] = False
    co_want = spline.hermite_cubic_coefficients_with_backward_differences(xm[keep], times)
    co_got = data.hermite_coeffs_cuda(xm[keep].to(dev), times.to(dev), fill_missing=True).cpu()
    assert torch.allclose(co_got, co_want, rtol=1e-4, atol=1e-5)


def test_noise_table_cache_follows_weight_updates_and_plan_is_safe_across_streams(dev):
    """The per-step noise table of the tcgen05 path is reused while (step times, weights) are unchanged; it must be
    rebuilt after a parameter update, and one plan used from two streams must serialise its solves."""
    B, H, C, L, K = 48, 64, 5, 1, 17
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=11)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    y0d, td = y0.to(dev), times.to(dev)
    with torch.no_grad():
        a1 = snsde_b200.sdeint(mg, y0d, td, dt=1.0, seed=5, precision="tc")
        a2 = snsde_b200.sdeint(mg, y0d, td, dt=1.0, seed=5, precision="tc")          # cached table
        assert torch.equal(a1, a2)
        for p_ in mg.noise_t.parameters():
            p_.mul_(1.5)                                                               # in-place update -> new version
        b1 = snsde_b200.sdeint(mg, y0d, td, dt=1.0, seed=5, precision="tc")
        b_ref = snsde_b200.sdeint(mg, y0d, td, dt=1.0, seed=5, precision="fp32")       # FMA kernel: no table at all
        assert not torch.equal(a1, b1)
        close(b1, b_ref)
        # two streams, one plan, different step grids back to back
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        torch.cuda.synchronize()
        outs = []
        for i in range(6):
            with torch.cuda.stream(s1 if i % 2 == 0 else s2):
                ts = td if i % 2 == 0 else td[::2].contiguous()
                outs.append((i, snsde_b200.sdeint(mg, y0d, ts, dt=1.0, seed=5, precision="tc")))
        torch.cuda.synchronize()
        want_full = snsde_b200.sdeint(mg, y0d, td, dt=1.0, seed=5, precision="tc")
        for i, z in outs:
            assert torch.equal(z, want_full if i % 2 == 0 else want_full[::2])


def test_m_split_cta_pairs_agree_with_the_single_cta_kernel(dev, monkeypatch):
    """hidden 256: the cluster launch (CTA pairs, each owning 128 output features, operand halves exchanged through
    DSMEM) against the single-CTA general kernel (streamed weights) and the oracle; ragged batch (odd number of pairs'
    rows), control model, in-kernel Philox."""
    for (io, no, B, C, L) in ((4, 17, 37, 14, 1), (3, 6, 20, 5, 1), (6, 17, 9, 7, 1)):
        m, times, coeffs, y0 = make_problem(io, no, B, 256, C, L, 14, seed=500 + io)
        mg = m.to(dev)
        mg.set_X(coeffs.to(dev), times.to(dev))
        with torch.no_grad():
            a = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=21, precision="tc")
            plan = snsde_b200.plans_of(mg)[("euler", "tc", str(dev))]
            assert plan.kernel == "tcgen05_general"
            dW = snsde_b200.philox_increments(21, plan.step_plan(times, 1.0, times), B, 256, dev).cpu()
            monkeypatch.setenv("SNSDE_TCG_NO_MSPLIT", "1")
            b = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=21, precision="tc")
            monkeypatch.delenv("SNSDE_TCG_NO_MSPLIT")
        m.to("cpu"); m.set_X(coeffs, times)
        want = solver.sdeint(m, y0, times, 1.0, solver.BrownianTable(dW))
        close(a, want)
        close(b, want)
        close(a, b, rtol=2e-5)


def test_natural_spline_with_missing_values_on_device_matches_the_reference_golden(golden_dir, dev):
    """snsde_natural_coeffs_missing vs coefficients minted by the reference's in-tree builder (interpolate.py:56-153):
    NaNs inside, at either end, a whole series missing; plus a larger random case against the pinned oracle."""
    from snsde_b200 import data
    cases = [c for c in torch.load(golden_dir / "spline_golden.pt") if c.get("missing")]
    assert len(cases) == 2
    for c in cases:
        want = torch.cat([c["a"], c["b"], c["two_c"], c["three_d"]], dim=-1)
        got = data.natural_coeffs_cuda(c["x"].to(dev), c["times"].to(dev), missing=True)
        close(got, want, rtol=2e-5)
    g = torch.Generator().manual_seed(9)
    B, K, C = 33, 40, 5
    times = torch.cat([torch.zeros(1), torch.rand(K - 1, generator=g) + 0.1]).cumsum(0)
    x = torch.randn(B, K, C, generator=g).cumsum(1)
    x[torch.rand(B, K, C, generator=g) < 0.4] = float("nan")
    x[3, :, 2] = float("nan")
    want = torch.cat(spline.natural_cubic_spline_coeffs(times, x), dim=-1)
    got = data.natural_coeffs_cuda(x.to(dev), times.to(dev), missing=True)
    close(got, want, rtol=5e-5)
    full = torch.randn(4, 9, 3, generator=g)                     # no NaN: the missing-value kernel equals the plain builder
    close(data.natural_coeffs_cuda(full.to(dev), times[:9].to(dev), missing=True),
          torch.cat(spline.natural_cubic_spline_coeffs(times[:9], full), dim=-1), rtol=2e-5)


# ---- the warp-shuffle kernel (hidden <= 32, csrc/snsde_warp.cu) against the interpreter kernel ----
def _solve_with_variant(m, coeffs, times, y0, ts, dt, method, dev, warp, monkeypatch, **kw):
    """A fresh plan with the warp-shuffle form enabled / disabled (SNSDE_NO_WARP is read when the weights are set)."""
    if warp:
        monkeypatch.delenv("SNSDE_NO_WARP", raising=False)
    else:
        monkeypatch.setenv("SNSDE_NO_WARP", "1")
    snsde_b200.engine._PLANS.pop(m, None)
    with torch.no_grad():
        out = snsde_b200.sdeint(m, y0, ts, dt=dt, method=method, precision="fp32", **kw)
    plan = next(iter(snsde_b200.plans_of(m).values()))
    variant = plan.variant
    snsde_b200.engine._PLANS.pop(m, None)
    monkeypatch.delenv("SNSDE_NO_WARP", raising=False)
    return out, variant


WARP_CASES = [  # family, io, no, H, HH, C, L, B, method
    ("tutorial", 0, 0, 32, 32, 2, 1, 64, "euler"),            # BASELINE c1's function
    ("tutorial", 0, 0, 20, 28, 3, 2, 7, "euler"),              # ragged widths, 7 mat-vecs
    ("benchmark", 4, 17, 32, 32, 5, 1, 19, "euler"), ("benchmark", 6, 17, 32, 32, 7, 1, 1500, "milstein"),
    ("benchmark", 2, 16, 16, 16, 4, 2, 33, "euler"), ("benchmark", 3, 18, 32, 32, 3, 1, 40, "euler"),
    ("benchmark", 1, 19, 24, 30, 3, 2, 9, "euler"), ("benchmark", 0, 5, 8, 8, 32, 1, 5, "milstein"),
    ("benchmark", 5, 9, 32, 17, 3, 3, 1, "milstein"), ("benchmark", 4, 13, 31, 31, 6, 1, 2400, "euler"),     # several pairs per CTA
    ("benchmark", 1, 14, 32, 32, 3, 1, 12, "euler"), ("benchmark", 3, 3, 5, 9, 2, 4, 3, "euler"),
    ("benchmark", 1, 18, 32, 32, 3, 1, 1300, "euler"),
    # SRK: six passes per step through the same mat-vec code (drift / diffusion at the stage states)
    ("benchmark", 4, 17, 32, 32, 5, 1, 19, "srk"), ("benchmark", 3, 18, 32, 32, 3, 1, 40, "srk"), ("benchmark", 1, 19, 24, 30, 3, 2, 9, "srk"),
    ("benchmark", 6, 6, 16, 16, 4, 2, 300, "srk"), ("tutorial", 0, 0, 32, 32, 2, 1, 64, "srk"), ("benchmark", 5, 0, 8, 12, 3, 1, 5, "srk"),
]


@pytest.mark.parametrize("family,io,no,H,HH,C,L,B,method", WARP_CASES)
def test_warp_kernel_agrees_with_the_interpreter_and_the_oracle(family, io, no, H, HH, C, L, B, method, dev, monkeypatch):
    K = 12
    m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, K, seed=7 * H + io + no, HH=HH, family=family, spacing=0.5)
    if no == 7:
        y0 = y0.abs() + 0.5
    dt = 0.5
    ts = torch.cat([times[:1], times[3:4], (times[5:6] + times[6:7]) / 2, times[-1:]])
    dW = torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(3)) * dt ** 0.5
    dU = dt * (dW / 2 + torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(4)) * (dt / 12) ** 0.5)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    args = (mg, coeffs, times.to(dev), y0.to(dev), ts.to(dev), dt, method, dev)
    bm = snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev))
    a, va = _solve_with_variant(*args, True, monkeypatch, bm=bm)
    b, vb = _solve_with_variant(*args, False, monkeypatch, bm=bm)
    assert vb == "interpreter" and va == "warp", (va, vb)
    close(a, b, rtol=5e-6)                      # same arithmetic; four partial sums per output instead of one
    # Philox mode: the two forms draw the same stream
    pa, _ = _solve_with_variant(*args, True, monkeypatch, seed=21)
    pb, _ = _solve_with_variant(*args, False, monkeypatch, seed=21)
    close(pa, pb, rtol=5e-6)
    if B <= 64:
        m.to("cpu"); m.set_X(coeffs, times)
        want = solver.sdeint(m, y0, ts, dt, solver.BrownianTable(dW, dU=dU), method=method)
        close(a, want)


def test_warp_kernel_fused_final_index_and_row_offset(dev, monkeypatch):
    """Per-row capture and batch sharding on the warp-shuffle form: shards reproduce the full batch bit for bit."""
    B, H, C, L, K = 70, 32, 4, 1, 10
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=5)
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    fi = torch.randint(1, K, (B,), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        full = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), y0.to(dev), seed=9, precision="fp32")
        assert next(iter(snsde_b200.plans_of(mg).values())).variant == "warp"
        stream = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=9, precision="fp32")
        parts = []
        for lo, hi in ((0, 24), (24, 70)):
            mg.set_X(coeffs[lo:hi].to(dev), times.to(dev))
            parts.append(snsde_b200.solve_final(mg, times.to(dev), fi[lo:hi].to(dev), y0[lo:hi].to(dev), seed=9,
                                                precision="fp32", row_offset=lo))
    want = stream[fi.to(dev), torch.arange(B, device=dev)]
    assert torch.equal(full, want)
    assert torch.equal(torch.cat(parts), full)


@pytest.mark.parametrize("precision", ["fp32"])
def test_cached_coefficient_table_follows_weights_and_step_times(dev, precision, monkeypatch):
    """The row-independent diffusion coefficient (noise_t) is tabulated once per (weights, evaluation times) on the fp32
    kernels: a solve after an in-place weight update, or on another time grid, must not read the old table."""
    for warp in (True, False):
        if warp:
            monkeypatch.delenv("SNSDE_NO_WARP", raising=False)
        else:
            monkeypatch.setenv("SNSDE_NO_WARP", "1")
        B, H, C, L, K = 9, 32, 4, 1, 8
        m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=31)
        mg = m.to(dev)
        mg.set_X(coeffs.to(dev), times.to(dev))
        dW = (torch.randn(K - 1, B, H, generator=torch.Generator().manual_seed(4))).to(dev)
        bm = lambda: snsde_b200.BrownianIncrements(dW)                       # noqa: E731
        with torch.no_grad():
            z1 = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=bm(), precision=precision)
            z1b = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=bm(), precision=precision)   # table reused
            assert torch.equal(z1, z1b)
            mg.noise_t[2].bias.add_(0.7)                                      # in-place update (an optimizer step)
            z2 = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=bm(), precision=precision)
            snsde_b200.engine._PLANS.pop(mg, None)                            # fresh plan, fresh table
            z2f = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, bm=bm(), precision=precision)
            assert not torch.equal(z1, z2) and torch.equal(z2, z2f)
            ts = times[:5]
            z3 = snsde_b200.sdeint(mg, y0.to(dev), ts.to(dev), dt=1.0, bm=snsde_b200.BrownianIncrements(dW[:4]), precision=precision)
            assert torch.equal(z3, z2f[:5])
            half = times.to(dev) * 0.5                                        # other evaluation times, same step count
            mg.set_X(coeffs.to(dev), half)
            z4 = snsde_b200.sdeint(mg, y0.to(dev), half, dt=0.5, bm=bm(), precision=precision)
            snsde_b200.engine._PLANS.pop(mg, None)
            z4f = snsde_b200.sdeint(mg, y0.to(dev), half, dt=0.5, bm=bm(), precision=precision)
            assert torch.equal(z4, z4f)
        snsde_b200.engine._PLANS.pop(mg, None)


def test_large_launches_of_small_models_run_on_the_interpreter_and_agree_with_the_warp_kernel(dev):
    """hidden <= 32 is eligible for the warp-owned kernel, which serves launches of up to 4096 rows; rows never interact,
    so the first rows of a 5000-row launch (interpreter kernel) equal the same rows solved alone (warp-owned kernel)."""
    B, H, C, L, K = 5000, 32, 3, 1, 7
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=77)
    mg = m.to(dev)
    with torch.no_grad():
        mg.set_X(coeffs.to(dev), times.to(dev))
        big = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=1.0, seed=5, precision="fp32")
        assert next(iter(snsde_b200.plans_of(mg).values())).variant == "warp"          # eligibility is a plan property
        mg.set_X(coeffs[:96].to(dev), times.to(dev))
        small = snsde_b200.sdeint(mg, y0[:96].to(dev), times.to(dev), dt=1.0, seed=5, precision="fp32")
    close(big[:, :96], small, rtol=5e-6)
    m.to("cpu"); m.set_X(coeffs[:16], times)
    plan = next(iter(snsde_b200.plans_of(mg).values()))
    dW = snsde_b200.philox_increments(5, plan.step_plan(times, 1.0, times), 16, H, dev).cpu()
    close(big[:, :16], solver.sdeint(m, y0[:16], times, 1.0, solver.BrownianTable(dW)))


def test_warp_kernel_randomised_shapes_against_the_interpreter(dev, monkeypatch):
    """Seeded sweep over small models at the edges of the warp-owned kernel's envelope (width 1, 32 channels, 5 layers,
    single rows, zero and one step, every method): both fp32 kernels must agree, and the plan must pick the warp form."""
    rng = np.random.default_rng(123)
    n_checked = 0
    for trial in range(36):
        io = int(rng.integers(0, 7))
        no = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19]))
        method = ["euler", "milstein", "srk"][trial % 3]
        if method == "milstein" and no in (14, 15, 18, 19):
            method = "euler"                               # full vjp through noise_y: interpreter only (not this test)
        H = int(rng.choice([1, 2, 3, 7, 16, 31, 32]))
        HH = H if io in (0, 2, 4, 6) else int(rng.choice([1, 5, 32]))
        C = int(rng.choice([1, 2, 9, 32]))
        L = int(rng.choice([1, 2, 5]))
        B = int(rng.choice([1, 2, 3, 65, 130]))
        K = int(rng.choice([1, 2, 6]))
        m, times, coeffs, y0 = make_problem(io, no, B, H, C, L, max(K, 2), seed=1000 + trial, HH=HH, spacing=0.5)
        ts = times[:K]
        S = K - 1
        g = torch.Generator().manual_seed(trial)
        dW = torch.randn(S, B, H, generator=g) * 0.5 ** 0.5
        dU = 0.5 * (dW / 2 + torch.randn(S, B, H, generator=g) * (0.5 / 12) ** 0.5)
        mg = m.to(dev)
        mg.set_X(coeffs.to(dev), times.to(dev))
        args = (mg, coeffs, times.to(dev), y0.to(dev), ts.to(dev), 0.5, method, dev)
        for kw in (dict(bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev))), dict(seed=trial)):
            a, va = _solve_with_variant(*args, True, monkeypatch, **kw)
            b, vb = _solve_with_variant(*args, False, monkeypatch, **kw)
            assert (va, vb) == ("warp", "interpreter"), (trial, io, no, method, H, HH, C, L, va, vb)
            assert a.shape == (K, B, H) and torch.isfinite(a).all()
            assert torch.equal(a[0].cpu(), y0)
            close(a, b, rtol=5e-6)
            n_checked += 1
    assert n_checked == 72


@pytest.mark.parametrize("method,K", [("euler", 2700), ("srk", 950)])
def test_warp_kernel_long_trajectories_read_their_tables_from_global_memory(method, K, dev, monkeypatch):
    """Step / emit / point tables beyond the shared-memory budget (96 KB) stay in global memory: same results."""
    B, H, C, L = 3, 8, 2, 1
    m, times, coeffs, y0 = make_problem(4, 17, B, H, C, L, K, seed=9, spacing=0.01)
    dt = 0.01
    mg = m.to(dev)
    mg.set_X(coeffs.to(dev), times.to(dev))
    ts = times[[0, 7, K // 2, K - 1]]
    args = (mg, coeffs, times.to(dev), y0.to(dev), ts.to(dev), dt, method, dev)
    a, va = _solve_with_variant(*args, True, monkeypatch, seed=3)
    b, vb = _solve_with_variant(*args, False, monkeypatch, seed=3)
    assert (va, vb) == ("warp", "interpreter")
    assert torch.isfinite(a).all()
    close(a, b, rtol=2e-5)                     # thousands of steps of fp32 re-association

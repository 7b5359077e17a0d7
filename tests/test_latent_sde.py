"""LatentSDE (rest): the augmented system f_aug / g_aug of
/root/reference/torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:29-147 (posterior drift + KL path accumulator,
``sdeint_adjoint(..., names={'drift': 'f_aug', 'diffusion': 'g_aug'})``, default method 'srk').

CPU part: the oracle against the goldens minted from the reference's own class, the host logic (describe / pack /
weight count / patch dispatch, also on the real class when /root/reference is mounted).
GPU part (``-m gpu``): the CUDA path through the C ABI against those goldens and the oracle - forward (euler,
milstein, srk; Philox replay), the patched ``forward`` (out, latent, logqp) and the backward pass (euler).
Tolerances: forward 1e-4 of the tensor scale (north star); gradients 1e-4 of each gradient tensor's max norm.
"""
import copy
import ctypes
import pathlib
import sys

import pytest
import torch

import snsde_b200
from oracle import latent as olatent
from oracle import solver, spline
from snsde_b200 import _lib, packing
from snsde_b200.modules import LatentSDEParams

ROOT = pathlib.Path(__file__).resolve().parents[1]
REF = pathlib.Path("/root/reference")


@pytest.fixture(scope="module")
def cases(golden_dir):
    return torch.load(golden_dir / "latent_golden.pt")


def _oracle_model(c):
    if c["kind"] == "fg":
        C, H, HH, L = c["dims"]
    else:
        _, _, C, H, HH, L = c["dims"]
    th, mu, sg = c["prior"]
    m = olatent.LatentSDE(C, H, HH, L, theta=th, mu=mu, sigma=sg)
    m.load_state_dict(c["state_dict"])          # strict: same names, shapes, buffers as the reference class
    return m


def close(got, want, rtol=1e-4, what=""):
    got, want = got.detach().cpu().float(), want.detach().cpu().float()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(float(want.abs().max()), 1.0)
    err = float((got - want).abs().max())
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} > {rtol:g} * {scale:.3g}"


# ---- CPU: oracle pinned by the reference's own class ------------------------------------------------------------
def test_oracle_f_aug_g_aug_match_the_reference_class(cases):
    fg = [c for c in cases if c["kind"] == "fg"]
    assert len(fg) == 4
    with torch.no_grad():
        for c in fg:
            m = _oracle_model(c)
            for i, t in enumerate(c["t"]):
                f, g = m.f_aug(t, c["y"]), m.g_aug(t, c["y"])
                assert torch.allclose(f, c["f_aug"][i], rtol=1e-6, atol=1e-6 * float(c["f_aug"][i].abs().max()))
                assert torch.equal(g, c["g_aug"][i])
                assert float(g[:, -1].abs().max()) == 0.0 and torch.all(g[:, :-1] == c["prior"][2])


def test_oracle_forward_matches_the_reference_class(cases):
    fw = [c for c in cases if c["kind"] == "forward"]
    assert [c["method"] for c in fw] == [None, "euler", "euler", "srk", "milstein"]
    with torch.no_grad():
        for c in fw:
            m = _oracle_model(c)
            pred, lat, logqp = m(c["coeffs"], c["times"], bm=solver.BrownianTable(c["dW"], dU=c["dU"]), method=c["method"])
            close(lat, c["latent"], 1e-6, "latent")
            close(pred, c["pred"], 1e-6, "pred")
            close(logqp, c["logqp"], 1e-6, "logqp")
            assert float(c["logqp"]) > 0.0


def test_oracle_kl_channel_is_half_the_squared_drift_mismatch():
    """The extra channel integrates 0.5 |(f - h) / sigma|^2 and carries no noise: under Euler its final value is the
    left Riemann sum of that quantity along the latent path (closed-form check of the restatement)."""
    torch.manual_seed(3)
    m = olatent.LatentSDE(2, 6, 7, 2, theta=0.8, mu=0.3, sigma=0.6)
    B, K = 3, 6
    times = torch.linspace(0, 1, K)
    y0 = torch.cat([torch.randn(B, 5), torch.zeros(B, 1)], 1)
    steps = solver.step_times(times, solver.solver_dt(times))
    dW = torch.randn(len(steps), B, 6) * 0.3
    with torch.no_grad():
        ys = solver.sdeint(m, y0, torch.tensor([t for t, _ in steps] + [steps[-1][1]]), solver.solver_dt(times),
                           solver.BrownianTable(dW), method="euler", names=olatent.NAMES)
        acc = torch.zeros(B)
        for k, (t0, t1) in enumerate(steps):
            lat = ys[k][:, :-1]
            u = (m.f(torch.tensor(t0), lat) - 0.8 * (0.3 - lat)) / 0.6
            acc = acc + torch.tensor(t1 - t0) * 0.5 * (u ** 2).sum(1)
    assert torch.allclose(ys[-1][:, -1], acc, rtol=1e-5, atol=1e-6)


# ---- CPU: host logic --------------------------------------------------------------------------------------------
def test_describe_pack_and_weight_count():
    m = LatentSDEParams(3, 9, 12, 3, theta=0.7, mu=0.2, sigma=0.3)
    desc = packing.describe(m)
    assert desc == dict(family=_lib.FAMILY_LATENT_SDE, input_option=0, noise_option=0, input_channels=3, hidden=9,
                        hidden_hidden=12, num_hidden_layers=3)
    blob = packing.pack(m, desc)
    assert blob.numel() == 12 * 10 + 12 + 2 * (12 * 12 + 12) + 8 * 12 + 8 + 3
    assert blob[-3:].tolist() == pytest.approx([0.7, 0.2, 0.3])
    cd = _lib.ModelDesc(method=2, precision=0, **desc)
    assert _lib.load().snsde_weight_count(ctypes.byref(cd)) == blob.numel()
    keys, gkeys = packing.blob_keys(desc), packing.grad_keys(desc)
    assert keys[:len(gkeys)] == gkeys and keys[len(gkeys):] == ["theta", "mu", "sigma"]
    assert all(k in dict(m.named_parameters()) for k in gkeys)
    # same state_dict layout as the oracle restatement (and through it, the reference class)
    o = olatent.LatentSDE(3, 9, 12, 3, theta=0.7, mu=0.2, sigma=0.3)
    assert [(k, tuple(v.shape)) for k, v in o.state_dict().items()] == [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    bad = _lib.ModelDesc(family=_lib.FAMILY_LATENT_SDE, input_option=0, noise_option=0, input_channels=1, hidden=1,
                         hidden_hidden=4, num_hidden_layers=1, method=0, precision=0)
    assert _lib.load().snsde_weight_count(ctypes.byref(bad)) == _lib.ERR_BAD_ARG


def test_patch_replaces_forward_of_a_latent_sde_and_sdeint_requires_names():
    from snsde_b200 import engine
    m = snsde_b200.patch(LatentSDEParams(2, 5, 6, 1))
    assert snsde_b200.wrapper_kind(m) == "latent_sde" and m.forward.__func__ is engine._forward_latent
    dup = copy.deepcopy(m)
    assert dup.forward.__self__ is dup
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(snsde_b200.EngineError):
            m(torch.zeros(2, 3, 8), torch.arange(4.0))
        with torch.no_grad(), pytest.raises(snsde_b200.EngineError):       # no plan without a device: refusal comes first
            snsde_b200.sdeint(m, torch.zeros(2, 5), torch.arange(4.0), dt=1.0)


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted (GPU box)")
def test_the_reference_own_latent_sde_is_recognised_and_packs_identically():
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    try:
        import make_golden as mg
    finally:
        sys.path.pop(0)
    mg.install_shims()
    try:
        ref = mg.load_latent_module()
        torch.manual_seed(0)
        real = ref.LatentSDE(3, 9, 12, 3, theta=0.7, mu=0.2, sigma=0.3)
        desc = packing.describe(real)
        assert desc["family"] == _lib.FAMILY_LATENT_SDE and (desc["hidden"], desc["hidden_hidden"], desc["num_hidden_layers"]) == (9, 12, 3)
        own = LatentSDEParams(3, 9, 12, 3, theta=0.7, mu=0.2, sigma=0.3)
        own.load_state_dict(real.state_dict())
        assert torch.equal(packing.pack(real, desc), packing.pack(own, packing.describe(own)))
        from snsde_b200 import engine
        patched = snsde_b200.patch(real)
        assert snsde_b200.wrapper_kind(patched) == "latent_sde" and patched.forward.__func__ is engine._forward_latent
        if not torch.cuda.is_available():
            with torch.no_grad(), pytest.raises(snsde_b200.EngineError):
                patched(torch.zeros(2, 3, 12), torch.arange(4.0))
    finally:
        for name in ("torchcde", "torchsde", "torchdiffeq", "controldiffeq"):
            sys.modules.pop(name, None)


# ---- GPU: the CUDA path through the C ABI -------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _engine_model(c, dev):
    _, _, C, H, HH, L = c["dims"]
    th, mu, sg = c["prior"]
    m = LatentSDEParams(C, H, HH, L, theta=th, mu=mu, sigma=sg)
    m.load_state_dict(c["state_dict"])
    return snsde_b200.patch(m.to(dev))


@pytest.mark.gpu
def test_patched_forward_reproduces_the_reference_goldens(cases, dev):
    """(out, latent, logqp) of the reference class's forward, for the default 'srk', 'euler' and 'milstein'."""
    for c in [c for c in cases if c["kind"] == "forward"]:
        model = _engine_model(c, dev)
        kw = {} if c["method"] is None else {"method": c["method"]}
        with torch.no_grad():
            pred, lat, logqp = model(c["coeffs"].to(dev), c["times"].to(dev),
                                     bm=snsde_b200.BrownianIncrements(c["dW"].to(dev), c["dU"].to(dev)), **kw)
        plan = next(iter(snsde_b200.plans_of(model).values()))
        assert plan.kernel == "fma_fp32" and plan.variant == "warp" and plan.desc["family"] == _lib.FAMILY_LATENT_SDE
        close(lat, c["latent"], 1e-4, f"latent {c['method']}")
        close(pred, c["pred"], 1e-4, f"pred {c['method']}")
        close(logqp, c["logqp"], 1e-4, f"logqp {c['method']}")


LATENT_SHAPES = [  # C, H, HH, L, B, K, method       one warp / several warps per row group, ragged batch, deep MLP
    (2, 5, 6, 1, 7, 6, "euler"), (2, 5, 6, 1, 7, 6, "srk"), (3, 32, 32, 2, 33, 7, "milstein"), (3, 33, 40, 2, 9, 7, "euler"), (4, 65, 64, 1, 12, 6, "srk"), (2, 32, 32, 3, 64, 9, "srk"),
    (3, 17, 130, 2, 5, 6, "milstein"), (2, 129, 96, 1, 300, 5, "euler"), (2, 8, 8, 1, 1030, 5, "srk"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,HH,L,B,K,method", LATENT_SHAPES)
def test_augmented_solve_matches_the_oracle(C, H, HH, L, B, K, method, dev, monkeypatch):
    torch.manual_seed(H * 7 + B)
    m = olatent.LatentSDE(C, H, HH, L, theta=0.9, mu=-0.2, sigma=0.45)
    times = torch.linspace(0, 1.5, K)
    dt = solver.solver_dt(times)
    steps = solver.step_times(times, dt)
    h = torch.tensor([b - a for a, b in steps]).view(-1, 1, 1)
    dW = torch.randn(len(steps), B, H) * h.sqrt()
    dU = h * (dW / 2 + torch.randn(len(steps), B, H) * (h / 12).sqrt())
    y0 = torch.cat([torch.randn(B, H - 1) * 0.7, torch.zeros(B, 1)], 1)
    want = solver.sdeint(m, y0, times, dt, solver.BrownianTable(dW, dU=dU), method=method, names=olatent.NAMES)
    mg = LatentSDEParams(C, H, HH, L)
    mg.load_state_dict(m.state_dict())
    mg = mg.to(dev)
    with torch.no_grad():
        got = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt, method=method, names=olatent.NAMES,
                                bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev)))
    close(got, want, 1e-4, "augmented states")
    close(got[..., -1], want[..., -1], 1e-4, "KL accumulator")
    plan = next(iter(snsde_b200.plans_of(mg).values()))
    assert plan.variant == ("warp" if max(H, HH) <= 32 else "interpreter")
    if plan.variant == "warp":                              # the same solve on the interpreter kernel
        monkeypatch.setenv("SNSDE_NO_WARP", "1")
        snsde_b200.engine._PLANS.pop(mg, None)
        with torch.no_grad():
            ref = snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt, method=method, names=olatent.NAMES,
                                    bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev)))
        assert next(iter(snsde_b200.plans_of(mg).values())).variant == "interpreter"
        monkeypatch.delenv("SNSDE_NO_WARP")
        snsde_b200.engine._PLANS.pop(mg, None)
        close(got, ref, 5e-6, "warp-owned vs interpreter kernel")
    assert float(want[-1, :, -1].min()) > 0.0
    with pytest.raises(ValueError, match="augmented"):
        snsde_b200.sdeint(mg, y0.to(dev), times.to(dev), dt=dt)


@pytest.mark.gpu
def test_philox_replay_is_bit_identical_and_the_kl_channel_sees_no_noise(dev):
    torch.manual_seed(5)
    C, H, HH, L, B, K = 2, 16, 24, 2, 33, 8
    mg = LatentSDEParams(C, H, HH, L).to(dev)
    times = torch.linspace(0, 1, K, device=dev)
    dt = solver.solver_dt(times.cpu())
    y0 = torch.cat([torch.randn(B, H - 1), torch.zeros(B, 1)], 1).to(dev)
    for method in ("euler", "srk"):
        with torch.no_grad():
            a = snsde_b200.sdeint(mg, y0, times, dt=dt, method=method, names=olatent.NAMES, seed=77)
            plan = snsde_b200.plan_for(mg, method, "auto", dev)
            sp = plan.step_plan(times, dt, None)
            dW, dU = snsde_b200.philox_increments(77, sp, B, H, dev, with_U=True)
            b = snsde_b200.sdeint(mg, y0, times, dt=dt, method=method, names=olatent.NAMES,
                                  bm=snsde_b200.BrownianIncrements(dW, dU))
            noisy = dW.clone(); noisy[..., -1] += 3.0          # increments of the KL channel must not matter (g_aug = 0 there)
            c = snsde_b200.sdeint(mg, y0, times, dt=dt, method=method, names=olatent.NAMES,
                                  bm=snsde_b200.BrownianIncrements(noisy, dU))
        assert torch.equal(a, b), method
        if method == "euler":
            assert torch.equal(b, c)
        else:
            close(c, b, 1e-6, "srk with perturbed KL-channel increments")


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,HH,L,B,K,method", [
    (3, 9, 12, 2, 7, 6, "euler"), (2, 33, 40, 1, 10, 5, "euler"), (2, 70, 64, 3, 6, 5, "euler"),
    (3, 9, 12, 2, 7, 6, "srk"), (2, 33, 40, 1, 10, 5, "srk"),
    # enough rows for 4-row groups in the reverse sweeps (one and two warps per group), ragged last group
    (2, 9, 12, 1, 330, 5, "euler"), (2, 9, 12, 1, 330, 5, "srk"), (3, 33, 40, 1, 161, 4, "srk")])
def test_backward_matches_autograd_through_the_oracle(C, H, HH, L, B, K, method, dev):
    """dL/d(aug_y0) and dL/d(every drift parameter) with L mixing latent states and the KL accumulator."""
    torch.manual_seed(40 + H)
    m = olatent.LatentSDE(C, H, HH, L, theta=1.1, mu=0.15, sigma=0.5)
    times = torch.linspace(0, 1, K)
    dt = solver.solver_dt(times)
    steps = solver.step_times(times, dt)
    S = len(steps)
    h = torch.tensor([b - a for a, b in steps]).view(S, 1, 1)
    dW = torch.randn(S, B, H) * h.sqrt()
    dU = h * (dW / 2 + torch.randn(S, B, H) * (h / 12).sqrt())
    y0 = torch.cat([torch.randn(B, H - 1) * 0.5, torch.zeros(B, 1)], 1)
    w = torch.randn(K, B, H)
    mo = copy.deepcopy(m).double()
    y0o = y0.double().requires_grad_(True)
    zo = solver.sdeint_with_grad(mo, y0o, times.double(), dt, solver.BrownianTable(dW.double(), dU=dU.double()), method=method,
                                 names=olatent.NAMES)
    ((zo * w.double()).sum() + 3.0 * zo[-1, :, -1].mean()).backward()
    mg = LatentSDEParams(C, H, HH, L)
    mg.load_state_dict(m.state_dict())
    mg = mg.to(dev)
    y0g = y0.to(dev).requires_grad_(True)
    zg = snsde_b200.sdeint(mg, y0g, times.to(dev), dt=dt, method=method, names=olatent.NAMES,
                           bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev)))
    close(zg, zo.float(), 1e-4, "states under autograd")
    ((zg * w.to(dev)).sum() + 3.0 * zg[-1, :, -1].mean()).backward()

    def gclose(got, want, name):
        scale = max(float(want.abs().max()), 1e-6)
        err = float((got.detach().cpu().double() - want).abs().max())
        assert err <= 1e-4 * scale, f"{name}: grad max abs err {err:.3e} > 1e-4 * {scale:.3g}"
    gclose(y0g.grad, y0o.grad, "aug_y0")
    named_o = dict(mo.named_parameters())
    checked = 0
    for name, p in mg.named_parameters():
        if name.split(".")[0] in ("linear_in", "linears", "linear_out"):
            gclose(p.grad, named_o[name].grad, name)
            checked += 1
    assert checked == 2 * (L + 1)


@pytest.mark.gpu
def test_training_step_through_the_patched_latent_forward(dev):
    """loss = mse(out) + logqp, backward through the engine (method 'euler'): the gradients of the drift network, the
    initial network, the embedding and q(y0) equal autograd through the oracle forward on identical increments."""
    torch.manual_seed(9)
    B, K, C, H, HH, L = 6, 7, 3, 10, 16, 2
    m = olatent.LatentSDE(C, H, HH, L, theta=1.0, mu=0.0, sigma=0.5)
    times = torch.linspace(0, 1, K)
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(torch.randn(B, K, C).cumsum(1) * 0.3, times)
    S = len(solver.step_times(times, solver.solver_dt(times)))
    dW = torch.randn(S, B, H) * solver.solver_dt(times) ** 0.5
    target = torch.randn(B, K, H)
    mo = copy.deepcopy(m).double()
    pred, _, logqp = mo(coeffs.double(), times.double(), bm=solver.BrownianTable(dW.double()), method="euler", with_grad=True)
    ((pred - target.double()).pow(2).mean() + 0.1 * logqp).backward()
    mg = LatentSDEParams(C, H, HH, L)
    mg.load_state_dict(m.state_dict())
    mg = snsde_b200.patch(mg.to(dev)).train()
    predg, latg, logqpg = mg(coeffs.to(dev), times.to(dev), bm=snsde_b200.BrownianIncrements(dW.to(dev)), method="euler")
    close(predg, pred.float(), 1e-4, "pred")
    close(logqpg, logqp.float(), 1e-4, "logqp")
    ((predg - target.to(dev)).pow(2).mean() + 0.1 * logqpg).backward()
    named_o = dict(mo.named_parameters())
    for name, p in mg.named_parameters():
        want = named_o[name].grad
        assert p.grad is not None and want is not None, name
        scale = max(float(want.abs().max()), 1e-6)
        err = float((p.grad.cpu().double() - want).abs().max())
        assert err <= 1e-4 * scale, f"{name}: {err:.3e} vs scale {scale:.3g}"
    # the reference's default method ('srk', latent_sde.py:107-109) trains too: reverse sweep of the SRK solve
    # (increments scaled per step: the linspace grid ends in a sliver step of ~1e-7, where an O(sqrt(dt)) increment would make
    # the SRK weights I_kkk / h cancel catastrophically in fp32 - in torchsde as much as here)
    h = torch.tensor([b - a for a, b in solver.step_times(times, solver.solver_dt(times))]).view(-1, 1, 1)
    dW = torch.randn(S, B, H) * h.sqrt()
    dU = h * (dW / 2 + torch.randn(S, B, H) * (h / 12).sqrt())
    mo2 = copy.deepcopy(m).double()
    pred, _, logqp = mo2(coeffs.double(), times.double(), bm=solver.BrownianTable(dW.double(), dU=dU.double()), with_grad=True)
    ((pred - target.double()).pow(2).mean() + 0.1 * logqp).backward()
    mg.zero_grad(set_to_none=True)
    predg, _, logqpg = mg(coeffs.to(dev), times.to(dev), bm=snsde_b200.BrownianIncrements(dW.to(dev), dU.to(dev)))
    close(predg, pred.float(), 1e-4, "pred (srk)")
    ((predg - target.to(dev)).pow(2).mean() + 0.1 * logqpg).backward()
    named_o = dict(mo2.named_parameters())
    for name, p in mg.named_parameters():
        want = named_o[name].grad
        scale = max(float(want.abs().max()), 1e-6)
        err = float((p.grad.cpu().double() - want).abs().max())
        assert err <= 1e-4 * scale, f"srk {name}: {err:.3e} vs scale {scale:.3g}"

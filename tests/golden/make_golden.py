"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is real reference code here and what is shim:

* REAL: ``Diffusion_model`` / ``NeuralSDE`` / ``NeuralSDE_forecasting`` from
  /root/reference/benchmark_{classification,forecasting}/models_sde/neuralsde.py and
  the in-tree spline /root/reference/benchmark_classification/controldiffeq/interpolate.py,
  imported unmodified.
* SHIM: the reference imports three packages that do not exist in this image
  (torchsde, torchcde, torchdiffeq).  ``torchcde.CubicSpline`` /
  ``torchcde.hermite_cubic_coefficients_with_backward_differences`` and
  ``torchsde.sdeint`` are provided by the oracle; ``torchdiffeq`` is an empty stub
  (only the CDE baselines use it).  So ``fg_golden.pt`` pins f/g *given* the oracle's
  spline read, ``spline_golden.pt`` pins that spline read against the reference's own
  in-tree spline, and ``forward_golden.pt`` pins the reference's output-time
  selection/gather given the oracle solver.

  ``forward_golden.pt`` also holds the torch-ists wrapper (nsde_model.py, loaded by file path as the
  reference's own test does), whose default method is 'srk'.

  ``latent_golden.pt`` holds the reference's own ``LatentSDE`` (latent_sde.py, loaded by file path; ``torchsde.SDEIto``
  shimmed as an nn.Module base, ``sdeint_adjoint(..., names=...)`` provided by the oracle solver): f_aug / g_aug values
  and full forward outputs ``(out, latent, logqp)``.

Outputs: fg_golden.pt, spline_golden.pt, forward_golden.pt, latent_golden.pt (a few hundred KB in total).
"""
import importlib
import pathlib
import sys
import types

import torch

HERE = pathlib.Path(__file__).resolve().parent
REPO = HERE.parents[1]
REF = pathlib.Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle import solver as osolver          # noqa: E402
from oracle import spline as ospline          # noqa: E402


def install_shims():
    tcde = types.ModuleType("torchcde")
    tcde.CubicSpline = ospline.CubicSpline
    tcde.hermite_cubic_coefficients_with_backward_differences = (
        ospline.hermite_cubic_coefficients_with_backward_differences)
    tsde = types.ModuleType("torchsde")

    def sdeint(sde, y0, ts, dt, bm=None, method="euler", options=None, names=None, **kw):
        return osolver.sdeint(sde, y0, ts, dt, bm, method=method, options=options, names=names)

    class SDEIto(torch.nn.Module):          # torchsde.SDEIto: an nn.Module carrying sde_type / noise_type
        def __init__(self, noise_type):
            super().__init__()
            self.sde_type, self.noise_type = "ito", noise_type

    tsde.SDEIto = SDEIto

    tsde.sdeint_adjoint = sdeint

    tsde.sdeint = sdeint
    tdiff = types.ModuleType("torchdiffeq")
    tdiff.odeint = tdiff.odeint_adjoint = None
    sys.modules.update(torchcde=tcde, torchsde=tsde, torchdiffeq=tdiff)


def load_reference_module(root, real_cde=True):
    for name in list(sys.modules):
        if name.split(".")[0] in ("models_sde", "controldiffeq"):
            del sys.modules[name]
    importlib.invalidate_caches()
    sys.path.insert(0, str(root))
    try:
        # load the file directly: models_sde/__init__.py drags in the CDE/GRU baselines
        if real_cde:
            import controldiffeq  # noqa: F401  (real in-tree package, torchdiffeq stubbed)
        else:
            # benchmark_forecasting/controldiffeq drags in TorchDiffEqPack -> matplotlib (absent);
            # neuralsde.py imports controldiffeq (:21) but never uses it.
            sys.modules["controldiffeq"] = types.ModuleType("controldiffeq")
        spec = importlib.util.spec_from_file_location(
            f"ref_neuralsde_{root.name}", root / "models_sde" / "neuralsde.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod, sys.modules["controldiffeq"]
    finally:
        sys.path.pop(0)


def load_torch_ists_module():
    """torch-ists' copy of the wrapper + vector field, loaded by file path the way the reference's own test does
    (/root/reference/tests/test_neuralsde_core_alignment.py:48-53): the torch_ists package itself is not importable."""
    path = REF / "torch-ists" / "torch_ists" / "diff_module" / "NSDE" / "nsde_model.py"
    spec = importlib.util.spec_from_file_location("ref_torch_ists_nsde_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_latent_module():
    """The reference's LatentSDE, loaded by file path (the torch_ists package itself is not importable)."""
    path = REF / "torch-ists" / "torch_ists" / "diff_module" / "NSDE" / "latent_sde.py"
    spec = importlib.util.spec_from_file_location("ref_torch_ists_latent_sde", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def latent_golden(ref_l):
    """LatentSDE (latent_sde.py:29-147): f_aug / g_aug at fixed (t, y) and the full forward on the oracle solver."""
    out = []
    with torch.no_grad():
        for i, (C, H, HH, L, theta, mu, sigma) in enumerate(((3, 5, 6, 1, 1.0, 0.0, 0.5), (2, 9, 8, 3, 0.7, 0.2, 0.3),
                                                              (4, 33, 40, 2, 1.3, -0.1, 0.8), (2, 4, 4, 1, 1.0, 0.1, 5e-8))):
            torch.manual_seed(600 + i)
            m = ref_l.LatentSDE(C, H, HH, L, theta=theta, mu=mu, sigma=sigma)
            B = 4
            y = torch.randn(B, H) * 0.8
            tq = [torch.tensor(0.0), torch.tensor(0.37), torch.tensor(2.5)]
            out.append(dict(kind="fg", dims=(C, H, HH, L), prior=(theta, mu, sigma),
                            state_dict={k: v.clone() for k, v in m.state_dict().items()}, y=y, t=torch.stack(tq),
                            f_aug=torch.stack([m.f_aug(t, y) for t in tq]),
                            g_aug=torch.stack([m.g_aug(t, y) for t in tq])))
        for i, (method, grid) in enumerate(((None, "linspace"), ("euler", "linspace"), ("euler", "arange"), ("srk", "arange"),
                                            ("milstein", "linspace"))):
            B, K, C, H, HH, L = 5, 9, 3, 8, 12, 2
            torch.manual_seed(700 + i)
            m = ref_l.LatentSDE(C, H, HH, L, theta=0.9, mu=0.1, sigma=0.4)
            m.qy0_mean.data.fill_(0.3); m.qy0_logvar.data.fill_(-1.2)
            times = torch.linspace(0, 1, K) if grid == "linspace" else torch.arange(K, dtype=torch.float32) * 0.25
            x = torch.randn(B, K, C).cumsum(1) * 0.3
            coeffs = ospline.hermite_cubic_coefficients_with_backward_differences(x, times)
            steps = osolver.step_times(times, osolver.solver_dt(times))
            h = torch.tensor([b - a for a, b in steps]).view(-1, 1, 1)
            dW = torch.randn(len(steps), B, H) * h.sqrt()
            dU = h * (dW / 2 + torch.randn(len(steps), B, H) * (h / 12).sqrt())
            kw = {} if method is None else {"method": method}
            pred, latent, logqp = m(coeffs, times, bm=osolver.BrownianTable(dW, dU=dU), **kw)
            out.append(dict(kind="forward", method=method, dims=(B, K, C, H, HH, L), prior=(0.9, 0.1, 0.4), n_steps=len(steps),
                            state_dict={k: v.clone() for k, v in m.state_dict().items()},
                            times=times, coeffs=coeffs, dW=dW, dU=dU, pred=pred, latent=latent, logqp=logqp))
    torch.save(out, HERE / "latent_golden.pt")
    print("latent_golden.pt:", len(out), "cases")


def make_inputs(seed, B, K, C, H):
    g = torch.Generator().manual_seed(seed)
    times = torch.linspace(0.0, 2.0, K)
    x = torch.randn(B, K, C, generator=g).cumsum(1) * 0.5
    coeffs = ospline.hermite_cubic_coefficients_with_backward_differences(x, times)
    y = torch.randn(B, H, generator=g) * 1.5
    return times, coeffs, y


def fg_golden(ref):
    cases = []
    with torch.no_grad():
        for io in range(7):
            for no in range(20):
                B, K, C, H, L = 3, 6, 3, 4, 2
                HH = 6 if io in (1, 3, 5) else H      # HH != H is only legal without emb / opt 0
                torch.manual_seed(1000 + 20 * io + no)
                m = ref.Diffusion_model(C, H, HH, L, theta=0.7, sigma=-0.3,
                                        input_option=io, noise_option=no)
                times, coeffs, y = make_inputs(77 + io, B, K, C, H)
                m.set_X(coeffs, times)
                tq = [times[0], times[2], (times[2] + times[3]) / 2, times[-1]]
                cases.append(dict(
                    input_option=io, noise_option=no, dims=(B, K, C, H, HH, L),
                    state_dict={k: v.clone() for k, v in m.state_dict().items()},
                    times=times, coeffs=coeffs, y=y, t=torch.stack(tq),
                    f=torch.stack([m.f(t, y) for t in tq]),
                    g=torch.stack([m.g(t, y) for t in tq])))
        # the reference test's own fixture recipe (tests/test_neuralsde_core_alignment.py:56-65,108-114)
        for name, (io, no) in {"lsde": (2, 16), "lnsde": (4, 17), "gsde": (6, 17)}.items():
            B, K, C, H, L = 2, 5, 3, 4, 2
            torch.manual_seed(4242 + io)
            m = ref.Diffusion_model(input_channels=C, hidden_channels=H, hidden_hidden_channels=H,
                                    num_hidden_layers=L, input_option=io, noise_option=no)
            times = torch.linspace(0.0, 1.0, steps=K)
            values = torch.linspace(0.1, 1.0, steps=B * K * C, dtype=torch.float32).reshape(B, K, C)
            coeffs = ospline.hermite_cubic_coefficients_with_backward_differences(values, t=times)
            y = torch.linspace(0.2, 0.9, steps=B * H, dtype=torch.float32).reshape(B, H)
            m.set_X(coeffs, times)
            t = times[2]
            cases.append(dict(
                input_option=io, noise_option=no, dims=(B, K, C, H, H, L), name=name,
                state_dict={k: v.clone() for k, v in m.state_dict().items()},
                times=times, coeffs=coeffs, y=y, t=t.unsqueeze(0),
                f=m.f(t, y).unsqueeze(0), g=m.g(t, y).unsqueeze(0)))
    torch.save(cases, HERE / "fg_golden.pt")
    print("fg_golden.pt:", len(cases), "cases")


def spline_golden(cde):
    g = torch.Generator().manual_seed(5)
    out = []
    for K, C, B in ((2, 2, 2), (3, 2, 2), (9, 3, 4), (50, 5, 2)):
        times = torch.cat([torch.zeros(1), torch.rand(K - 1, generator=g) + 0.2]).cumsum(0)
        x = torch.randn(B, K, C, generator=g)
        a, b, c2, d3 = cde.natural_cubic_spline_coeffs(times, x)
        sp = cde.NaturalCubicSpline(times, (a, b, c2, d3))
        tq = torch.cat([times, (times[1:] + times[:-1]) / 2,
                        torch.tensor([times[0] - 0.3, times[-1] + 0.3])])
        ev = torch.stack([sp.evaluate(t) for t in tq])
        out.append(dict(times=times, x=x, a=a, b=b, two_c=c2, three_d=d3, tq=tq, evaluate=ev))
    # missing values (interpolate.py:56-153): NaNs inside, at either end, a whole series missing
    for K, C, B in ((7, 3, 3), (12, 2, 2)):
        times = torch.cat([torch.zeros(1), torch.rand(K - 1, generator=g) + 0.2]).cumsum(0)
        x = torch.randn(B, K, C, generator=g)
        x[torch.rand(B, K, C, generator=g) < 0.35] = float("nan")
        x[0, 0, 0] = float("nan"); x[0, -1, 1] = float("nan"); x[1, :, 0] = float("nan"); x[1, 2, 1] = 0.5
        a, b, c2, d3 = cde.natural_cubic_spline_coeffs(times, x)
        out.append(dict(times=times, x=x, a=a, b=b, two_c=c2, three_d=d3, missing=True))
    torch.save(out, HERE / "spline_golden.pt")
    print("spline_golden.pt:", len(out), "cases")


def forward_golden(ref_c, ref_f, ref_t=None):
    out = []
    with torch.no_grad():
        # classification wrapper: final_index gather
        for io, no, fi in ((4, 17, [3, 5, 7, 7, 2]), (2, 16, [0, 7, 4, 1, 7]), (1, 18, [7, 7, 7, 7, 7])):
            B, K, C, H, L = 5, 8, 3, 4, 1
            torch.manual_seed(99 + io)
            func = ref_c.Diffusion_model(C, H, H, L, input_option=io, noise_option=no)
            model = ref_c.NeuralSDE(func, C, H, 2, initial=False)
            model.linear = torch.nn.Identity()
            times, coeffs, z0 = make_inputs(11 + io, B, K, C, H)
            times = torch.arange(K, dtype=torch.float32)
            final_index = torch.tensor(fi)
            S = K - 1
            dW = torch.randn(S, B, H)
            z = model(times, [coeffs], final_index, z0=z0, bm=osolver.BrownianTable(dW))
            out.append(dict(kind="classification", input_option=io, noise_option=no, dims=(B, K, C, H, H, L),
                            state_dict={k: v.clone() for k, v in func.state_dict().items()},
                            times=times, coeffs=coeffs, final_index=final_index, z0=z0, dW=dW, z=z))
        # forecasting wrapper: natural-spline 4-tuple coeffs, stream, tail slice
        B, K, C, H, L, output_time = 3, 7, 2, 4, 2, 3
        torch.manual_seed(123)
        func = ref_f.Diffusion_model(C, H, H, L, input_option=4, noise_option=17)
        model = ref_f.NeuralSDE_forecasting(func, C, output_time, H, 2, initial=True)
        model.linear = torch.nn.Identity()
        times = torch.linspace(0, K - 1, K)
        x = torch.randn(B, K, C).cumsum(1) * 0.3
        coeffs4 = ospline.natural_cubic_spline_coeffs(times, x)
        dW = torch.randn(K - 1, B, H)
        z = model(times, coeffs4, None, bm=osolver.BrownianTable(dW))
        out.append(dict(kind="forecasting", input_option=4, noise_option=17, dims=(B, K, C, H, H, L),
                        output_time=output_time,
                        state_dict={k: v.clone() for k, v in func.state_dict().items()},
                        initial_network={k: v.clone() for k, v in model.initial_network.state_dict().items()},
                        times=times, coeffs=torch.cat(coeffs4, -1), dW=dW, z=z))
        # the wrappers WITH their own neighbours in eval mode: z0 = initial_network(X(t0)) and the real
        # read-out heads (classification: Linear, BatchNorm1d (non-trivial running statistics), ReLU, Dropout, Linear;
        # forecasting: Linear, ReLU, Linear)
        B, K, C, H, L, O = 6, 9, 3, 8, 1, 2
        torch.manual_seed(777)
        func = ref_c.Diffusion_model(C, H, H, L, input_option=4, noise_option=17)
        model = ref_c.NeuralSDE(func, C, H, O, initial=True)
        bn = model.linear[1]
        bn.running_mean.copy_(torch.randn(H) * 0.3); bn.running_var.copy_(torch.rand(H) + 0.5)
        bn.weight.data.copy_(torch.rand(H) + 0.5); bn.bias.data.copy_(torch.randn(H) * 0.2)
        model.eval()
        times = torch.arange(K, dtype=torch.float32)
        x = torch.randn(B, K, C).cumsum(1) * 0.3
        coeffs = ospline.hermite_cubic_coefficients_with_backward_differences(x, times)
        final_index = torch.tensor([8, 3, 5, 8, 1, 6])
        dW = torch.randn(K - 1, B, H)
        pred = model(times, [coeffs], final_index, bm=osolver.BrownianTable(dW))
        out.append(dict(kind="classification_full", input_option=4, noise_option=17, dims=(B, K, C, H, H, L), out_channels=O,
                        state_dict={k: v.clone() for k, v in func.state_dict().items()},
                        model_state={k: v.clone() for k, v in model.state_dict().items()},
                        times=times, coeffs=coeffs, final_index=final_index, dW=dW, pred=pred))
        B, K, C, H, L, O, output_time = 4, 8, 2, 8, 1, 3, 3
        torch.manual_seed(778)
        func = ref_f.Diffusion_model(C, H, H, L, input_option=6, noise_option=17)
        model = ref_f.NeuralSDE_forecasting(func, C, output_time, H, O, initial=True)
        model.eval()
        times = torch.linspace(0, K - 1, K)
        x = torch.randn(B, K, C).cumsum(1) * 0.3
        coeffs4 = ospline.natural_cubic_spline_coeffs(times, x)
        dW = torch.randn(K - 1, B, H)
        pred = model(times, coeffs4, None, bm=osolver.BrownianTable(dW))
        out.append(dict(kind="forecasting_full", input_option=6, noise_option=17, dims=(B, K, C, H, H, L), out_channels=O,
                        output_time=output_time,
                        state_dict={k: v.clone() for k, v in func.state_dict().items()},
                        model_state={k: v.clone() for k, v in model.state_dict().items()},
                        times=times, coeffs=torch.cat(coeffs4, -1), dW=dW, pred=pred))
        # torch-ists wrapper (nsde_model.py:45-84): forward(coeffs, times) -> (head(z), z), every knot, default method
        # 'srk' (needs the space-time Levy integrals beside the increments), linspace grid with a sliver step
        if ref_t is not None:
            for method, io, no in ((None, 4, 17), ("euler", 6, 17), ("srk", 2, 16)):
                B, K, C, H, L = 4, 9, 3, 8, 2
                torch.manual_seed(321 + io)
                func = ref_t.Diffusion_model(C, H, H, L, input_option=io, noise_option=no)
                model = ref_t.NeuralSDE(func, C, H, 2, initial=True)
                times = torch.linspace(0, 1, K)
                x = torch.randn(B, K, C).cumsum(1) * 0.3
                coeffs = ospline.hermite_cubic_coefficients_with_backward_differences(x, times)
                S = len(osolver.step_times(times, osolver.solver_dt(times)))
                dW = torch.randn(S, B, H) * (1.0 / (K - 1)) ** 0.5
                dU = torch.randn(S, B, H) * (1.0 / (K - 1)) ** 1.5 / 3 ** 0.5
                kw = {} if method is None else {"method": method}
                pred, z = model(coeffs, times, bm=osolver.BrownianTable(dW, dU=dU), **kw)
                out.append(dict(kind="torch_ists", method=method, input_option=io, noise_option=no,
                                dims=(B, K, C, H, H, L), n_steps=S,
                                state_dict={k: v.clone() for k, v in func.state_dict().items()},
                                model_state={k: v.clone() for k, v in model.state_dict().items()},
                                times=times, coeffs=coeffs, dW=dW, dU=dU, z=z, pred=pred))
    torch.save(out, HERE / "forward_golden.pt")
    print("forward_golden.pt:", len(out), "cases")


if __name__ == "__main__":
    install_shims()
    if "--only-latent" in sys.argv:         # the other fixtures are left untouched
        latent_golden(load_latent_module())
        sys.exit(0)
    ref_c, cde = load_reference_module(REF / "benchmark_classification")
    fg_golden(ref_c)
    spline_golden(cde)
    ref_f, _ = load_reference_module(REF / "benchmark_forecasting", real_cde=False)
    ref_c, _ = load_reference_module(REF / "benchmark_classification")
    forward_golden(ref_c, ref_f, load_torch_ists_module())
    latent_golden(load_latent_module())

"""CPU-only checks of the host side: C-ABI surface, step plan, packing, sharding (gloo)."""
import ctypes
import os
import pathlib
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import snsde_b200
from oracle import solver, vector_field
from snsde_b200 import _lib, packing

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "snsde.h").read_text()
    declared = set(re.findall(r"\b(snsde_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.snsde_abi_version() == _lib.ABI_VERSION == 4
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", nm), name


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Step) == 40 == snsde_b200.stepplan.STEP_DTYPE.itemsize
    assert ctypes.sizeof(_lib.Emit) == 12 == snsde_b200.stepplan.EMIT_DTYPE.itemsize
    assert ctypes.sizeof(_lib.ModelDesc) == 36
    assert ctypes.sizeof(_lib.Point) == 20 == snsde_b200.stepplan.POINT_DTYPE.itemsize
    for (n, _), (m, _t) in zip(_lib.Point._fields_, snsde_b200.stepplan.POINT_DTYPE.descr):
        assert n == m
    for (n, _), (m, _t) in zip(_lib.Step._fields_, snsde_b200.stepplan.STEP_DTYPE.descr):
        assert n == m


@pytest.mark.parametrize("io", range(7))
@pytest.mark.parametrize("no", range(20))
def test_weight_count_matches_state_dict(io, no):
    H, C, L = 8, 3, 3
    HH = 12 if io in (1, 3, 5) else H
    m = vector_field.DiffusionModel(C, H, HH, L, input_option=io, noise_option=no)
    desc = packing.describe(m)
    assert (desc["input_option"], desc["noise_option"], desc["hidden_hidden"], desc["num_hidden_layers"]) == (io, no, HH, L)
    blob = packing.pack(m, desc)
    assert blob.numel() == sum(v.numel() for v in m.state_dict().values())
    cd = _lib.ModelDesc(method=0, precision=0, **desc)
    assert _lib.load().snsde_weight_count(ctypes.byref(cd)) == blob.numel()


def test_weight_count_tutorial_and_validation_errors():
    lib = _lib.load()
    m = vector_field.TutorialLSDEFunc(2, 32, 16, 2)
    desc = packing.describe(m)
    assert desc["family"] == _lib.FAMILY_TUTORIAL_LSDE and desc["hidden_hidden"] == 16 and desc["num_hidden_layers"] == 2
    cd = _lib.ModelDesc(method=0, precision=0, **desc)
    assert lib.snsde_weight_count(ctypes.byref(cd)) == packing.pack(m, desc).numel()
    bad = _lib.ModelDesc(family=0, input_option=4, noise_option=20, input_channels=3, hidden=4, hidden_hidden=4,
                         num_hidden_layers=1, method=0, precision=0)
    assert lib.snsde_weight_count(ctypes.byref(bad)) == _lib.ERR_BAD_ARG
    assert b"Unknown noise_option 20" in lib.snsde_last_error()
    bad.noise_option, bad.hidden_hidden = 17, 8          # emb needs HH == H
    assert lib.snsde_weight_count(ctypes.byref(bad)) == _lib.ERR_BAD_ARG
    bad.hidden_hidden, bad.method = 4, 3                 # unknown method id
    assert lib.snsde_weight_count(ctypes.byref(bad)) == _lib.ERR_UNSUPPORTED
    bad.method = 2                                       # srk (torch-ists default) is implemented
    assert lib.snsde_weight_count(ctypes.byref(bad)) > 0
    bad.method, bad.noise_option = 1, 18                 # milstein through the noise network (full vjp) is supported
    assert lib.snsde_weight_count(ctypes.byref(bad)) > 0
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_BAD_ARG)


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = vector_field.DiffusionModel(3, 4, 4, 1, input_option=4, noise_option=17)
    m.set_X(torch.zeros(2, 4, 12), torch.arange(5.0))
    with torch.no_grad(), pytest.raises(snsde_b200.EngineError):
        snsde_b200.sdeint(m, torch.zeros(2, 4), torch.arange(5.0), dt=1.0)


@pytest.mark.parametrize("ts,dt", [
    (torch.linspace(0, 1, 20), 0.05),                    # tutorial grid: steps do not land on knots
    (torch.linspace(0, 1, 64), None),                    # float32 sliver step
    (torch.arange(73.0) + 1, None),                      # Sepsis: linspace(1,72,72)-like integer grid
    (torch.tensor([0.0, 0.13, 0.25, 0.5]), 0.1),
    (torch.tensor([0.0, 0.05, 0.06, 0.07, 1.0]), 0.5),   # several outputs inside one step
])
def test_step_plan_replays_the_oracle_loop(ts, dt):
    dt = solver.solver_dt(ts) if dt is None else dt
    sp = snsde_b200.build_step_plan(ts.numpy(), dt, ts.numpy())
    ref = solver.step_times(ts, dt)
    assert sp.n_steps == len(ref) and sp.n_out == len(ts)
    for (a, b), s in zip(ref, sp.steps):
        assert np.float32(a) == s["t0"] and np.float32(np.float32(b) - np.float32(a)) == s["h"]
        assert s["sqrt_h"] == np.sqrt(s["h"])
        idx = int(torch.bucketize(torch.tensor(a), ts).sub(1).clamp(0, len(ts) - 2))
        assert idx == s["interval"] and s["frac"] == np.float32(np.float32(a) - ts[idx].numpy())
    # emits: contiguous ranges, one per output, lerp weights reproduce the oracle's interpolation
    assert sp.n_init_emits == 1 and sp.emits[0]["slot"] == 0
    assert list(sp.emits["slot"]) == list(range(len(ts)))
    prev_end = 1
    for s in sp.steps:
        assert s["emit_begin"] == prev_end and s["emit_end"] >= s["emit_begin"]
        prev_end = s["emit_end"]
    assert prev_end == len(sp.emits)
    # replay a scalar "identity" SDE through the plan and compare with the oracle on y(t) = t path
    class Lin(torch.nn.Module):
        sde_type, noise_type = "ito", "diagonal"
        def f(self, t, y): return torch.ones_like(y) * 2.0
        def g(self, t, y): return torch.zeros_like(y)
    y0 = torch.tensor([[0.5]])
    out = solver.sdeint(Lin(), y0, ts, dt, solver.BrownianTable(torch.zeros(sp.n_steps, 1, 1)))
    y, got, e = np.float32(0.5), {0: np.float32(0.5)}, 1
    for s in sp.steps:
        yn = np.float32(y + np.float32(2.0) * s["h"])
        for em in sp.emits[s["emit_begin"]:s["emit_end"]]:
            got[int(em["slot"])] = np.float32(em["w_prev"] * y + em["w_curr"] * yn)
        y = yn
    assert np.allclose([got[i] for i in range(len(ts))], out[:, 0, 0].numpy(), rtol=1e-6, atol=1e-7)


def test_step_plan_rejects_bad_input():
    with pytest.raises(ValueError):
        snsde_b200.build_step_plan(np.array([0.0, 0.0, 1.0]), 0.1)
    with pytest.raises(ValueError):
        snsde_b200.build_step_plan(np.array([0.0, 1.0]), 0.0)
    sp = snsde_b200.build_step_plan(np.array([3.0]), 0.1)          # a single output time: no steps
    assert sp.n_steps == 0 and len(sp.emits) == 1


def test_final_index_slots_match_reference_bookkeeping(golden_dir):
    from oracle import wrapper
    times = torch.arange(8.0)
    for fi in ([3, 5, 7, 7, 2], [0, 7, 4, 1, 7], [7] * 5, [0] * 3, [1, 1, 6]):
        fi = torch.tensor(fi)
        ts_ref, g_ref = wrapper.output_times_for_final_index(times, fi)
        ts, slots = snsde_b200.final_index_slots(times, fi)
        assert torch.equal(ts, ts_ref) and torch.equal(slots, g_ref)
        assert torch.equal(ts[slots], times[fi])


def test_shard_bounds_cover_rows_exactly():
    from snsde_b200.dist import shard_bounds
    for n in (1, 7, 1024, 1025, 8191):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import snsde_b200
from snsde_b200.dist import solve_final_sharded, shard_bounds
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
for n in (8, 7):
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 1.5
    def solve(lo, hi, out):          # stands in for engine.solve_final(..., row_offset=lo, out=out)
        assert (lo, hi) == shard_bounds(n, rank, world)
        out.copy_(full[lo:hi])
    got = solve_final_sharded(solve, n, 3, "cpu")
    assert torch.equal(got, full), (rank, n)
dist.barrier(); dist.destroy_process_group(); print("ok", rank)
"""


def test_two_rank_gloo_all_gather_of_latents(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


# ---- drop-in boundary against the REAL reference classes (build container only: needs /root/reference) ----
REF = pathlib.Path("/root/reference")


def _golden_tools():
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    try:
        import make_golden
    finally:
        sys.path.pop(0)
    return make_golden


def _drop_shims():
    for name in ("torchcde", "torchsde", "torchdiffeq", "controldiffeq"):
        sys.modules.pop(name, None)


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("bench_dir,io,no", [("benchmark_classification", 4, 17), ("benchmark_classification", 1, 18),
                                              ("benchmark_forecasting", 6, 17), ("benchmark_classification", 2, 5)])
def test_packing_accepts_the_reference_own_modules(bench_dir, io, no):
    mg = _golden_tools()
    mg.install_shims()
    try:
        ref, _ = mg.load_reference_module(REF / bench_dir, real_cde=(bench_dir == "benchmark_classification"))
        torch.manual_seed(0)
        func = ref.Diffusion_model(input_channels=5, hidden_channels=32, hidden_hidden_channels=32, num_hidden_layers=2,
                                   input_option=io, noise_option=no)
        desc = packing.describe(func)
        assert (desc["input_option"], desc["noise_option"], desc["hidden"], desc["num_hidden_layers"]) == (io, no, 32, 2)
        blob = packing.pack(func, desc)
        own = vector_field.DiffusionModel(5, 32, 32, 2, input_option=io, noise_option=no)
        own.load_state_dict(func.state_dict())
        assert torch.equal(blob, packing.pack(own, packing.describe(own)))        # same bytes from either module
        cd = _lib.ModelDesc(method=0, precision=0, **desc)
        assert _lib.load().snsde_weight_count(ctypes.byref(cd)) == blob.numel()
        assert all(k in dict(func.named_parameters()) for k in packing.blob_keys(desc))   # the backward returns one grad per key
    finally:
        _drop_shims()


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted (GPU box)")
def test_patch_dispatches_on_each_of_the_three_reference_wrappers():
    """patch() must be correct for the classification NeuralSDE, NeuralSDE_forecasting and the torch-ists NeuralSDE
    (the forecasting / torch-ists forwards were replaced by the classification one)."""
    import copy
    import io as _io
    mg = _golden_tools()
    mg.install_shims()
    try:
        ref_c, _ = mg.load_reference_module(REF / "benchmark_classification")
        ref_f, _ = mg.load_reference_module(REF / "benchmark_forecasting", real_cde=False)
        ref_t = mg.load_torch_ists_module()
        mk = lambda ref: ref.Diffusion_model(5, 32, 32, 1, input_option=4, noise_option=17)       # noqa: E731
        cls = snsde_b200.patch(ref_c.NeuralSDE(mk(ref_c), 5, 32, 3, initial=False))
        fore = snsde_b200.patch(ref_f.NeuralSDE_forecasting(mk(ref_f), 5, 4, 32, 3, initial=True))
        ists = snsde_b200.patch(ref_t.NeuralSDE(mk(ref_t), 5, 32, 3, initial=True))
        assert [snsde_b200.wrapper_kind(m) for m in (cls, fore, ists)] == ["classification", "forecasting", "torch_ists"]
        from snsde_b200 import engine
        assert cls.forward.__func__ is engine._forward_classification
        assert fore.forward.__func__ is engine._forward_forecasting
        assert "forward" not in ists.__dict__                                   # torch-ists forward(coeffs, times) untouched
        assert cls._solve_sde_path.__func__ is engine._solve_sde_path_benchmark
        assert fore._solve_sde_path.__func__ is engine._solve_sde_path_benchmark
        assert ists._solve_sde_path.__func__ is engine._solve_sde_path_torch_ists
        # a deep copy drives ITS OWN func; plans (ctypes handles) are not part of the module
        dup = copy.deepcopy(cls)
        assert dup._solve_sde_path.__self__ is dup and dup.forward.__self__ is dup and dup.func is not cls.func
        assert not any("snsde" in k for k in cls.func.__dict__)
        torch.save(cls.state_dict(), _io.BytesIO())
        times = torch.arange(6.0)
        coeffs = torch.zeros(2, 5, 20)
        if not torch.cuda.is_available():       # every wrapper reaches the engine, which refuses without a GPU
            with torch.no_grad():
                with pytest.raises(snsde_b200.EngineError):
                    cls(times, [coeffs], torch.tensor([5, 3]), z0=torch.zeros(2, 32))
                with pytest.raises(snsde_b200.EngineError):
                    dup(times, [coeffs], torch.tensor([5, 3]), z0=torch.zeros(2, 32))
                with pytest.raises(snsde_b200.EngineError):
                    fore(times, (coeffs[..., :5], coeffs[..., 5:10], coeffs[..., 10:15], coeffs[..., 15:]), None)
                with pytest.raises(snsde_b200.EngineError):
                    ists(coeffs, times)                                        # default method 'srk'
    finally:
        _drop_shims()


def test_torch_ists_style_wrapper_defaults_to_srk():
    class IstsStyle(torch.nn.Module):                      # signature of torch-ists NeuralSDE._solve_sde_path (nsde_model.py:63)
        def __init__(self):
            super().__init__()
            self.func = vector_field.DiffusionModel(3, 8, 8, 1, input_option=4, noise_option=17)

        def _solve_sde_path(self, times, y0, kwargs):
            raise AssertionError("replaced by patch()")

        def forward(self, coeffs, times, **kwargs):
            return self._solve_sde_path(times, torch.zeros(coeffs.shape[0], 8), kwargs)
    m = snsde_b200.patch(IstsStyle())
    assert snsde_b200.wrapper_kind(m) == "torch_ists" and "forward" not in m.__dict__
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(snsde_b200.EngineError):
            m(torch.zeros(2, 3, 12), torch.arange(4.0))
    with pytest.raises(ValueError, match="not implemented"):
        with torch.no_grad():
            m(torch.zeros(2, 3, 12), torch.arange(4.0), method="heun")


def test_step_plan_dense_states_and_srk_points():
    sp = snsde_b200.build_step_plan(np.array([0.0, 0.25, 1.0], dtype=np.float32), 0.4, np.linspace(0, 1, 5, dtype=np.float32),
                                    method="srk")
    S = sp.n_steps
    assert S == 3 and sp.points.shape == (S, 4)
    h, t0 = sp.steps["h"], sp.steps["t0"]
    for i, c in enumerate((0.0, 0.25, 0.5, 1.0)):
        t = (t0 + np.float32(c) * h).astype(np.float32)
        assert np.array_equal(sp.points["t"][:, i], t)
        assert np.allclose(sp.points["sin_t"][:, i], np.sin(t))
        kn = np.linspace(0, 1, 5, dtype=np.float32)
        idx = np.clip(np.searchsorted(kn, t, side="left") - 1, 0, 3)
        assert np.array_equal(sp.points["interval"][:, i], idx)
    d = sp.dense()
    assert d.n_out == S + 1 and len(d.emits) == S + 1 and np.array_equal(d.emits["slot"], np.arange(S + 1))
    assert np.array_equal(d.steps["emit_begin"], np.arange(1, S + 1)) and np.all(d.emits["w_curr"] == 1)
    slot, k, wp, wc = sp.output_map()
    # ts[1] = 0.25 falls inside step 0 (0 -> 0.4): lerp of Y[0], Y[1]; ts[2] = 1.0 is the end of the last step
    assert list(slot) == [0, 1, 2] and list(k) == [-1, 0, S - 1]
    assert abs(wp[1] - 0.375) < 1e-6 and abs(wc[1] - 0.625) < 1e-6 and wp[2] == 0 and wc[2] == 1


def test_bench_algorithmic_costs_match_the_survey_contract():
    """bench.py's roofline numerators are SURVEY 8(d)'s per-SDE-step figures (layer-wise form, L = 1)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    flops = {"c1": 10368, "c2": 140544, "c3": 37504, "c4": 132096, "c5": 532480}
    for k, f in flops.items():
        assert bench.alg_flops_per_sde_step(bench.WORKLOADS[k]) == f
    # bytes: one 16C-byte spline row per step (models that read the control) + outputs + z0, amortised over S
    assert abs(bench.alg_bytes_per_sde_step(bench.WORKLOADS["c2"], 1) - (560 + 4 * 128 * 2 / 200)) < 1e-9
    assert bench.alg_bytes_per_sde_step(bench.WORKLOADS["c4"], 2) < 16          # input option 3 never reads X(t)
    assert abs(bench.alg_bytes_per_sde_step(bench.WORKLOADS["c5"], 10) - (224 + 4 * 256 * 11 / 500)) < 1e-9


def test_control_at_equals_the_spline_read_and_carries_gradients():
    """engine._control_at (the torch-op spline read the training path of the patched LatentSDE.forward uses,
    latent_sde.py:100-101) equals CubicSpline.evaluate at knots, inside intervals and when clamped outside the grid."""
    from snsde_b200 import engine
    from oracle import spline
    torch.manual_seed(2)
    B, K, C = 3, 6, 4
    times = torch.cat([torch.zeros(1), torch.rand(K - 1) + 0.2]).cumsum(0)
    x = torch.randn(B, K, C)
    coeffs = spline.hermite_cubic_coefficients_with_backward_differences(x, times).requires_grad_(True)
    X = spline.CubicSpline(coeffs.detach(), times)
    for t in (times[0], times[2], (times[2] + times[3]) / 2, times[-1], times[0] - 0.3, times[-1] + 0.4):
        got = engine._control_at(coeffs, times, t)
        assert torch.allclose(got, X.evaluate(t), rtol=1e-6, atol=1e-6)
    engine._control_at(coeffs, times, times[0]).sum().backward()
    assert coeffs.grad is not None and float(coeffs.grad[:, 0, :C].abs().sum()) == B * C       # X(t0) = a of interval 0


def test_vectorised_step_plan_is_bit_identical_to_the_loop():
    """stepplan.build_step_plan (sequential float32 accumulation + searchsorted) against the statement-for-statement
    loop, over integer / linspace (sliver last step) / irregular grids, dt below and above the knot spacing, output
    times on and between knots, both methods."""
    from snsde_b200 import stepplan as sp
    rng = np.random.default_rng(0)
    n = 0
    for trial in range(600):
        kind = trial % 5
        K = int(rng.integers(2, 60))
        if kind == 0: knots = np.arange(K, dtype=np.float32)
        elif kind == 1: knots = np.linspace(0, 1, K).astype(np.float32)
        elif kind == 2: knots = np.cumsum(np.concatenate([[rng.normal()], rng.random(K - 1) + 0.05])).astype(np.float32)
        elif kind == 3: knots = (np.arange(K) * 0.25 + 3).astype(np.float32)
        else: knots = np.linspace(-2, 5, K).astype(np.float32)
        dmin = float((knots[1:] - knots[:-1]).min())
        dt = [max(dmin, 1e-3), dmin * 0.37, dmin * 2.3, 1e-3 if dmin < 0.2 else 0.05][trial % 4]
        ts = knots[np.sort(rng.choice(K, size=int(rng.integers(1, K + 1)), replace=False))]
        if trial % 7 == 0 and len(ts) > 2:
            ts = np.unique(np.concatenate([ts, (ts[:-1] + ts[1:]) / 2])).astype(np.float32)
        for method in ("euler", "srk"):
            a, b = sp.build_step_plan(ts, dt, knots, method=method), sp.build_step_plan_loop(ts, dt, knots, method=method)
            assert (a.n_out, a.n_init_emits, a.n_knots) == (b.n_out, b.n_init_emits, b.n_knots)
            assert a.steps.tobytes() == b.steps.tobytes() and a.emits.tobytes() == b.emits.tobytes()
            assert (a.points is None) == (b.points is None) and (a.points is None or a.points.tobytes() == b.points.tobytes())
            n += 1
    assert n == 1200
    for bad in (lambda: sp.build_step_plan(np.float32([16777216, 16777300]), 0.5),       # fl32(2^24 + 0.5) == 2^24: no advance
                lambda: sp.build_step_plan(np.float32([0, 1]), 0.0)):
        with pytest.raises(ValueError):
            bad()

"""GPU tests at BASELINE.json's full sizes: oracle comparison where the CPU finishes in seconds, plus
size-independent properties (two independent kernels agree, sharding is bit-invariant, the fused gather equals
gathering the streamed output, noise-free solves ignore the seed)."""
import pytest
import torch

import snsde_b200
from snsde_b200 import data, modules
from oracle import solver, vector_field, wrapper

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1.0)


def quantile_err(a, b, q):
    """q-quantile of |a-b| relative to the scale of b.  Used where the max norm is meaningless: the Neural GSDE
    (multiplicative noise x geometric drift) over 200 unit steps is ill-conditioned - the reference's own fp32
    path differs from its fp64 evaluation by up to 15% of the scale in rare elements that regrow from ~1e-8
    (median 2e-7; measured with the oracle) - so parity is asserted on quantiles."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    d = (a - b).abs().flatten()
    k = max(1, int(q * d.numel()))
    return float(d.kthvalue(k).values) / max(float(b.abs().max()), 1.0)


def conditioned_tol(o32, want32, run64, with_fp64=False):
    """Parity tolerance that respects the conditioning of the workload: 1e-4, or 3x the distance between the
    reference path evaluated in fp32 and in fp64 when that is larger (fp32 itself cannot do better).
    ``with_fp64`` also returns the fp64 trajectory: on these chaotic horizons two fp32-class evaluations each sit
    ~delta from the exact trajectory and up to ~2 delta from each other, so the max-norm check is made against the
    closer of the two references (the quantile checks stay against the fp32 oracle)."""
    import copy
    want64 = run64(copy.deepcopy(o32).double())
    tol = max(RTOL, 3.0 * rel_err(want32, want64))
    return (tol, want64) if with_fp64 else tol


def workload(io, no, B, H, C, S, L=1, seed=0, natural=False):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = modules.DiffusionModelParams(C, H, H, L, input_option=io, noise_option=no)
    times = torch.arange(S + 1, dtype=torch.float32)
    x = (torch.randn(B, S + 1, C, generator=g) * 0.1).cumsum(1)
    x[..., 0] = times
    coeffs = (data.natural_cubic_coeffs if natural else data.hermite_backward_difference_coeffs)(x, times)
    z0 = torch.randn(B, H, generator=g) * 0.1
    fi = torch.randint(2, S + 1, (B,), generator=g)
    return m, times, coeffs, z0, fi


def oracle_model(m, io, no, C, H, L=1):
    o = vector_field.DiffusionModel(C, H, H, L, input_option=io, noise_option=no)
    o.load_state_dict(m.state_dict())
    return o


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def test_c2_full_size(dev):
    """Neural LNSDE (4,17), B=1024, H=128, C=35, 200 Euler steps, per-row final_index."""
    io, no, B, H, C, S = 4, 17, 1024, 128, 35, 200
    m, times, coeffs, z0, fi = workload(io, no, B, H, C, S)
    mg = m.to(dev)
    cg, tg, zg, fg = coeffs.to(dev), times.to(dev), z0.to(dev), fi.to(dev)
    with torch.no_grad():
        mg.set_X(cg, tg)
        z_tc = snsde_b200.solve_final(mg, tg, fg, zg, seed=5, precision="tc")
        z_fma = snsde_b200.solve_final(mg, tg, fg, zg, seed=5, precision="fp32")
        assert snsde_b200.plans_of(mg)[("euler", "tc", str(dev))].kernel == "tcgen05"
        # (1) two independent kernels, identical Philox stream
        assert rel_err(z_tc, z_fma) <= RTOL
        # (2) sharding invariance, bit for bit (4 shards, as in BASELINE config c4's 4-GPU split)
        ts, slots = snsde_b200.final_index_slots(tg, fg)
        plan = snsde_b200.plans_of(mg)[("euler", "tc", str(dev))]
        sp = plan.step_plan(ts, 1.0, tg)
        parts = [plan.forward(zg[lo:lo + 256], sp, coeffs=cg[lo:lo + 256], row_slot=slots[lo:lo + 256], seed=5, row_offset=lo)
                 for lo in range(0, B, 256)]
        assert torch.equal(torch.cat(parts), z_tc)
        # (3) fused final_index capture == gather of the streamed [K, B, H] output
        full = snsde_b200.sdeint(mg, zg, tg, dt=1.0, seed=5, precision="tc")
        assert torch.equal(full[fg, torch.arange(B, device=dev)], z_tc)
        # (4) the oracle on the engine's own increments
        dW = snsde_b200.philox_increments(5, plan.step_plan(tg, 1.0, tg), B, H, dev).cpu()
    want = wrapper.classification_latent(oracle_model(m.cpu(), io, no, C, H), times, coeffs, fi, z0, solver.BrownianTable(dW))
    assert rel_err(z_tc, want) <= RTOL and rel_err(z_fma, want) <= RTOL


def test_c3_full_size_milstein(dev):
    """Neural GSDE (6,17), Milstein, B=2048, H=64."""
    io, no, B, H, C, S = 6, 17, 2048, 64, 35, 200
    m, times, coeffs, z0, fi = workload(io, no, B, H, C, S, seed=1)
    mg = m.to(dev)
    cg, tg, zg, fg = coeffs.to(dev), times.to(dev), z0.to(dev), fi.to(dev)
    with torch.no_grad():
        mg.set_X(cg, tg)
        z_tc = snsde_b200.solve_final(mg, tg, fg, zg, method="milstein", seed=9, precision="tc")
        z_fma = snsde_b200.solve_final(mg, tg, fg, zg, method="milstein", seed=9, precision="fp32")
        z_eul = snsde_b200.solve_final(mg, tg, fg, zg, method="euler", seed=9, precision="tc")
        plan = snsde_b200.plans_of(mg)[("milstein", "tc", str(dev))]
        dW = snsde_b200.philox_increments(9, plan.step_plan(tg, 1.0, tg), B, H, dev).cpu()
    # ill-conditioned workload (see quantile_err): two fp32-class kernels agree on all but a sliver of elements
    assert quantile_err(z_tc, z_fma, 0.5) <= 1e-6 and quantile_err(z_tc, z_fma, 0.99) <= RTOL
    assert rel_err(z_tc, z_eul) > 1e-3                     # the Milstein correction is really applied
    sl = slice(0, 256)                                     # oracle Milstein needs autograd: a 256-row slice
    want = wrapper.classification_latent(oracle_model(m.cpu(), io, no, C, H), times, coeffs[sl], fi[sl], z0[sl],
                                         solver.BrownianTable(dW[:, sl]), method="milstein")
    # rows of a shard see the same increments as in the full batch (global-row keyed stream)
    assert quantile_err(z_tc[sl], want, 0.5) <= 1e-6 and quantile_err(z_tc[sl], want, 0.99) <= RTOL
    assert quantile_err(z_fma[sl], want, 0.5) <= 1e-6 and quantile_err(z_fma[sl], want, 0.99) <= RTOL
    # hard bound on EVERY element: no further from the fp32 reference than a few times the reference's own fp32-vs-fp64
    # distance on this trajectory (tests/golden/conditioning.json: up to 0.2 of the scale for this model at 200 steps)
    import copy
    o64 = copy.deepcopy(oracle_model(m.cpu(), io, no, C, H)).double()
    want64 = wrapper.classification_latent(o64, times.double(), coeffs[sl].double(), fi[sl], z0[sl].double(),
                                           solver.BrownianTable(dW[:, sl].double()), method="milstein")
    bound = max(RTOL, 3.0 * rel_err(want, want64))
    assert rel_err(z_tc[sl], want) <= bound and rel_err(z_fma[sl], want) <= bound, (rel_err(z_tc[sl], want), rel_err(z_fma[sl], want), bound)
    # strict gate on the well-conditioned part of the horizon: the first 24 steps of the FULL batch, 1e-4
    with torch.no_grad():
        mg.set_X(cg[:, :24], tg[:25])
        z24 = snsde_b200.sdeint(mg, zg, tg[:25], dt=1.0, method="milstein", seed=9, precision="tc")
    o = oracle_model(m.cpu(), io, no, C, H)
    o.set_X(coeffs[sl, :24], times[:25])
    want24 = solver.sdeint(o, z0[sl], times[:25], 1.0, solver.BrownianTable(dW[:24, sl]), method="milstein")
    assert rel_err(z24[:, sl], want24) <= RTOL


def test_c4_shape_state_network_noise(dev):
    """Neural SDE (3,18), Speech shape, 1024 rows (one GPU's shard of B=4096), 160 steps, last knot."""
    io, no, B, H, C, S = 3, 18, 1024, 128, 21, 160
    m, times, coeffs, z0, _ = workload(io, no, B, H, C, S, seed=2)
    mg = m.to(dev)
    tg, zg = times.to(dev), z0.to(dev)
    ts = times[[0, -1]]
    with torch.no_grad():
        mg.set_X(coeffs.to(dev), tg)
        z = snsde_b200.sdeint(mg, zg, ts.to(dev), dt=1.0, seed=3, row_offset=2048)
        plan = next(iter(snsde_b200.plans_of(mg).values()))
        sp = plan.step_plan(ts, 1.0, tg)
        halves = [plan.forward(zg[lo:lo + 512], sp, coeffs=None, seed=3, row_offset=2048 + lo) for lo in (0, 512)]
        assert torch.equal(torch.cat(halves, dim=1), z)
        dW = snsde_b200.philox_increments(3, sp, B, H, dev, row_offset=2048).cpu()
    o = oracle_model(m.cpu(), io, no, C, H)
    o.set_X(coeffs, times)
    want = solver.sdeint(o, z0, ts, 1.0, solver.BrownianTable(dW))

    def run64(o64):
        o64.set_X(coeffs.double(), times.double())
        return solver.sdeint(o64, z0.double(), ts.double(), 1.0, solver.BrownianTable(dW.double()))
    # ONE reference for the max norm: the reference path evaluated in fp64 (the closest thing to its exact trajectory);
    # tolerance 1e-4 or 3x the distance at which the reference's own fp32 evaluation sits from it (~6e-5: |z| reaches ~190)
    tol, want64 = conditioned_tol(o, want, run64, with_fp64=True)
    err = rel_err(z, want64)
    assert err <= tol, (err, rel_err(z, want), tol)
    assert quantile_err(z, want, 0.999) <= RTOL


def test_c5_shape_slice_natural_spline_tail(dev):
    """Neural LNSDE (4,17), MuJoCo shape: H=256, C=14, 500 steps, natural-spline coeffs, last 10 knots; 128-row slice."""
    io, no, B, H, C, S = 4, 17, 128, 256, 14, 500
    m, times, coeffs, z0, _ = workload(io, no, B, H, C, S, seed=3, natural=True)
    mg = m.to(dev)
    ts = torch.cat([times[:1], times[-10:]])
    with torch.no_grad():
        mg.set_X(coeffs.to(dev), times.to(dev))
        z = snsde_b200.sdeint(mg, z0.to(dev), ts.to(dev), dt=1.0, seed=4)
        plan = next(iter(snsde_b200.plans_of(mg).values()))
        dW = snsde_b200.philox_increments(4, plan.step_plan(ts, 1.0, times), B, H, dev).cpu()
    o = oracle_model(m.cpu(), io, no, C, H)
    o.set_X(coeffs, times)
    want = solver.sdeint(o, z0, ts, 1.0, solver.BrownianTable(dW))
    assert z.shape == (11, B, H)

    def run64(o64):
        o64.set_X(coeffs.double(), times.double())
        return solver.sdeint(o64, z0.double(), ts.double(), 1.0, solver.BrownianTable(dW.double()))
    tol = conditioned_tol(o, want, run64)          # 500 steps, |z| reaches ~470: fp32 vs fp64 reference ~1e-4
    assert rel_err(z, want) <= tol, (rel_err(z, want), tol)
    assert quantile_err(z, want, 0.999) <= RTOL


def test_noise_free_solves_ignore_the_seed(dev):
    m, times, coeffs, z0, fi = workload(2, 0, 300, 64, 5, 40, seed=4)
    mg = m.to(dev)
    with torch.no_grad():
        mg.set_X(coeffs.to(dev), times.to(dev))
        a = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), z0.to(dev), seed=1)
        b = snsde_b200.solve_final(mg, times.to(dev), fi.to(dev), z0.to(dev), seed=2)
    assert torch.equal(a, b)

#!/usr/bin/env python
"""SDE-steps/sec benchmark of the Neural-SDE solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]

One "step" = one forward trajectory solve of one synthetic batch (B rows x S solver steps);
SDE-steps/sec = B_global * S * K / seconds.  Default workload: BASELINE config c2 - Neural
LNSDE (input_option 4, noise_option 17), Sepsis shape, B=1024 rows per GPU, hidden 128, C=35,
200 Euler steps, per-row final_index capture.  N GPUs shard the batch (1024 rows each, weak
scaling) and all-gather the final latents over NCCL.

Printed JSON (one line, rank 0): see the prompt contract; `value` = device-resident solve timed
with CUDA events, `e2e` = the public API with HOST buffers (H2D of the inputs and D2H of the
latents inside the timed region), `roofline` = the solve kernel against MEASURED_PEAKS.json,
`cpu_baseline` = the CPU oracle (port of the reference path) on this box's host cores.
`--impl reference` times that CPU port alone (torchsde/torchcde are not installable: no
network), with all host threads.
"""
import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: model (input_option, noise_option), rows per GPU, hidden, channels, solver steps, layers, method, output
    "c1": dict(family="tutorial", io=0, no=0, B=64, H=32, C=2, S=50, L=1, method="euler", out="stream",
               desc="tutorial Neural LSDE func, B=64 H=32 50 Euler steps (dt=0.02 on linspace(0,1,51))"),
    "c2": dict(family="benchmark", io=4, no=17, B=1024, H=128, C=35, S=200, L=1, method="euler", out="final_index",
               desc="Neural LNSDE (4,17) Sepsis-shape B=1024/GPU H=128 C=35 200 Euler steps, final_index capture"),
    "c3": dict(family="benchmark", io=6, no=17, B=2048, H=64, C=35, S=200, L=1, method="milstein", out="final_index",
               desc="Neural GSDE (6,17) Milstein B=2048/GPU H=64 C=35 200 steps, final_index capture"),
    "c4": dict(family="benchmark", io=3, no=18, B=1024, H=128, C=21, S=160, L=1, method="euler", out="last",
               desc="Neural SDE (3,18) Speech-shape B=1024/GPU H=128 C=21 160 Euler steps, last knot"),
    "c5": dict(family="benchmark", io=4, no=17, B=1024, H=256, C=14, S=500, L=1, method="euler", out="tail10",
               desc="Neural LNSDE (4,17) MuJoCo-shape B=1024/GPU H=256 C=14 500 Euler steps, last 10 knots"),
}
N_INPUT_SETS = 3          # rotated so that consecutive steps never re-read inputs from L2


def alg_flops_per_sde_step(w):
    """algorithmic multiply-adds x2 per batch row per solver step (layer-wise form)."""
    H, C, L, io, no = w["H"], w["C"], w["L"], w["io"], w["no"]
    if w["family"] == "tutorial":
        return 2 * C * H + 4 * H * H + 2 * (L + 1) * H * H + 2 * H * H
    ctl = io in (0, 2, 4, 6)
    emb = io in (2, 4, 6)
    tau = 2 if io in (3, 4, 5, 6) else 0
    f = (2 * C * H if ctl else 0) + 2 * (H + tau) * H + (4 * H * H if emb else 0) + (L - 1) * 2 * H * H + 2 * H * H
    if no in (18, 19):
        f += 2 * (H + 2) * H + 2 * H * H
    if no in (14, 15):
        f += 2 * (H + 2) * H
    return f


def alg_bytes_per_sde_step(w, n_out):
    """one 4C-float spline row per step + outputs and z0 amortised over S steps."""
    H, C, S = w["H"], w["C"], w["S"]
    ctl = w["family"] == "tutorial" or w["io"] in (0, 2, 4, 6)
    return (16 * C if ctl else 0) + 4.0 * H * n_out / S + 4.0 * H / S


def make_inputs(w, B, seed, keep_raw=None):
    """Synthetic inputs of SURVEY 8d on the CPU (fp32): integer knots, time channel + random-walk channels,
    Hermite/backward-difference coefficients (natural spline for c5), z0 ~ N(0, 0.1^2), random final_index."""
    from snsde_b200 import data
    g = torch.Generator().manual_seed(seed)
    S, C, H = w["S"], w["C"], w["H"]
    if w["family"] == "tutorial":
        times = torch.linspace(0, 1, S + 1)
    else:
        times = torch.arange(S + 1, dtype=torch.float32)
    x = (torch.randn(B, S + 1, C, generator=g) * 0.1).cumsum(1)
    x[..., 0] = times
    build = data.natural_cubic_coeffs if w["out"] == "tail10" else data.hermite_backward_difference_coeffs
    coeffs = build(x, times).contiguous()
    z0 = torch.randn(B, H, generator=g) * 0.1
    if w["out"] == "final_index":
        final_index = torch.randint(2, S + 1, (B,), generator=g)
    else:
        final_index = torch.full((B,), S, dtype=torch.long)
    if keep_raw is not None:
        keep_raw.append(x.contiguous())
    return times, coeffs, z0, final_index


def make_model(w, seed=0):
    from snsde_b200 import modules
    torch.manual_seed(seed)
    if w["family"] == "tutorial":
        return modules.TutorialLSDEParams(w["C"], w["H"], w["H"], w["L"])
    return modules.DiffusionModelParams(w["C"], w["H"], w["H"], w["L"], theta=1.0, sigma=1.0,
                                        input_option=w["io"], noise_option=w["no"])


def output_times(w, times):
    if w["out"] == "stream":
        return times
    if w["out"] == "tail10":
        return torch.cat([times[:1], times[-10:]])
    if w["out"] == "last":
        return times[[0, -1]]
    return None          # final_index: derived per batch


class ClockSampler:
    """nvidia-smi SM clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# CPU port of the reference path (the oracle), used for cpu_baseline and --impl reference
# ---------------------------------------------------------------------------------------------------
def cpu_reference_solver(w, B, seed=0):
    """Returns (callable running ONE solve, S_executed, description).  Mirrors what the reference does
    per forward: set_X, unique(final_index) output times, torchsde-style step loop with f/g evaluated
    op by op, Brownian increments drawn per step, gather."""
    from oracle import solver, vector_field, wrapper
    torch.manual_seed(seed)
    if w["family"] == "tutorial":
        m = vector_field.TutorialLSDEFunc(w["C"], w["H"], w["H"], w["L"])
    else:
        m = vector_field.DiffusionModel(w["C"], w["H"], w["H"], w["L"], input_option=w["io"], noise_option=w["no"])
    times, coeffs, z0, final_index = make_inputs(w, B, seed)

    class DrawnBM:                       # N(0, (t1-t0) I) per call, like BrownianInterval but without its tree
        def __call__(self, t0, t1):
            return torch.randn(B, w["H"]) * (t1 - t0).sqrt()

    def run():
        with torch.no_grad():
            if w["out"] == "final_index":
                return wrapper.classification_latent(m, times, coeffs, final_index, z0, DrawnBM(), method=w["method"])
            m.set_X(coeffs, times)
            dt = 1.0 / w["S"] if w["family"] == "tutorial" else solver.solver_dt(times)
            return solver.sdeint(m, z0, output_times(w, times), dt, DrawnBM(), method=w["method"])
    return run, w["S"]


def time_cpu_reference(w, B, budget_s=12.0, max_repeats=200, warmup=1):
    """Bounded sample: full solves of the workload until ~budget_s of CPU work (10-30 s band of the contract)."""
    torch.set_num_threads(os.cpu_count())
    run, S = cpu_reference_solver(w, B)
    for _ in range(warmup):
        run()
    ts = []
    while sum(ts) < budget_s and len(ts) < max_repeats:
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
    return B * S / statistics.median(ts), ts


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = w["B"]
    t0 = time.perf_counter()
    run, S = cpu_reference_solver(w, B)
    torch.set_num_threads(os.cpu_count())
    for _ in range(min(args.warmup, 1)):
        run()
    ts = []
    for _ in range(args.steps):
        a = time.perf_counter(); run(); ts.append(time.perf_counter() - a)
        if time.perf_counter() - t0 > 240:          # bounded: the whole arm ends within a few minutes
            break
    sec = sum(ts)
    value = B * S * len(ts) / sec
    sample = f"{len(ts)} full solves of the workload (B={B}, S={S}) on the CPU"
    line = {
        "impl": "reference", "metric": "SDE-steps/sec", "value": value, "unit": "SDE-steps/s",
        "n_gpus": args.gpus, "steps": len(ts), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sec / len(ts),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "rows_timed": B,
                   "note": "CPU port of the reference path (oracle/): torchsde/torchcde 0.2.5 are not installable offline"},
        "cpu_baseline": {"value": value, "unit": "SDE-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "SDE-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def _cpu_list(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus += list(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(local_rank):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU, so the host<->device
    copies of the e2e leg do not cross the socket interconnect (e2e erratic on the 8-GPU box).  Sources, in
    order: sysfs numa_node of the GPU's PCI function; the "CPU Affinity" column of `nvidia-smi topo -m`."""
    try:
        cpus, where = [], None
        pr = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(pr, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            f = pathlib.Path(f"/sys/bus/pci/devices/{bus}/numa_node")
            if f.exists():
                node = int(f.read_text())
                if node >= 0:
                    cpus = _cpu_list(pathlib.Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
                    where = f"numa node {node} (sysfs)"
        if not cpus:
            out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            lines = [l for l in out.splitlines() if l.strip()]
            hdr = next(l for l in lines if "CPU Affinity" in l)
            cols = [c.strip() for c in hdr.split("\t")]
            ci = cols.index("CPU Affinity")
            row = next(l for l in lines if l.startswith(f"GPU{local_rank}\t") or l.startswith(f"GPU{local_rank} "))
            cells = [c.strip() for c in row.split("\t")]
            cpus = _cpu_list(cells[ci]) if ci < len(cells) else []
            where = f"cpus {cells[ci]} (nvidia-smi topo)" if cpus else None
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return where
        return None
    except Exception as exc:                             # noqa: BLE001
        if os.environ.get("BENCH_DEBUG"):
            print(f"[bench] NUMA binding skipped: {exc!r}", file=sys.stderr)
        return None


def oracle_model_for(w, model):
    from oracle import vector_field
    if w["family"] == "tutorial":
        m = vector_field.TutorialLSDEFunc(w["C"], w["H"], w["H"], w["L"])
    else:
        m = vector_field.DiffusionModel(w["C"], w["H"], w["H"], w["L"], input_option=w["io"], noise_option=w["no"])
    m.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()})
    return m


def parity_gate(w, model, inputs_host, inputs_dev, plan, dt, dev, row_offset, precision, rows=48, horizon=None, tol=1e-4):
    """"parity gate run with every measurement".  The engine solves the FULL batch of the workload (same
    kernel, same launch shape as the timed solve, in-kernel Philox increments); the CPU oracle integrates the first
    `rows` rows on the materialised increments (rows never interact); rel err = max|z - z_oracle| / max(|z_oracle|, 1e-3)
    over those rows' outputs.  `horizon`: number of solver steps checked (None = the whole trajectory)."""
    import snsde_b200
    from oracle import solver, wrapper
    times, coeffs, z0, fi = inputs_host
    times_d, coeffs_d, z0_d, fi_d = inputs_dev
    S = w["S"] if horizon is None else min(horizon, w["S"])
    if S < w["S"]:
        times, coeffs, times_d, coeffs_d = times[:S + 1], coeffs[:, :S], times_d[:S + 1], coeffs_d[:, :S]
        fi, fi_d = fi.clamp(max=S), fi_d.clamp(max=S)
    model.set_X(coeffs_d, times_d)
    seed = 4242
    with torch.no_grad():
        if w["out"] == "final_index":
            ts, slots = snsde_b200.final_index_slots(times, fi)
            sp = plan.step_plan(ts, dt, times)
            got = plan.forward(z0_d, sp, coeffs=coeffs_d, row_slot=slots.to(torch.int32), seed=seed, row_offset=row_offset)
        else:
            ts = times if w["out"] == "stream" else (torch.cat([times[:1], times[-min(10, S):]]) if w["out"] == "tail10" else times[[0, -1]])
            sp = plan.step_plan(ts, dt, times)
            got = plan.forward(z0_d, sp, coeffs=coeffs_d, seed=seed, row_offset=row_offset)
        dW = snsde_b200.philox_increments(seed, sp, rows, w["H"], dev, row_offset=row_offset).cpu()
        flags = plan.status()
    got = got.cpu()
    m = oracle_model_for(w, model)
    bm = solver.BrownianTable(dW)
    with torch.no_grad():
        if w["out"] == "final_index":
            want = wrapper.classification_latent(m, times, coeffs[:rows], fi[:rows], z0[:rows], bm, method=w["method"])
            got = got[:rows]
        else:
            m.set_X(coeffs[:rows], times)
            want = solver.sdeint(m, z0[:rows], ts, dt, bm, method=w["method"])
            got = got[:, :rows]
    err = float((got - want).abs().max()) / max(float(want.abs().max()), 1e-3)
    return {"rel_err": err, "tol": tol, "ok": bool(err <= tol and not (flags & 1)), "rows": rows, "solver_steps": S,
            "batch_solved": int(z0.shape[0]), "increments": "in-kernel Philox, materialised for the oracle",
            "oracle": "oracle/ (CPU port; torchsde/torchcde parity unpinned, see oracle/__init__.py)",
            "fp16_range_flag": int(flags)}


class Workload:
    """Device-resident state of one benchmark workload on this rank."""

    def __init__(self, name, w, precision, dev, rank, world):
        import snsde_b200
        self.name, self.w, self.dev, self.rank, self.world = name, w, dev, rank, world
        B, H, S = w["B"], w["H"], w["S"]
        self.row_offset = rank * B
        self.model = make_model(w).to(dev)
        self.raw_host = []
        self.sets_host = [make_inputs(w, B, seed=1000 * rank + i, keep_raw=self.raw_host) for i in range(N_INPUT_SETS)]
        self.sets_dev = [tuple(t.to(dev) for t in s) for s in self.sets_host]
        with torch.no_grad():
            self.plan = snsde_b200.plan_for(self.model, w["method"], precision, dev)
        self.dt = 1.0 / S if w["family"] == "tutorial" else snsde_b200.solver_dt(self.sets_host[0][0].numpy())
        t0 = time.perf_counter()
        self.step_plans, self.slots = [], []
        ts_fixed = output_times(w, self.sets_dev[0][0])
        for (times, coeffs, z0, fi) in self.sets_host:
            if w["out"] == "final_index":
                ts, sl = snsde_b200.final_index_slots(times, fi)
                self.step_plans.append(self.plan.step_plan(ts, self.dt, times)); self.slots.append(sl.to(torch.int32).to(dev))
            else:
                self.step_plans.append(self.plan.step_plan(ts_fixed, self.dt, times)); self.slots.append(None)
        self.host_plan_ms = 1e3 * (time.perf_counter() - t0) / N_INPUT_SETS      # reported, not part of `value`
        self.n_out = self.step_plans[0].n_out if w["out"] != "final_index" else 1
        Bg = B * world
        shape = (Bg, H) if w["out"] == "final_index" else (self.n_out, Bg, H)
        # two gather buffers: the all-gather of solve i runs on a side stream under solve i+1
        self.gather = [torch.empty(shape, device=dev) for _ in range(2)]
        self.side = torch.cuda.Stream(device=dev) if world > 1 else None
        self.gather_done = [None, None]

    def step(self, i):
        import torch.distributed as dist
        from snsde_b200 import dist as sdist
        w, B, H = self.w, self.w["B"], self.w["H"]
        times, coeffs, z0, fi = self.sets_dev[i % N_INPUT_SETS]
        sp, sl = self.step_plans[i % N_INPUT_SETS], self.slots[i % N_INPUT_SETS]
        buf = self.gather[i % 2]
        main = torch.cuda.current_stream(self.dev)
        if self.gather_done[i % 2] is not None:
            main.wait_event(self.gather_done[i % 2])                 # the gather that last used this buffer (2 steps ago)
        lo = self.row_offset
        if w["out"] == "final_index":
            self.plan.forward(z0, sp, coeffs=coeffs, row_slot=sl, seed=i, row_offset=lo, out=buf[lo:lo + B])
        else:
            z = self.plan.forward(z0, sp, coeffs=coeffs, seed=i, row_offset=lo)
        if self.world > 1:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                if w["out"] == "final_index":
                    sdist.all_gather_rows(buf, self.rank, self.world)
                else:
                    z.record_stream(self.side)
                    tmp = torch.empty((self.world, self.n_out, B, H), device=self.dev)
                    dist.all_gather_into_tensor(tmp, z)
                    buf.view(self.n_out, self.world, B, H).copy_(tmp.transpose(0, 1))
                ev = torch.cuda.Event()
                ev.record(self.side)
                self.gather_done[i % 2] = ev
        return buf

    def drain(self):
        """Make the main stream wait for every outstanding all-gather (end of a timed region)."""
        if self.side is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)


def train_block(w, wl, steps, warmup, precision, dev, rows=16):
    """one training step of the workload through the public API - forward solve that saves the solver
    states, reverse-sweep kernel, weight-gradient GEMMs - timed on the device, plus a gradient parity gate:
    dL/dz0 of the first `rows` rows (L = sum z^2 is row-separable) against fp64 autograd through the oracle on the
    materialised increments."""
    import snsde_b200
    from oracle import solver
    B, H, S = w["B"], w["H"], w["S"]
    times, coeffs, z0, fi = wl.sets_host[0]
    times_d, coeffs_d, z0_d, fi_d = wl.sets_dev[0]
    model = wl.model
    model.set_X(coeffs_d, times_d)
    opt = torch.optim.Adam(model.parameters(), lr=1e-6)

    def one(i, with_opt):
        zz = z0_d.clone().requires_grad_(True)
        z = snsde_b200.solve_final(model, times, fi, zz, method="euler", seed=i, precision=precision, row_offset=wl.row_offset, dt=wl.dt)
        (z * z).sum().backward()
        if with_opt:
            opt.step()
        opt.zero_grad(set_to_none=True)
        return zz.grad

    out = {}
    for key, with_opt in (("fwd_bwd_ms", False), ("fwd_bwd_optimizer_ms", True)):
        for i in range(warmup):
            one(i, with_opt)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            one(i, with_opt)
        b.record()
        torch.cuda.synchronize()
        out[key] = a.elapsed_time(b) / steps
    # gradient parity (weights as they are now)
    seed = 777
    zz = z0_d.clone().requires_grad_(True)
    z = snsde_b200.solve_final(model, times, fi, zz, method="euler", seed=seed, precision=precision, row_offset=wl.row_offset, dt=wl.dt)
    (z * z).sum().backward()
    opt.zero_grad(set_to_none=True)
    got = zz.grad[:rows].cpu().double()
    plan = snsde_b200.plan_for(model, "euler", precision, dev)
    ts, _ = snsde_b200.final_index_slots(times, fi)
    dW = snsde_b200.philox_increments(seed, plan.step_plan(ts, wl.dt, times), rows, H, dev, row_offset=wl.row_offset).cpu().double()
    m = oracle_model_for(w, model).double()
    m.set_X(coeffs[:rows].double(), times.double())
    y0 = z0[:rows].double().requires_grad_(True)
    z_all = solver.sdeint_with_grad(m, y0, times.double(), wl.dt, solver.BrownianTable(dW))
    zo = z_all[fi[:rows], torch.arange(rows)]
    (zo * zo).sum().backward()
    err = float((got - y0.grad).abs().max()) / max(float(y0.grad.abs().max()), 1e-6)
    out.update({"value": B * S / (out["fwd_bwd_ms"] * 1e-3), "unit": "SDE-steps/s (forward + backward)", "rows_per_gpu": B, "solver_steps": S,
                "forward_kernel": plan.kernel, "backward_kernel": "fp32 reverse sweep + cuBLAS weight-gradient GEMMs",
                "grad_parity": {"what": "dL/dz0, L = sum z^2", "rel_err": err, "tol": 1e-4, "ok": bool(err <= 1e-4), "rows": rows,
                                "oracle": "fp64 autograd through oracle.solver (CPU), same Philox increments"}})
    return out


def roofline_block(name, w, plan, kern_ms, n_out_rows):
    hbm_peak, tf_peak, peak_src = peaks()
    B, S = w["B"], w["S"]
    abytes = alg_bytes_per_sde_step(w, n_out_rows)
    aflops = alg_flops_per_sde_step(w)
    sde_per_s_kernel = B * S / (kern_ms * 1e-3)                   # one launch, this GPU
    hbm_gbs = sde_per_s_kernel * abytes / 1e9
    tflops = sde_per_s_kernel * aflops / 1e12
    hbm_frac, tf_frac = hbm_gbs / hbm_peak, tflops / tf_peak
    if tf_frac >= hbm_frac:
        roof = {"bound": "tensor", "achieved": tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf_frac}
    else:
        roof = {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac}
    traffic = None
    tj = ROOT / "profiles" / "traffic.json"
    if tj.exists():
        rec = json.loads(tj.read_text()).get(name)
        if rec and rec["kernel"] == plan.kernel and B == WORKLOADS[name]["B"]:
            traffic = rec["bytes"]                 # measured once with ncu --set full (see profiles/)
    roof.update({"traffic": traffic, "alg_bytes_per_launch": abytes * B * S, "kernel": plan.kernel, "kernel_ms_per_launch": kern_ms, "peak_source": peak_src,
                 "alg_bytes_per_sde_step": abytes, "alg_flops_per_sde_step": aflops,
                 "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": hbm_frac},
                 "tensor": {"achieved_tflops": tflops, "peak_tflops": tf_peak, "frac": tf_frac},
                 "note": "latency-bound chain of small dependent GEMMs (SURVEY 8d); both fractions reported"})
    return roof


def time_resident(wl, steps, warmup, barrier, max_over_ranks, clock_index=None):
    """W warm-up + K timed device-resident solves (CUDA events on the launching stream; max over ranks)."""
    with torch.no_grad():
        for i in range(warmup):
            wl.step(i)
        wl.drain()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        end = torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(clock_index) if clock_index is not None else None
        if sampler:
            sampler.__enter__()
        barrier()
        launches0 = wl.plan.launches
        t_wall = time.perf_counter()
        for i in range(steps):
            evs[i][0].record()
            wl.step(i)
            evs[i][1].record()
        wl.drain()                                  # the last all-gathers are part of the timed region
        end.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        launches = wl.plan.launches - launches0
        if sampler and t_wall < 1.5:                # keep the GPU under load long enough for >= 10 clock samples
            # a FIXED number of extra solves agreed by all ranks: every solve carries a collective at N > 1, so a
            # time-based loop would leave the ranks with different collective counts (and hang the next one)
            n_extra = int(max_over_ranks(min(20000.0, 1.5 / max(t_wall / steps, 1e-5))))
            for j in range(n_extra):
                wl.step(j)
                if j % 8 == 7:
                    torch.cuda.synchronize()
            wl.drain()
            torch.cuda.synchronize()
        if sampler:
            sampler.__exit__()
    dev_ms = max_over_ranks(evs[0][0].elapsed_time(end))          # whole K-step region on the device, gathers included
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)   # the solve launches alone (roofline numerator)
    return dev_ms, kern_ms, launches, (sampler.summary() if sampler else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tc"])
    ap.add_argument("--rows", type=int, default=None, help="rows per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c3/c4/c5 blocks of the default line")
    ap.add_argument("--solver-steps", type=int, default=None, help="override the workload's S (launch-overhead sweeps)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.rows:
        w["B"] = args.rows
    if args.solver_steps:
        w["S"] = args.solver_steps
    if args.impl == "reference":
        return run_reference_arm(args, w)
    args.warmup = max(args.warmup, 3)
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:      # no-op when libsnsde.so is current (it ships with the snapshot)
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("snsde_build", ROOT / "stable-neural-sdes_b200" / "build.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as exc:                          # noqa: BLE001
            print(f"[bench] could not build libsnsde.so: {exc}", file=sys.stderr)

    import torch.distributed as dist
    import snsde_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    affinity0 = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa(local_rank)
    if world > 1:
        # The path's one collective moves 0.5 MB per rank and runs UNDER the next solve (side stream).  Cap NCCL's channel
        # count so its CTAs take few issue slots from the persistent solve kernel, but not below what the gather needs to
        # finish within a solve - measured at N=8 (ms per solve): default channels 0.325, 4 channels 0.301, 1 channel 0.621
        # (the gather itself becomes the bottleneck).  An explicit NCCL_MAX_NCHANNELS wins.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B, H, S = w["B"], w["H"], w["S"]
    Bg = B * world
    row_offset = rank * B
    wl = Workload(args.workload, w, args.precision, dev, rank, world)
    model, plan, dt, method, n_out = wl.model, wl.plan, wl.dt, w["method"], wl.n_out
    pinned = [tuple(t.pin_memory() for t in s) for s in wl.sets_host]
    raw_pinned = [x.pin_memory() for x in wl.raw_host]

    # parity gate of THIS measurement (rank 0's shard): full-batch launch vs the oracle on a row slice
    # (c1/c2: the whole trajectory; c3/c4/c5: the first 24 solver steps - their full horizons are ill-conditioned in fp32,
    # DESIGN 3 - with the full-horizon error reported beside it, not gated)
    parity = None
    if rank == 0:
        gate_h = None if args.workload in ("c1", "c2") else 24
        parity = parity_gate(w, model, wl.sets_host[0], wl.sets_dev[0], plan, dt, dev, row_offset, args.precision, horizon=gate_h)
        if gate_h is not None:
            full = parity_gate(w, model, wl.sets_host[0], wl.sets_dev[0], plan, dt, dev, row_offset, args.precision)
            parity["full_horizon_rel_err_not_gated"] = full["rel_err"]

    # end-to-end step: the public API with HOST (pinned) inputs; H2D + solve + D2H of the latents.  Steps are
    # double-buffered over two CUDA streams so the H2D of step i+1 overlaps the solve of step i; every step's
    # copies and result read-back are inside the timed region.
    e2e_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    host_outs = [torch.empty((B, H) if w["out"] == "final_index" else (n_out, B, H)).pin_memory() for _ in range(2)]

    from snsde_b200 import data as sdata
    build_dev = sdata.natural_coeffs_cuda if w["out"] == "tail10" else sdata.hermite_coeffs_cuda

    def enqueue_e2e(i, raw=False):
        times, coeffs, z0, fi = pinned[i % N_INPUT_SETS]
        t_d = times.to(dev, non_blocking=True)
        if raw:       # ship the raw path x [B,K,C] (4x fewer bytes) and build the coefficients on device
            c_d = build_dev(raw_pinned[i % N_INPUT_SETS].to(dev, non_blocking=True), t_d)
        else:
            c_d = coeffs.to(dev, non_blocking=True)
        z_d = z0.to(dev, non_blocking=True)
        model.set_X(c_d, t_d)
        # knots and final_index are consumed by the HOST side of the engine (step plan, unique/slot bookkeeping of
        # neuralsde.py:95-103): handing it the host copies keeps the enqueue free of device->host syncs, so the H2D of
        # the next step overlaps this step's solve (a device final_index costs a sync behind the 115 MB copy).
        if os.environ.get("BENCH_E2E_DEVICE_BOOKKEEPING"):        # A/B aid: the reference's calling convention (device tensors)
            times, fi = t_d, fi.to(dev, non_blocking=True)
        if w["out"] == "final_index":
            z = snsde_b200.solve_final(model, times, fi, z_d, method=method, seed=i, precision=args.precision,
                                       row_offset=row_offset, dt=dt)
        else:
            z = snsde_b200.sdeint(model, z_d, output_times(w, times), dt=dt, method=method, seed=i,
                                  precision=args.precision, row_offset=row_offset)
        if world > 1:
            buf = torch.empty((world, *z.shape), device=dev)
            dist.all_gather_into_tensor(buf, z)
        host_outs[i % 2].copy_(z, non_blocking=True)

    def run_e2e(n, raw=False):
        pending = None
        for i in range(n):
            st = e2e_streams[i % 2]
            with torch.cuda.stream(st):
                enqueue_e2e(i, raw)
            if pending is not None:
                pending.synchronize()              # step i-1 fully done (result is in host memory)
            pending = st
        pending.synchronize()
    host_out = host_outs[0]

    # per step: knots + coefficients + z0 (+ the int32 slot table derived from final_index on the host)
    h2d = sum(t.numel() * t.element_size() for t in pinned[0][:3]) + (pinned[0][3].numel() * 4 if w["out"] == "final_index" else 0)
    h2d_raw = h2d - pinned[0][1].numel() * 4 + (raw_pinned[0].numel() * 4 if raw_pinned else 0)
    d2h = host_out.numel() * 4

    dev_ms, kern_ms, launches, clocks = time_resident(wl, args.steps, args.warmup, barrier, max_over_ranks, clock_index=local_rank)

    with torch.no_grad():
        run_e2e(args.warmup)
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_raw_s = None
        if w["family"] == "benchmark":             # secondary: raw path over PCIe + coefficient construction on device
            run_e2e(args.warmup, raw=True)
            barrier()
            t0 = time.perf_counter()
            run_e2e(args.steps, raw=True)
            barrier()
            e2e_raw_s = max_over_ranks(time.perf_counter() - t0)

    value = Bg * S * args.steps / (dev_ms * 1e-3)
    e2e_value = Bg * S * args.steps / e2e_s

    # ---- the other BASELINE configs (c1 tutorial, 64 rows; c3 single GPU; c4 = B 4096 over 4 GPUs; c5 = B 8192 over 8 GPUs): device-resident
    # value + roofline + parity, 1024 (c3: 2048) rows per GPU at whatever N this run has ----
    configs = {}
    if args.workload == "c2" and not args.no_configs and not args.rows and not args.solver_steps:
        del pinned, raw_pinned
        for name in os.environ.get("BENCH_CONFIGS", "c1,c3,c4,c5").split(","):       # env: debugging aid (subset / order)
            cw = dict(WORKLOADS[name])
            cwl = Workload(name, cw, args.precision, dev, rank, world)
            # long-horizon trajectories of these models are ill-conditioned in fp32 (tests/test_fullsize_gpu.py):
            # the gate attached to the measurement checks the full-batch launch over the first 24 solver steps
            cpar = parity_gate(cw, cwl.model, cwl.sets_host[0], cwl.sets_dev[0], cwl.plan, cwl.dt, dev, rank * cw["B"],
                               args.precision, horizon=None if name == "c1" else 24) if rank == 0 else None
            c_dev_ms, c_kern_ms, c_launches, _ = time_resident(cwl, args.steps, args.warmup, barrier, max_over_ranks)
            if rank == 0:
                nrows = 1 if cw["out"] == "final_index" else cwl.n_out
                target_n = {"c1": 1, "c3": 1, "c4": 4, "c5": 8}[name]
                configs[name] = {
                    "workload": cw["desc"], "value": cw["B"] * world * cw["S"] * args.steps / (c_dev_ms * 1e-3),
                    "unit": "SDE-steps/s", "global_rows": cw["B"] * world, "rows_per_gpu": cw["B"], "solver_steps": cw["S"],
                    "ms_per_step": c_dev_ms / args.steps, "kernel_ms": c_kern_ms, "kernel": cwl.plan.kernel,
                    "fp32_variant": cwl.plan.variant,
                    "gpu_launches": c_launches, "host_step_plan_ms": cwl.host_plan_ms,
                    "baseline_config_n_gpus": target_n, "is_baseline_shape": world == target_n,
                    "roofline": roofline_block(name, cw, cwl.plan, c_kern_ms, nrows), "parity": cpar}
                if name == "c1" and not args.no_cpu_baseline:      # BASELINE configs[0]: the reference's own CPU-runnable case
                    aff = os.sched_getaffinity(0)
                    os.sched_setaffinity(0, affinity0)
                    v1, ts1 = time_cpu_reference(cw, cw["B"], budget_s=3.0)
                    os.sched_setaffinity(0, aff)
                    configs[name]["cpu_baseline"] = {"value": v1, "unit": "SDE-steps/s", "cores": os.cpu_count(), "kind": "port",
                                                     "sample": f"median of {len(ts1)} full solves (B={cw['B']}, S={cw['S']}); {sum(ts1):.1f}s of CPU work"}
            del cwl
            torch.cuda.empty_cache()

    train = None
    if args.workload == "c2" and not args.no_configs and not args.solver_steps and w["method"] == "euler":
        tb = train_block(w, wl, max(3, args.steps // 4), 2, args.precision, dev)      # every rank trains its shard; rank 0 reports
        train = tb if rank == 0 else None

    if rank == 0:
        n_out_rows = 1 if w["out"] == "final_index" else n_out
        roof = roofline_block(args.workload, w, plan, kern_ms, n_out_rows)
        cpu = None
        if not args.no_cpu_baseline:
            os.sched_setaffinity(0, affinity0)          # the CPU baseline gets every host core again
            v, ts_cpu = time_cpu_reference(w, B)
            cpu = {"value": v, "unit": "SDE-steps/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"median of {len(ts_cpu)} full solves of the workload (B={B}, S={S}); {sum(ts_cpu):.1f}s of CPU work"}
        line = {
            "metric": "SDE-steps/sec", "value": value, "unit": "SDE-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "global_rows": Bg, "rows_per_gpu": B,
                       "solver_steps": S, "precision": args.precision, "kernel": plan.kernel, "fp32_variant": plan.variant,
                       "parallelism": f"batch-shard x{world}" + (" + NCCL all-gather of final latents (side stream, double-buffered: "
                                                                 "the gather of solve i runs under solve i+1)" if world > 1 else ""),
                       "l2": f"rotating {N_INPUT_SETS} input sets ({N_INPUT_SETS * h2d / 1e6:.0f} MB) > 126 MB L2",
                       "brownian": "in-kernel Philox4x32-10", "numa_node": numa_node},
            "e2e": {"value": e2e_value, "unit": "SDE-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches, "host_step_plan_ms": wl.host_plan_ms,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
        }
        if e2e_raw_s is not None:
            # NOT the headline: a different host interface (the raw path x instead of the reference's precomputed
            # coefficients; they are built on device by snsde_hermite_coeffs / snsde_natural_coeffs)
            line["e2e_raw_path"] = {"value": Bg * S * args.steps / e2e_raw_s, "unit": "SDE-steps/s",
                                    "h2d_bytes_per_step": h2d_raw, "d2h_bytes_per_step": d2h,
                                    "ms_per_step": 1e3 * e2e_raw_s / args.steps,
                                    "note": "host buffers = raw path x[B,K,C]; coefficients built on device per step"}
        if configs:
            line["configs"] = configs
        if train:
            line["train"] = train
        print(json.dumps(line), flush=True)
    bad = rank == 0 and ((parity and not parity["ok"]) or any(c["parity"] and not c["parity"]["ok"] for c in configs.values())
                         or (train is not None and not train["grad_parity"]["ok"]))
    if world > 1:
        dist.destroy_process_group()
    if bad:
        raise SystemExit("[bench] parity gate failed (see `parity` in the JSON line): the measurement is not valid")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""HBM roofline of the coefficient-construction kernel on the c2 input shape, x4 batches.

    python profiles/bench_coeffs.py      -> one JSON line (achieved GB/s of algorithmic bytes vs MEASURED_PEAKS.json)
"""
import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import snsde_b200                                # noqa: E402
from snsde_b200 import data                      # noqa: E402

dev = torch.device("cuda", 0)
B, K, C = 4096, 201, 35
times = torch.arange(K, dtype=torch.float32, device=dev)
xs = [torch.randn(B, K, C, device=dev).cumsum(1) for _ in range(3)]        # 3 x 115 MB inputs, 3 x 461 MB outputs > L2
outs = [torch.empty(B, K - 1, 4 * C, device=dev) for _ in range(3)]
for i in range(3):
    data.hermite_coeffs_cuda(xs[i], times, out=outs[i])
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for i, (a, b) in enumerate(evs):
    a.record()
    data.hermite_coeffs_cuda(xs[i % 3], times, out=outs[i % 3])
    b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
alg = B * K * C * 4 + B * (K - 1) * 4 * C * 4
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); ref = data.hermite_backward_difference_coeffs(xs[0], times); t1.record(); torch.cuda.synchronize()
res = [json.dumps({"kernel": "hermite_coeffs_kernel", "shape": [B, K, C], "ms": ms, "alg_bytes": alg,
                  "achieved_gbs": alg / ms / 1e6, "peak_gbs": peak, "frac": alg / ms / 1e6 / peak,
                  "torch_ops_ms": t0.elapsed_time(t1), "max_abs_diff_vs_torch_ops": float((ref - outs[0]).abs().max())})]
del xs, outs, ref

# natural cubic spline on the c5 (MuJoCo-forecast) input shape: B=8192, K=501, C=14
B, K, C = 8192, 501, 14
times = torch.arange(K, dtype=torch.float32, device=dev)
xs = [torch.randn(B, K, C, device=dev).cumsum(1) for _ in range(2)]        # 2 x 230 MB in, 2 x 917 MB out > L2
outs = [torch.empty(B, K - 1, 4 * C, device=dev) for _ in range(2)]
for i in range(2):
    data.natural_coeffs_cuda(xs[i], times, out=outs[i])
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for i, (a, b) in enumerate(evs):
    a.record()
    data.natural_coeffs_cuda(xs[i % 2], times, out=outs[i % 2])
    b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
alg = B * K * C * 4 + B * (K - 1) * 4 * C * 4            # read x once, write the coefficients once
t0.record(); ref = data.natural_cubic_coeffs(xs[0], times); t1.record(); torch.cuda.synchronize()
res.append(json.dumps({"kernel": "natural_coeffs_kernel", "shape": [B, K, C], "ms": ms, "alg_bytes": alg,
                       "achieved_gbs": alg / ms / 1e6, "peak_gbs": peak, "frac": alg / ms / 1e6 / peak,
                       "torch_ops_ms": t0.elapsed_time(t1), "bit_identical_to_torch_ops": bool(torch.equal(ref, outs[0]))}))
print("\n".join(res))

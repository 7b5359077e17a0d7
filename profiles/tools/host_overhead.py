"""Host enqueue cost of one solve vs its device time (workload c1 by default): K back-to-back `plan.forward` calls,
wall time of the enqueue loop alone (no synchronisation inside) against CUDA-event time of the same K solves.

    python profiles/tools/host_overhead.py [workload] [K] [method]
"""
import pathlib
import sys
import time

import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 3:
    w["method"] = sys.argv[3]
wl = bench.Workload(name, w, "auto", dev, 0, 1)
with torch.no_grad():
    for i in range(10):
        wl.step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(K):
        wl.step(i)
    b.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
print(f"{name} ({w['method']}): kernel={wl.plan.kernel}/{wl.plan.variant} K={K} host enqueue {1e6 * t_enq / K:.1f} us/solve, "
      f"device (events) {1e3 * a.elapsed_time(b) / K:.1f} us/solve, wall {1e6 * t_all / K:.1f} us/solve")

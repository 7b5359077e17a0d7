#!/usr/bin/env python
"""Per-step critical-path breakdown of snsde_tc_kernel from its clock64 trace (debug aid).

    SNSDE_TC_TRACE=gpurun_out/trace.txt python bench.py --steps 1 --warmup 3 --no-cpu-baseline
    python profiles/trace_tc.py gpurun_out/trace.txt

Events (CTA 0; SM-local clock): epilogue thread 0 / MMA lane 0 / prefetch thread 0 / producer thread 0.
"""
import sys
import numpy as np

EV = ["EPI_ACC0", "EPI_LD0", "EPI_DONE0", "EPI_ACC1", "EPI_LD1", "EPI_DONE1", "EPI_SHADOW_END",
      "MMA_WAKE0", "MMA_COMMIT0", "MMA_WAKE1", "MMA_COMMIT1", "MMA_X_DONE", "PREP_DONE", "PROD_DONE",
      "EPI_PFULL", "EPI_PREPARED"]
t = np.loadtxt(sys.argv[1], dtype=np.int64)
t = t[20:-5]                                  # steady state
ix = {n: i for i, n in enumerate(EV)}
g = lambda n: t[:, ix[n]].astype(np.float64)
prev = lambda n: np.roll(g(n), 1)
rows = [
    ("step period (EPI_DONE1 -> EPI_DONE1)", g("EPI_DONE1") - prev("EPI_DONE1")),
    ("L0: hand-over(prev step) -> MMA warp awake", g("MMA_WAKE0") - prev("EPI_DONE1")),
    ("L0: MMA issue (wake -> commit issued)", g("MMA_COMMIT0") - g("MMA_WAKE0")),
    ("L0: commit issued -> epilogue sees accumulators", g("EPI_ACC0") - g("MMA_COMMIT0")),
    ("L0: TMEM loads", g("EPI_LD0") - g("EPI_ACC0")),
    ("L0: epilogue math + operand write + hand-over", g("EPI_DONE0") - g("EPI_LD0")),
    ("L1: hand-over -> MMA warp awake", g("MMA_WAKE1") - g("EPI_DONE0")),
    ("L1: MMA issue", g("MMA_COMMIT1") - g("MMA_WAKE1")),
    ("L1: commit issued -> epilogue sees accumulators", g("EPI_ACC1") - g("MMA_COMMIT1")),
    ("L1: TMEM loads", g("EPI_LD1") - g("EPI_ACC1")),
    ("L1: SDE update + operand write + hand-over", g("EPI_DONE1") - g("EPI_LD1")),
    ("shadow: emits etc. after hand-over", g("EPI_SHADOW_END") - g("EPI_DONE1")),
    ("MMA warp: X(t) segment after L1 commit", g("MMA_X_DONE") - g("MMA_COMMIT1")),
]
if t[:, ix["EPI_PFULL"]].any():
    rows += [
        ("shadow: step data visible after shadow end (pfull)", g("EPI_PFULL") - prev("EPI_SHADOW_END")),
        ("shadow: diffusion of the new state (prepare)", g("EPI_PREPARED") - g("EPI_PFULL")),
        ("shadow: prepared -> accumulators of L0 seen", g("EPI_ACC0") - g("EPI_PREPARED")),
        ("prefetch warp: slot s ready before epilogue needs it by", g("EPI_PFULL") - g("PREP_DONE")),
    ]
for name, d in rows:
    d = d[1:]
    print(f"{name:55s} median {np.median(d):8.0f}  p10 {np.percentile(d, 10):8.0f}  p90 {np.percentile(d, 90):8.0f} cycles")

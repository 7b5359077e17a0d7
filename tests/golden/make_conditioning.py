"""Measures, with the CPU oracle alone, how far the reference path evaluated in fp32 is from the same path evaluated
in fp64 at the BASELINE shapes (row slices; rows never interact) and freezes it in conditioning.json.

    python tests/golden/make_conditioning.py

Why: the long-horizon full-size GPU tests (tests/test_fullsize_gpu.py) use the tolerance
max(1e-4, 3 * ||ref_fp32 - ref_fp64||) for c3/c4/c5 - no fp32 implementation can sit closer to the exact trajectory
than the reference's own fp32 evaluation does.  This file is the evidence behind that relaxed tolerance;
the strict 1e-4 gate for those models is the teacher-forced / short-horizon tests.
"""
import copy
import json
import pathlib
import sys

import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parents[0]))

from oracle import solver, wrapper                      # noqa: E402
from test_fullsize_gpu import oracle_model, workload    # noqa: E402

CASES = {  # name: io, no, rows, H, C, S, method, natural, seed (as in tests/test_fullsize_gpu.py)
    "c2": (4, 17, 64, 128, 35, 200, "euler", False, 0),
    "c3": (6, 17, 64, 64, 35, 200, "milstein", False, 1),
    "c4": (3, 18, 64, 128, 21, 160, "euler", False, 2),
    "c5": (4, 17, 32, 256, 14, 500, "euler", True, 3),
}


def rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.abs().max()), 1.0)


def main():
    out = {}
    for name, (io, no, B, H, C, S, method, natural, seed) in CASES.items():
        m, times, coeffs, z0, fi = workload(io, no, B, H, C, S, seed=seed, natural=natural)
        o32 = oracle_model(m, io, no, C, H)
        o64 = copy.deepcopy(o32).double()
        dW = torch.randn(S, B, H, generator=torch.Generator().manual_seed(100 + seed))
        o32.set_X(coeffs, times)
        o64.set_X(coeffs.double(), times.double())
        a = solver.sdeint(o32, z0, times, 1.0, solver.BrownianTable(dW), method=method)
        b = solver.sdeint(o64, z0.double(), times.double(), 1.0, solver.BrownianTable(dW.double()), method=method)
        d = (a.double() - b).abs()
        scale = max(float(b.abs().max()), 1.0)
        out[name] = {"model": [io, no], "rows": B, "hidden": H, "solver_steps": S, "method": method,
                     "max_rel": float(d.max()) / scale, "median_rel": float(d.flatten().median()) / scale,
                     "q999_rel": float(d.flatten().kthvalue(int(0.999 * d.numel())).values) / scale,
                     "max_rel_first_24_steps": float(d[:25].max()) / max(float(b[:25].abs().max()), 1.0),
                     "state_scale": scale}
        print(name, out[name])
    (HERE / "conditioning.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""B200-native Neural-SDE integration engine (drop-in for the reference's torchsde.sdeint path).

Public surface (mirrors the reference's operator interface for this path):
    sdeint(sde, y0, ts, dt, method=..., bm=..., seed=...)   -> [len(ts), B, H]
    solve_final(sde, times, final_index, z0, ...)            -> [B, H]   (fused gather)
    patch(model)                                             -> swaps the engine into any of the reference's three NeuralSDE wrappers
    (under autograd, 'euler', 'srk' and elementwise-diffusion 'milstein': backward through the reverse-sweep kernels; methods euler / milstein / srk;
     a LatentSDE is patched / solved as its augmented system)
    BrownianIncrements(dW), philox_increments(...), Plan, build_step_plan, dist helpers
"""
from . import dist, packing, stepplan                                    # noqa: F401
from ._lib import EngineError, LIB_PATH                                  # noqa: F401
from .engine import (BrownianIncrements, Plan, final_index_slots, initial_state, patch, philox_increments,  # noqa: F401
                     plans_of, readout_head, sdeint, solve_final, wrapper_kind)
from .engine import _plan_for as plan_for                                # noqa: F401
from .stepplan import build_step_plan, solver_dt                          # noqa: F401

__all__ = ["sdeint", "solve_final", "patch", "Plan", "BrownianIncrements", "philox_increments", "plans_of", "plan_for", "wrapper_kind", "initial_state", "readout_head",
           "final_index_slots", "build_step_plan", "solver_dt", "dist", "EngineError"]

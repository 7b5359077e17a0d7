"""Oracle (test infrastructure): the callers either side of the solve.

Restates reference ``NeuralSDE.forward`` output-time bookkeeping
(benchmark_classification/models_sde/neuralsde.py:84-120), the forecasting variant
(benchmark_forecasting/models_sde/neuralsde.py:158-186) and the torch-ists variant
(torch-ists/torch_ists/diff_module/NSDE/nsde_model.py:76-84) on top of
``oracle.solver.sdeint``.  The read-out head / z0 producers are boundary neighbours
 and are NOT applied here: these functions return the latent ``z``.
"""
import torch

from .solver import sdeint, solver_dt


def output_times_for_final_index(times, final_index):
    """neuralsde.py:95-103.  Returns (ts, gather_index) with gather_index into dim 0 of z_t."""
    sorted_fi, inverse = final_index.unique(sorted=True, return_inverse=True)
    if 0 in sorted_fi:
        sorted_fi = sorted_fi[1:]
        gather_index = inverse
    else:
        gather_index = inverse + 1
    if len(times) - 1 in sorted_fi:
        sorted_fi = sorted_fi[:-1]
    ts = torch.cat([times[0].unsqueeze(0), times[sorted_fi], times[-1].unsqueeze(0)])
    return ts, gather_index


def classification_latent(func, times, coeffs, final_index, z0, bm, method="euler"):
    """z at each row's own final knot, ``[B, H]`` (neuralsde.py:86-116 minus the head)."""
    func.set_X(coeffs, times)
    ts, gidx = output_times_for_final_index(times, final_index)
    z_t = sdeint(func, z0, ts, solver_dt(times), bm, method=method)
    idx = gidx.unsqueeze(-1).expand(z_t.shape[1:]).unsqueeze(0)
    return z_t.gather(dim=0, index=idx).squeeze(0)


def streamed_latent(func, times, coeffs, z0, bm, method="euler"):
    """All knots, ``[B, K, H]`` (forecasting :178-183, torch-ists :81-83)."""
    func.set_X(coeffs, times)
    z_t = sdeint(func, z0, times, solver_dt(times), bm, method=method)
    return z_t.transpose(0, 1)

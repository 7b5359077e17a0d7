"""Host-side mirror of the reference's operator interface for the SDE solve.

``sdeint`` has the keyword surface of ``torchsde.sdeint`` as the reference calls it
(/root/reference/benchmark_classification/models_sde/neuralsde.py:78-82 and the tutorial
notebooks' cell 7) and returns the same ``[len(ts), B, H]`` tensor; ``patch`` swaps it into
a reference ``NeuralSDE`` at the ``_solve_sde_path`` seam (neuralsde.py:71-82; torch-ists
variant nsde_model.py:63-74) and fuses the ``final_index`` gather of ``forward`` (:91-116).

PyTorch is plumbing here (device memory, streams); all arithmetic of the path runs in the
CUDA kernels behind include/snsde.h.  There is no CPU/eager fallback.
"""
import ctypes
import types
import warnings
import weakref

import numpy as np
import torch

from . import _lib, packing, stepplan


class BrownianIncrements:
    """Explicit Brownian increments ``dW[S, B, H]`` (parity mode).  Also satisfies torchsde's
    ``bm(t0, t1)`` protocol (sequential replay), so the same object can be handed to real
    torchsde through the reference's ``**kwargs`` pass-through (neuralsde.py:84,105,82)."""

    def __init__(self, dW):
        self.dW = dW
        self._k = 0

    def __call__(self, t0, t1):
        w = self.dW[self._k]
        self._k += 1
        return w


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_HOST_COPIES = {}


def _host_array(t):
    """fp32 numpy copy of a small tensor; device tensors are cached per (object, version) so
    steady-state calls do not synchronise."""
    if isinstance(t, np.ndarray):
        return np.ascontiguousarray(t, dtype=np.float32)
    if not t.is_cuda:
        return t.detach().to(torch.float32).contiguous().numpy()
    key = id(t)
    hit = _HOST_COPIES.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    arr = t.detach().to(torch.float32).cpu().contiguous().numpy()
    if len(_HOST_COPIES) > 256:
        _HOST_COPIES.clear()
    _HOST_COPIES[key] = (weakref.ref(t), t._version, arr)
    return arr


class Plan:
    """Owns one ``snsde_plan`` (model descriptor + device weight images) on one device."""

    def __init__(self, desc, method="euler", precision="auto", device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.EngineError("snsde: no CUDA device - this engine has no CPU fallback")
        if method not in _lib.METHOD:
            raise ValueError(f"snsde: method {method!r} not implemented (euler, milstein; 'srk' is future work)")
        if precision not in _lib.PRECISION:
            raise ValueError(f"snsde: precision {precision!r} not in {sorted(_lib.PRECISION)}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.desc = dict(desc)
        self.method, self.precision = method, precision
        self._cdesc = _lib.ModelDesc(method=_lib.METHOD[method], precision=_lib.PRECISION[precision], **desc)
        h = ctypes.c_void_p()
        _lib.check(self.lib.snsde_plan_create(ctypes.byref(self._cdesc), self.device.index or 0, ctypes.byref(h)))
        self._h = h
        self._plans = {}
        self.weights_version = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.snsde_plan_destroy(h)

    @property
    def hidden(self):
        return self.desc["hidden"]

    @property
    def uses_control(self):
        return self.desc["family"] == _lib.FAMILY_TUTORIAL_LSDE or self.desc["input_option"] in (0, 2, 4, 6)

    @property
    def kernel(self):
        return {0: "fma_fp32", 1: "tcgen05", 2: "tcgen05_general"}[_lib.check(self.lib.snsde_plan_kernel_kind(self._h))]

    @property
    def launches(self):
        return int(self.lib.snsde_plan_launch_count(self._h))

    def status(self):
        """Sticky device flags since the last call (synchronises the current stream): bit 0 = a tensor-core
        kernel saturated an operand beyond the fp16 range; rerun with ``precision='fp32'``."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        return _lib.check(self.lib.snsde_plan_status(self._h, ctypes.c_void_p(stream)))

    def set_weights(self, blob):
        blob = blob.detach().to(torch.float32).contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.snsde_plan_set_weights(self._h, _ptr(blob), blob.numel(), int(blob.is_cuda),
                                                   ctypes.c_void_p(stream)))

    def load_from(self, sde):
        ver = packing.weights_version(sde)
        if ver != self.weights_version:
            self.set_weights(packing.pack(sde, self.desc))
            self.weights_version = ver

    def forward(self, y0, plan, coeffs=None, row_slot=None, dW=None, seed=0, row_offset=0, out=None):
        """Enqueue one solve on the current stream.  Returns ``[n_out, B, H]`` or, with
        ``row_slot`` (int32 ``[B]``), the fused ``[B, H]`` gather."""
        dev = self.device
        H = self.hidden
        if y0.dim() != 2 or y0.shape[1] != H:
            raise ValueError(f"snsde: y0 must be [B, {H}], got {tuple(y0.shape)}")
        if y0.device != dev:
            raise ValueError(f"snsde: y0 is on {y0.device}, plan is on {dev}")
        y0 = y0.detach().to(torch.float32).contiguous()
        B = y0.shape[0]
        stride = 0
        if self.uses_control:
            C = self.desc["input_channels"]
            if coeffs is None:
                raise ValueError("snsde: this model reads the control path; call set_X(coeffs, times) first")
            if coeffs.dim() != 3 or coeffs.shape[0] != B or coeffs.shape[2] != 4 * C or coeffs.shape[1] != plan.n_knots - 1:
                raise ValueError(f"snsde: coeffs must be [B={B}, K-1={plan.n_knots - 1}, 4C={4 * C}], got {tuple(coeffs.shape)}")
            if coeffs.device != dev:
                raise ValueError("snsde: coeffs and y0 must be on the same device")
            coeffs = coeffs.detach().to(torch.float32)
            if coeffs.stride(2) != 1 or coeffs.stride(1) != 4 * C or coeffs.data_ptr() % 16 or coeffs.stride(0) % 4:
                coeffs = coeffs.contiguous()
            stride = coeffs.stride(0)
        else:
            coeffs = None
        if dW is not None:
            if tuple(dW.shape) != (plan.n_steps, B, H):
                raise ValueError(f"snsde: dW must be [S={plan.n_steps}, B={B}, H={H}], got {tuple(dW.shape)}")
            dW = dW.detach().to(device=dev, dtype=torch.float32).contiguous()
        if row_slot is not None:
            row_slot = row_slot.to(device=dev, dtype=torch.int32).contiguous()
            if tuple(row_slot.shape) != (B,):
                raise ValueError("snsde: row_slot must be [B]")
            shape = (B, H)
        else:
            shape = (plan.n_out, B, H)
        if out is None:
            out = torch.empty(shape, device=dev, dtype=torch.float32)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise ValueError(f"snsde: out must be a contiguous fp32 {shape} tensor on {dev}")
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.lib.snsde_forward(
            self._h, _ptr(coeffs), stride, plan.n_knots, _ptr(y0), B,
            ctypes.c_void_p(plan.steps.ctypes.data), plan.n_steps,
            ctypes.c_void_p(plan.emits.ctypes.data), len(plan.emits), plan.n_init_emits, plan.n_out,
            _ptr(row_slot), _ptr(dW), ctypes.c_uint64(seed & (2 ** 64 - 1)), ctypes.c_uint64(row_offset),
            _ptr(out), ctypes.c_void_p(stream)))
        return out

    def step_plan(self, ts, dt, knots):
        ts_h = _host_array(ts)
        kn_h = _host_array(knots) if (knots is not None and self.uses_control) else None
        key = (ts_h.tobytes(), float(dt), None if kn_h is None else kn_h.tobytes())
        sp = self._plans.get(key)
        if sp is None:
            if len(self._plans) > 64:
                self._plans.clear()
            sp = self._plans[key] = stepplan.build_step_plan(ts_h, dt, kn_h)
        return sp


def philox_increments(seed, plan, B, H, device, row_offset=0):
    """``dW[S, B, H]`` exactly as the kernels draw them for ``seed`` (for the oracle)."""
    lib = _lib.load()
    dev = torch.device(device)
    dW = torch.empty((plan.n_steps, B, H), device=dev, dtype=torch.float32)
    sq = np.ascontiguousarray(plan.steps["sqrt_h"])
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.snsde_philox_fill(ctypes.c_uint64(seed), ctypes.c_uint64(row_offset), plan.n_steps, B, H,
                                     ctypes.c_void_p(sq.ctypes.data), _ptr(dW), dev.index or 0,
                                     ctypes.c_void_p(stream)))
    return dW


def _plan_for(sde, method, precision, device):
    cache = sde.__dict__.setdefault("_snsde_plans", {})
    key = (method, precision, str(device))
    plan = cache.get(key)
    if plan is None:
        plan = cache[key] = Plan(packing.describe(sde), method=method, precision=precision, device=device)
    plan.load_from(sde)
    return plan


def _check_sde(sde, y0):
    if getattr(sde, "sde_type", "ito") != "ito" or getattr(sde, "noise_type", "diagonal") != "diagonal":
        raise ValueError("snsde: only Ito SDEs with diagonal noise are supported (neuralsde.py:137-138)")
    if torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in sde.parameters())):
        raise RuntimeError("snsde: the engine is forward-only (SURVEY 8f1); call it under torch.no_grad() "
                           "- refusing to silently drop gradients")


def _increments(bm, plan, B, H, device):
    if bm is None:
        return None
    if hasattr(bm, "dW"):
        return bm.dW
    rows = [bm(torch.tensor(float(s["t0"])), torch.tensor(float(s["t0"]) + float(s["h"]))) for s in plan.steps]
    return torch.stack(rows).to(device) if rows else torch.empty((0, B, H), device=device)


def _random_seed():
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def sdeint(sde, y0, ts, dt=1e-3, method=None, options=None, bm=None, seed=None, precision="auto",
           row_offset=0, names=None, out=None, **unused_kwargs):
    """Drop-in for ``torchsde.sdeint(sde, y0, ts, dt=..., method=...)`` on this path.

    ``sde`` is a reference ``Diffusion_model`` (or the tutorial ``NeuralLSDEFunc``) on which
    ``set_X(coeffs, times)`` has been called; the engine reads ``sde.coeffs``, ``sde.times`` and
    ``state_dict()`` and never calls Python ``f``/``g``.  ``options`` is accepted and ignored, as
    torchsde's Euler ignores ``options['dt']`` (neuralsde.py:39-46).  ``bm=None`` draws
    increments in-kernel (Philox, ``seed``); ``bm=BrownianIncrements(dW)`` replays a table.
    """
    if unused_kwargs:
        warnings.warn(f"Unexpected arguments {sorted(unused_kwargs)}")          # torchsde does the same
    if names is not None:
        raise ValueError("snsde: `names` remapping is not supported")
    method = "euler" if method is None else method
    _check_sde(sde, y0)
    plan = _plan_for(sde, method, precision, y0.device)
    sp = plan.step_plan(ts, dt, getattr(sde, "times", None))
    dW = _increments(bm, sp, y0.shape[0], plan.hidden, y0.device)
    if dW is None and seed is None:
        seed = _random_seed()
    return plan.forward(y0, sp, coeffs=getattr(sde, "coeffs", None), dW=dW, seed=seed or 0,
                        row_offset=row_offset, out=out)


def final_index_slots(times, final_index):
    """Output times and per-row slot for ``final_index`` - the bookkeeping of neuralsde.py:95-103."""
    uniq, inverse = torch.unique(final_index, sorted=True, return_inverse=True)
    has0 = bool((uniq[0] == 0).item())
    slots = inverse if has0 else inverse + 1
    interior = uniq[1:] if has0 else uniq
    if interior.numel() and int(interior[-1]) == len(times) - 1:
        interior = interior[:-1]
    ts = torch.cat([times[:1], times[interior], times[-1:]])
    return ts, slots


def solve_final(sde, times, final_index, z0, method=None, bm=None, seed=None, precision="auto",
                row_offset=0, out=None, dt=None):
    """``z`` at each row's own final knot, ``[B, H]``: sdeint + gather of neuralsde.py:105-116 fused
    (each row keeps only its own slot; the ``[n_unique, B, H]`` intermediate is never written)."""
    method = "euler" if method is None else method
    _check_sde(sde, z0)
    plan = _plan_for(sde, method, precision, z0.device)
    ts, slots = final_index_slots(times, final_index)
    sp = plan.step_plan(ts, stepplan.solver_dt(_host_array(times)) if dt is None else dt, times)
    dW = _increments(bm, sp, z0.shape[0], plan.hidden, z0.device)
    if dW is None and seed is None:
        seed = _random_seed()
    return plan.forward(z0, sp, coeffs=getattr(sde, "coeffs", None), row_slot=slots, dW=dW, seed=seed or 0,
                        row_offset=row_offset, out=out)


_ENGINE_KW = ("bm", "seed", "precision", "row_offset")


def patch(model, fuse_final_index=True):
    """Swap the engine into a reference ``NeuralSDE``-style module, in place.

    * ``model._solve_sde_path`` (both signatures: ``(times, ts, z0, kwargs)`` of the benchmark
      classes and ``(times, y0, kwargs)`` of torch-ists) calls :func:`sdeint`;
    * with ``fuse_final_index`` and a classification-style ``forward(times, coeffs,
      final_index, z0=None, stream=False, **kw)``, the non-stream branch calls
      :func:`solve_final` so the gather happens in the kernel.
    Returns ``model``.
    """
    func = model.func

    def _solve(self, times, *rest):
        if len(rest) == 3:
            ts, z0, kwargs = rest
        else:
            z0, kwargs = rest
            ts = times
        kwargs = dict(kwargs)
        if "method" not in kwargs:
            if len(rest) == 2:
                # torch-ists NeuralSDE defaults to 'srk' (nsde_model.py:67), which the engine does not implement:
                # refuse rather than silently integrating with a different scheme
                raise ValueError("snsde: this wrapper's default method is 'srk' (not implemented); pass method='euler' "
                                 "or method='milstein' explicitly")
            kwargs["method"] = "euler"                       # benchmark wrappers' default (neuralsde.py:75)
        if kwargs["method"] == "srk":
            raise ValueError("snsde: method 'srk' is not implemented (SURVEY 8f2); use 'euler' or 'milstein'")
        dt = stepplan.solver_dt(_host_array(times))
        return sdeint(func, z0, ts, dt=dt, **kwargs)

    model._solve_sde_path = types.MethodType(_solve, model)

    if fuse_final_index and hasattr(model, "_prepare_initial_state") and hasattr(model, "linear"):
        def _forward(self, times, coeffs, final_index, z0=None, stream=False, **kwargs):
            if isinstance(coeffs, (tuple, list)):
                coeffs = coeffs[0] if len(coeffs) == 1 else torch.cat(list(coeffs), dim=-1)
            func.set_X(coeffs, times)
            z0 = self._prepare_initial_state(times, z0)
            eng = {k: kwargs.pop(k) for k in _ENGINE_KW if k in kwargs}
            method = kwargs.pop("method", None)
            kwargs.pop("options", None)
            if stream:
                z_t = sdeint(func, z0, times, dt=stepplan.solver_dt(_host_array(times)), method=method, **eng, **kwargs)
                z_t = z_t.transpose(0, 1)
            else:
                z_t = solve_final(func, times, final_index, z0, method=method, **eng)
            return self.linear(z_t)

        model.forward = types.MethodType(_forward, model)
    return model

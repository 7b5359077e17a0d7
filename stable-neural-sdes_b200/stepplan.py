"""Host step plan: replays torchsde 0.2.5's fixed-step loop in float32 and tabulates, per
step, everything that depends on time only (so the kernel never syncs with the host).

Mirrors ``BaseSDESolver.integrate`` as driven from the reference
(/root/reference/benchmark_classification/models_sde/neuralsde.py:78-82):
``next_t = min(curr_t + dt, ts[-1])`` in ``ts.dtype`` arithmetic, ``while curr_t < out_t``;
outputs are linear interpolations between the two solver states bracketing ``ts[i]``.
The spline interval/fraction follow torchcde ``CubicSpline._interpret_t``
(``clamp(bucketize(t, knots) - 1, 0, K-2)``, ``frac = t - knots[idx]``).
"""
import numpy as np

STEP_DTYPE = np.dtype([("t0", "<f4"), ("h", "<f4"), ("sqrt_h", "<f4"), ("sin_t0", "<f4"), ("cos_t0", "<f4"),
                       ("interval", "<i4"), ("frac", "<f4"), ("emit_begin", "<i4"), ("emit_end", "<i4"),
                       ("reserved", "<i4")])
EMIT_DTYPE = np.dtype([("slot", "<i4"), ("w_prev", "<f4"), ("w_curr", "<f4")])
assert STEP_DTYPE.itemsize == 40 and EMIT_DTYPE.itemsize == 12


class StepPlan:
    __slots__ = ("steps", "emits", "n_init_emits", "n_out", "n_knots")

    def __init__(self, steps, emits, n_init_emits, n_out, n_knots):
        self.steps, self.emits = steps, emits
        self.n_init_emits, self.n_out, self.n_knots = n_init_emits, n_out, n_knots

    @property
    def n_steps(self):
        return len(self.steps)


def solver_dt(knots):
    """``dt = max(min(diff(times)), 1e-3)`` - neuralsde.py:32-33 (float32 difference)."""
    knots = np.asarray(knots, dtype=np.float32)
    return max(float((knots[1:] - knots[:-1]).min()), 1e-3)


def build_step_plan(ts, dt, knots=None):
    ts = np.ascontiguousarray(ts, dtype=np.float32).reshape(-1)
    if ts.size < 1:
        raise ValueError("ts must hold at least one time")
    if ts.size > 1 and not np.all(ts[1:] > ts[:-1]):
        raise ValueError("Evaluation times `ts` must be strictly increasing.")
    if not dt > 0:
        raise ValueError("dt must be positive")
    dt32 = np.float32(dt)
    last = ts[-1]
    t0s, t1s, ranges = [], [], []
    emits = [(0, 0.0, 1.0)]
    curr = prev = ts[0]
    for i in range(1, ts.size):
        out_t = ts[i]
        while curr < out_t:
            nxt = min(np.float32(curr + dt32), last)
            if not nxt > curr:
                raise ValueError("dt is too small to advance float32 time")
            t0s.append(curr)
            t1s.append(nxt)
            ranges.append([len(emits), len(emits)])
            prev, curr = curr, nxt
        span = np.float32(curr - prev)
        emits.append((i, np.float32(curr - out_t) / span, np.float32(out_t - prev) / span))
        ranges[-1][1] = len(emits)
    S = len(t0s)
    steps = np.zeros(S, dtype=STEP_DTYPE)
    if S:
        t0 = np.asarray(t0s, dtype=np.float32)
        t1 = np.asarray(t1s, dtype=np.float32)
        steps["t0"] = t0
        steps["h"] = t1 - t0
        steps["sqrt_h"] = np.sqrt(steps["h"])
        steps["sin_t0"] = np.sin(t0)
        steps["cos_t0"] = np.cos(t0)
        rg = np.asarray(ranges, dtype=np.int32)
        steps["emit_begin"], steps["emit_end"] = rg[:, 0], rg[:, 1]
        if knots is not None:
            knots = np.ascontiguousarray(knots, dtype=np.float32).reshape(-1)
            idx = np.clip(np.searchsorted(knots, t0, side="left") - 1, 0, knots.size - 2)
            steps["interval"] = idx
            steps["frac"] = t0 - knots[idx]
    em = np.zeros(len(emits), dtype=EMIT_DTYPE)
    em["slot"] = [e[0] for e in emits]
    em["w_prev"] = [e[1] for e in emits]
    em["w_curr"] = [e[2] for e in emits]
    return StepPlan(steps, em, 1, int(ts.size), 0 if knots is None else int(np.size(knots)))

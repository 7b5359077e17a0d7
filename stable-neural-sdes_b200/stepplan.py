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
POINT_DTYPE = np.dtype([("t", "<f4"), ("sin_t", "<f4"), ("cos_t", "<f4"), ("frac", "<f4"), ("interval", "<i4")])
assert STEP_DTYPE.itemsize == 40 and EMIT_DTYPE.itemsize == 12 and POINT_DTYPE.itemsize == 20
SRK_POINT_FRACTIONS = (0.0, 0.25, 0.5, 1.0)      # SRID2: f at t0 + {0, 1, 1/2} h, g at t0 + {0, 1/4, 1} h


class StepPlan:
    __slots__ = ("steps", "emits", "n_init_emits", "n_out", "n_knots", "points", "_dense")

    def __init__(self, steps, emits, n_init_emits, n_out, n_knots, points=None):
        self.steps, self.emits = steps, emits
        self.n_init_emits, self.n_out, self.n_knots = n_init_emits, n_out, n_knots
        self.points = points                     # [S, 4] snsde_point (method 'srk') or None
        self._dense = None

    @property
    def n_steps(self):
        return len(self.steps)

    def dense(self):
        """The same steps, emitting EVERY solver state: slot s+1 = state after step s (``[S+1, B, H]``).
        The training path saves these states for the reverse sweep and forms the requested outputs from them."""
        if self._dense is None:
            S = self.n_steps
            steps = self.steps.copy()
            steps["emit_begin"] = np.arange(1, S + 1, dtype=np.int32)
            steps["emit_end"] = np.arange(2, S + 2, dtype=np.int32)
            em = np.zeros(S + 1, dtype=EMIT_DTYPE)
            em["slot"] = np.arange(S + 1, dtype=np.int32)
            em["w_curr"] = 1.0
            self._dense = StepPlan(steps, em, 1, S + 1, self.n_knots, self.points)
        return self._dense

    def output_map(self):
        """How the requested outputs read the dense states: ``out[slot] = w_prev * Y[k] + w_curr * Y[k + 1]``
        with ``k`` the step that produced the emit (``k = -1``: the initial state, out = Y[0]).
        Returns (slot, k, w_prev, w_curr) arrays over the emits."""
        E = len(self.emits)
        k = np.full(E, -1, dtype=np.int64)
        for s in range(self.n_steps):
            k[self.steps["emit_begin"][s]:self.steps["emit_end"][s]] = s
        return self.emits["slot"].astype(np.int64), k, self.emits["w_prev"].copy(), self.emits["w_curr"].copy()


def solver_dt(knots):
    """``dt = max(min(diff(times)), 1e-3)`` - neuralsde.py:32-33 (float32 difference)."""
    knots = np.asarray(knots, dtype=np.float32)
    return max(float((knots[1:] - knots[:-1]).min()), 1e-3)


def srk_points(steps, knots=None):
    """Evaluation points of the SRID2 stages per step, ``[S, 4]``: ``t0 + c * h`` for c in (0, 1/4, 1/2, 1) in
    float32 arithmetic (torchsde forms ``t0 + C[j] * dt`` with 0-d float32 tensors), their time features and
    spline interval / fraction."""
    S = len(steps)
    pts = np.zeros((S, len(SRK_POINT_FRACTIONS)), dtype=POINT_DTYPE)
    for i, c in enumerate(SRK_POINT_FRACTIONS):
        t = (steps["t0"] + (np.float32(c) * steps["h"]).astype(np.float32)).astype(np.float32)
        pts["t"][:, i] = t
        pts["sin_t"][:, i] = np.sin(t)
        pts["cos_t"][:, i] = np.cos(t)
        if knots is not None and S:
            kn = np.ascontiguousarray(knots, dtype=np.float32).reshape(-1)
            idx = np.clip(np.searchsorted(kn, t, side="left") - 1, 0, kn.size - 2)
            pts["interval"][:, i] = idx
            pts["frac"][:, i] = t - kn[idx]
    return np.ascontiguousarray(pts)


def _step_grid(ts, dt32):
    """Step boundaries ``b[0..S]`` of torchsde's fixed-step loop in float32: ``b[k+1] = min(fl32(b[k] + dt), ts[-1])`` until
    ``ts[-1]`` is reached.  ``np.add.accumulate`` on float32 adds sequentially, i.e. exactly like the loop."""
    first, last = ts[0], ts[-1]
    parts, curr, total = [], first, 0
    while True:                                               # chunks of <= 2^20 steps (a restart from `curr` continues the same sums)
        est = (float(last) - float(curr)) / float(dt32)
        n = int(min(max(np.ceil(est) + 2, 2), 1 << 20))
        arr = np.full(n + 1, dt32, dtype=np.float32)
        arr[0] = curr
        grid = np.add.accumulate(arr, dtype=np.float32)
        hit = np.nonzero(grid >= last)[0]
        end = int(hit[0]) if hit.size else n
        if end and not np.all(grid[1:end + 1] > grid[:end]):
            raise ValueError("dt is too small to advance float32 time")
        parts.append(grid[:end] if hit.size else grid[:n])
        total += end
        if hit.size:
            break
        if total > (1 << 26):
            raise ValueError("more than 2^26 solver steps between ts[0] and ts[-1]")
        curr = grid[n]
    b = np.concatenate(parts + [np.asarray([last], dtype=np.float32)])
    return b


def build_step_plan(ts, dt, knots=None, method="euler"):
    """Float32 replay of torchsde's ``BaseSDESolver.integrate`` bookkeeping (``next_t = min(curr_t + dt, ts[-1])``,
    ``while curr_t < out_t``, linear interpolation of the two states bracketing an output time), vectorised: the grid is a
    sequential float32 accumulation, an output time belongs to the first step that reaches it."""
    ts = np.ascontiguousarray(ts, dtype=np.float32).reshape(-1)
    if ts.size < 1:
        raise ValueError("ts must hold at least one time")
    if ts.size > 1 and not np.all(ts[1:] > ts[:-1]):
        raise ValueError("Evaluation times `ts` must be strictly increasing.")
    if not dt > 0:
        raise ValueError("dt must be positive")
    dt32 = np.float32(dt)
    n_out = int(ts.size)
    em = np.zeros(n_out, dtype=EMIT_DTYPE)
    em["slot"] = np.arange(n_out, dtype=np.int32)
    em["w_curr"][0] = 1.0
    if n_out == 1:
        steps = np.zeros(0, dtype=STEP_DTYPE)
    else:
        b = _step_grid(ts, dt32)
        S = b.size - 1
        t0, t1 = b[:-1], b[1:]
        out_t = ts[1:]
        k = np.searchsorted(t1, out_t, side="left")              # the step whose end first reaches the output time
        prev, curr = t0[k], t1[k]
        span = (curr - prev).astype(np.float32)
        em["w_prev"][1:] = (curr - out_t).astype(np.float32) / span
        em["w_curr"][1:] = (out_t - prev).astype(np.float32) / span
        steps = np.zeros(S, dtype=STEP_DTYPE)
        steps["t0"] = t0
        steps["h"] = t1 - t0
        steps["sqrt_h"] = np.sqrt(steps["h"])
        steps["sin_t0"] = np.sin(t0)
        steps["cos_t0"] = np.cos(t0)
        idx_steps = np.arange(S)
        steps["emit_begin"] = 1 + np.searchsorted(k, idx_steps, side="left")     # emits are ordered by their step
        steps["emit_end"] = 1 + np.searchsorted(k, idx_steps, side="right")
        if knots is not None:
            knots = np.ascontiguousarray(knots, dtype=np.float32).reshape(-1)
            idx = np.clip(np.searchsorted(knots, t0, side="left") - 1, 0, knots.size - 2)
            steps["interval"] = idx
            steps["frac"] = t0 - knots[idx]
    points = srk_points(steps, knots) if method == "srk" else None
    return StepPlan(steps, em, 1, n_out, 0 if knots is None else int(np.size(knots)), points)


def build_step_plan_loop(ts, dt, knots=None, method="euler"):
    """The same plan built step by step, statement for statement like the solver loop (kept as the checker of the
    vectorised builder, tests/test_host_cpu.py)."""
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
    points = srk_points(steps, knots) if method == "srk" else None
    return StepPlan(steps, em, 1, int(ts.size), 0 if knots is None else int(np.size(knots)), points)

"""Oracle (test infrastructure): Philox4x32-10 counter RNG bits in numpy.

The engine draws Brownian increments in-kernel from Philox4x32-10 (Salmon et al.,
"Parallel random numbers: as easy as 1, 2, 3", SC'11 - the same generator as
curand's ``curandStatePhilox4_32_10_t``), keyed for world-size invariance:

    key     = (seed_lo, seed_hi)
    counter = (feature j, global_row >> 2, step s, 0x534E5344 "SNSD")
    -> 4 x u32 -> 4 normals for global rows 4p..4p+3 of feature j at step s.

The uniform->normal map (Box-Muller on MUFU intrinsics) is not bit-reproducible on
a CPU, so bit parity of increments is taken from the engine's own
``snsde_philox_fill`` kernel; this replica pins the *integer* part (known-answer
vectors of the Random123 distribution) and gives a double-precision reference of the
normal map for statistical/ulp-level checks.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
STREAM_TAG = 0x534E5344


def philox4x32_10(counter, key):
    """counter: uint32[..., 4], key: uint32[..., 2] -> uint32[..., 4]."""
    c = [np.asarray(counter[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c[0].astype(np.uint64)
            p1 = PHILOX_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def increment_bits(seed, step, rows, H):
    """u32 bits [len(rows), H, 4-lane-selected] for global ``rows`` at ``step``: returns the
    single u32 pair (r0,r1) words feeding the normal of each (row, j)."""
    rows = np.asarray(rows, dtype=np.int64)
    j = np.arange(H, dtype=np.uint32)[None, :]
    p = (rows >> 2).astype(np.uint32)[:, None]
    ctr = np.stack(np.broadcast_arrays(j, p, np.uint32(step), np.uint32(STREAM_TAG)), axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    key = np.broadcast_to(key, ctr.shape[:-1] + (2,))
    return philox4x32_10(ctr, key)          # [R, H, 4]; lane (row & 3) picks the normal


def normals_reference(seed, step, rows, H):
    """float64 reference of the engine's Box-Muller map (see csrc/snsde_rng.cuh):
    words (x0,x1) -> n0,n1 ; (x2,x3) -> n2,n3 with
    u = (x + 0.5) * 2^-32, r = sqrt(-2 ln u_a), n_even = r cos(2 pi u_b), n_odd = r sin(2 pi u_b)."""
    bits = increment_bits(seed, step, rows, H).astype(np.float64)
    u = (bits + 0.5) * 2.0 ** -32
    r01 = np.sqrt(-2.0 * np.log(u[..., 0]))
    r23 = np.sqrt(-2.0 * np.log(u[..., 2]))
    n = np.stack([r01 * np.cos(2 * np.pi * u[..., 1]), r01 * np.sin(2 * np.pi * u[..., 1]),
                  r23 * np.cos(2 * np.pi * u[..., 3]), r23 * np.sin(2 * np.pi * u[..., 3])], axis=-1)
    lane = (np.asarray(rows, dtype=np.int64) & 3)[:, None, None]
    return np.take_along_axis(n, np.broadcast_to(lane, n.shape[:-1] + (1,)), axis=-1)[..., 0]

"""Batch sharding across the GPUs of one box.

Rows never interact inside the solve, so rank r integrates rows
``[r*B/W, (r+1)*B/W)`` with no data-path collective; the Philox stream is keyed by the GLOBAL
row (``row_offset``), so results are bit-identical for any world size.  The one exchange step
is the all-gather of the final latents: each rank's kernel writes straight into its slice of
the gather buffer and ``all_gather_into_tensor`` runs in place (NCCL over NVLink/NVSwitch;
gloo on CPU for the tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_rows, rank, world_size):
    """Contiguous, balanced split; the first ``n_rows % world_size`` ranks get one extra row."""
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_buffer(n_rows_global, tail_shape, device, dtype=torch.float32):
    return torch.empty((n_rows_global, *tail_shape), device=device, dtype=dtype)


def all_gather_rows(buf, rank=None, world_size=None, group=None):
    """In-place all-gather of ``buf`` ([B_global, ...]) whose local slice is already filled."""
    world_size = dist.get_world_size(group) if world_size is None else world_size
    rank = dist.get_rank(group) if rank is None else rank
    if world_size == 1:
        return buf
    n = buf.shape[0]
    if n % world_size == 0:
        per = n // world_size
        dist.all_gather_into_tensor(buf, buf[rank * per:(rank + 1) * per], group=group)
    else:                                   # ragged split: pad every shard to the largest one
        per = -(-n // world_size)
        tmp = buf.new_empty((world_size * per, *buf.shape[1:]))
        lo, hi = shard_bounds(n, rank, world_size)
        mine = buf.new_zeros((per, *buf.shape[1:]))
        mine[:hi - lo] = buf[lo:hi]
        dist.all_gather_into_tensor(tmp, mine, group=group)
        for r in range(world_size):
            lo, hi = shard_bounds(n, r, world_size)
            buf[lo:hi] = tmp[r * per:r * per + hi - lo]
    return buf


def solve_final_sharded(solve_fn, n_rows_global, hidden, device, rank=None, world_size=None, group=None):
    """Run ``solve_fn(lo, hi, out_slice)`` on this rank's rows and all-gather the ``[B, H]`` latents.

    ``solve_fn`` must integrate global rows ``lo:hi`` with ``row_offset=lo`` and write the result
    into ``out_slice`` (e.g. ``engine.solve_final(..., row_offset=lo, out=out_slice)``).
    """
    world_size = dist.get_world_size(group) if world_size is None else world_size
    rank = dist.get_rank(group) if rank is None else rank
    buf = gather_buffer(n_rows_global, (hidden,), device)
    lo, hi = shard_bounds(n_rows_global, rank, world_size)
    solve_fn(lo, hi, buf[lo:hi])
    return all_gather_rows(buf, rank, world_size, group)

"""Multi-GPU split of the hot path for the one-process-per-GPU launch (torchrun): every (pixel, stratum)
sample is independent, so the flat stratum range [0, spp) is cut into `world` contiguous slices, each rank renders all
pixels for its slice into its own accumulation buffer (int64 fixed-point sums, include/rtb200.h RTB_ACCUM_SCALE), and
ONE sum-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) delivers the image to rank 0.  Integer sums are
exact, so the reduced image does not depend on the split.  There is no other data-path collective.
(The same job inside ONE process -- one host thread per GPU -- is rtb_render_multi of the C ABI.)"""
from __future__ import annotations


def split_samples(spp: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced slice of the stratum index range for `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    base, extra = divmod(spp, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pass_rows(step: int, world: int, rank: int, sqrt_spp: int, rows: int = 1) -> tuple[int, int]:
    """bench.py's progressive passes: pass `step` of rank `rank` is a block of `rows` consecutive rows s_j
    of the sqrt x sqrt stratum grid (rows * sqrt_spp strata); blocks are dealt round-robin over ranks and
    wrap around the grid."""
    rows = max(1, min(rows, sqrt_spp))
    block = (step * world + rank) % (sqrt_spp // rows)
    return block * rows * sqrt_spp, (block + 1) * rows * sqrt_spp


def strong_pass(step: int, world: int, rank: int, sqrt_spp: int, rows: int = 1) -> tuple[int, int]:
    """strong scaling: pass `step` is ONE block of `rows` rows for the whole job (pass_rows with world = 1), cut into
    `world` contiguous slices; returns this rank's slice"""
    lo, hi = pass_rows(step, 1, 0, sqrt_spp, rows)
    a, b = split_samples(hi - lo, world, rank)
    return lo + a, lo + b


def to_fixed(sums):
    """f64 radiance sums -> the int64 fixed-point units of the accumulation buffers (numpy or torch)"""
    import numpy as np
    return np.rint(np.asarray(sums, dtype=np.float64) * 4294967296.0).astype(np.int64)


def reduce_to_root(accum, root: int = 0):
    """Sum-reduce a per-rank accumulation tensor onto `root` with torch.distributed (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=root, op=dist.ReduceOp.SUM)
    return accum

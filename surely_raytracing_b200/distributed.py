"""Multi-GPU split of the hot path: every (pixel, stratum) sample is independent, so the
flat stratum range [0, spp) is cut into `world` contiguous slices, each rank renders all pixels for
its slice into its own fp32 accumulation buffer, and ONE sum-reduce (NCCL over NVLink on the GPU box,
gloo in the CPU tests) delivers the image to rank 0.  There is no other data-path collective."""
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


def reduce_to_root(accum, root: int = 0):
    """Sum-reduce a per-rank accumulation tensor onto `root` with torch.distributed (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=root, op=dist.ReduceOp.SUM)
    return accum

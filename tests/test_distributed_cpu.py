"""world_size-2 gloo test of the multi-GPU plumbing on CPU: sample split + one sum-reduce.
The renderer here is the oracle in KEYED mode (the CUDA library needs a GPU); what is under test is
the partition / reduce logic of surely_raytracing_b200.distributed that bench.py uses with NCCL."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from surely_raytracing_b200.distributed import pass_rows, split_samples

ROOT = Path(__file__).resolve().parent.parent


def test_split_is_a_partition():
    for spp in (1, 49, 961, 1936, 10000):
        for world in (1, 2, 3, 4, 8):
            parts = [split_samples(spp, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == spp
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        split_samples(10, 2, 2)


def test_pass_rows_cover_the_grid_once_per_cycle():
    for world in (1, 2, 4, 8):
        seen = set()
        steps = 100 // world if 100 % world == 0 else None
        if steps is None:
            continue
        for k in range(steps):
            for r in range(world):
                lo, hi = pass_rows(k, world, r, 100)
                assert hi - lo == 100 and lo % 100 == 0
                seen.add(lo)
        assert len(seen) == 100


def test_pass_rows_in_blocks_of_rows():
    """bench.py's default step is a block of 10 rows (one rtb_render call of 1000 strata): blocks tile the
    stratum range [0, sqrt^2), never straddle it, and the ranks of one step never share a block."""
    for world in (1, 2, 4, 8):
        for rows in (1, 4, 10, 25, 100):
            n_blocks = 100 // rows
            per_step = []
            for k in range(n_blocks * 2):
                blocks = [pass_rows(k, world, r, 100, rows) for r in range(world)]
                for lo, hi in blocks:
                    assert hi - lo == rows * 100 and lo % (rows * 100) == 0 and 0 <= lo and hi <= 100 * 100
                if world <= n_blocks:
                    assert len({lo for lo, _ in blocks}) == world
                per_step.append(blocks)
            seen = {lo for blocks in per_step for lo, _ in blocks}
            assert len(seen) == n_blocks            # a cycle visits every block
    assert pass_rows(3, 1, 0, 31, 1000) == (0, 31 * 31)   # more rows than the grid has: the whole grid


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from oracle import orc
    from surely_raytracing_b200.distributed import reduce_to_root, split_samples
    from surely_raytracing_b200.scenes import BuiltScene
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = orc.OracleScene(BuiltScene("c2", width=40, spp=25))
    lo, hi = split_samples(o.info.spp_used, world, rank)
    s, _ = o.render(lo, hi, sampler=orc.SAMPLER_KEYED, threads=1)
    from surely_raytracing_b200.distributed import to_fixed
    accum = torch.from_numpy(to_fixed(s))               # int64 fixed-point accumulation buffers, like the GPU path
    np.save(out_path + f".part{rank}.npy", accum.numpy())
    reduce_to_root(accum, 0)
    if rank == 0:
        np.save(out_path, accum.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_split_and_reduce_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import orc
    from surely_raytracing_b200.scenes import BuiltScene
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "reduced.npy"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    assert got.dtype == np.int64
    parts = [np.load(str(out) + f".part{r}.npy") for r in range(2)]
    assert np.array_equal(got, parts[0] + parts[1])       # integer sums: the reduce is exact, whatever the split
    o = orc.OracleScene(BuiltScene("c2", width=40, spp=25))
    full, _ = o.render(sampler=orc.SAMPLER_KEYED, threads=1)
    assert np.allclose(got / 4294967296.0, full, rtol=1e-12, atol=2 ** -31)


def test_strong_scaling_slices_tile_each_pass():
    """bench.py --scaling strong: the ranks' slices of a pass are contiguous, disjoint and cover the pass's rows."""
    from surely_raytracing_b200.distributed import strong_pass
    for world in (1, 2, 3, 8):
        for rows in (1, 10):
            for k in range(12):
                lo, hi = pass_rows(k, 1, 0, 100, rows)
                cuts = [strong_pass(k, world, r, 100, rows) for r in range(world)]
                assert cuts[0][0] == lo and cuts[-1][1] == hi
                assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
                sizes = [b - a for a, b in cuts]
                assert max(sizes) - min(sizes) <= 1

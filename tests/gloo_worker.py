"""CPU model of the multi-GPU path for tests/test_slabs_gloo.py: world_size ranks over the gloo backend.

Each rank owns the row slab lbm-graphcore_b200/slabs.py assigns it and keeps everything else POISONED with
NaN, so the run only comes out right if the halo plan is complete: per step it sweeps its own rows with
the oracle's fused timestep and then sends exactly what the CUDA sweep kernel stores into peer memory --
planes (2,5,6) of its top row to the rank above, planes (4,7,8) of its bottom row to the rank below
(ring).  Rank 0 gathers the slabs and compares with the single-domain oracle, bit for bit.

    RANK=r WORLD_SIZE=n MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tests/gloo_worker.py <nx> <ny> <iters> <seed>
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_by_path(name, *parts):
    """Load a pure-Python module of the product package without importing the package itself (its __init__
    needs the CUDA library, which is irrelevant to this host-logic test)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, *parts))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    nx, ny, iters, seed = (int(v) for v in sys.argv[1:5])
    import torch
    import torch.distributed as dist
    from oracle import oracle

    slabs = load_by_path("lbm_slabs", "lbm-graphcore_b200", "slabs.py")
    problem = load_by_path("lbm_problem", "lbm-graphcore_b200", "problem.py")
    dist.init_process_group(backend="gloo", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()

    rng = np.random.default_rng(seed)
    p = problem.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    obst = (rng.random((ny, nx)) < 0.08).astype(np.uint8)
    full = (problem.initial_cells(p) * (1.0 + 0.1 * rng.random((ny, nx, 9))).astype(np.float32)).astype(np.float32)

    r0, n = slabs.slab_rows(rank, world, ny)
    south, north = slabs.ring_neighbours(rank, world)
    top, bottom = r0 + n - 1, r0
    ghost_s, ghost_n = (r0 - 1) % ny, (r0 + n) % ny
    N, S = list(slabs.NORTH_PLANES), list(slabs.SOUTH_PLANES)

    a = np.full((ny, nx, 9), np.nan, dtype=np.float32)
    b = np.full((ny, nx, 9), np.nan, dtype=np.float32)
    a[r0:r0 + n] = full[r0:r0 + n]

    def exchange(buf):
        """top row's north-going planes -> north neighbour's south ghost row; bottom row's south-going planes
        -> south neighbour's north ghost row."""
        up = torch.from_numpy(np.ascontiguousarray(buf[top][:, N]))
        down = torch.from_numpy(np.ascontiguousarray(buf[bottom][:, S]))
        from_s, from_n = torch.empty_like(up), torch.empty_like(down)
        reqs = [dist.isend(up, north, tag=0), dist.isend(down, south, tag=1),
                dist.irecv(from_s, south, tag=0), dist.irecv(from_n, north, tag=1)]
        for q in reqs:
            q.wait()
        # with a single row per rank the two ghost rows may be the same global row: the planes are disjoint
        buf[ghost_s][:, N] = from_s.numpy()
        buf[ghost_n][:, S] = from_n.numpy()

    # one-off guarded accelerate on the owner of row ny-2 (NaN rows fail the guard and stay untouched)
    if ny >= 2 and slabs.owner_of_row(ny - 2, world, ny) == rank:
        oracle.accelerate_flow(op, a, obst)
    exchange(a)
    tot = np.zeros(iters, dtype=np.float64)
    for t in range(iters):
        tot[t] = oracle.timestep_rows(op, a, b, obst, r0, r0 + n, oracle.ACCUM_DOUBLE, 1)
        exchange(b)
        a, b = b, a
        b[...] = np.nan  # stale data must never be needed

    free = int((obst[r0:r0 + n] == 0).sum())
    av = slabs.combine_av_vels(dist, tot, free)
    parts = [None] * world
    dist.all_gather_object(parts, (r0, a[r0:r0 + n].copy()))
    if rank == 0:
        got = np.concatenate([c for _, c in sorted(parts, key=lambda rc: rc[0])], axis=0)
        want, want_av = oracle.run(op, full, obst, iters, accum=oracle.ACCUM_DOUBLE)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "slab model differs from the single-domain oracle"
        ulp = np.abs(av.view(np.int32).astype(np.int64) - want_av.view(np.int32).astype(np.int64)).max()
        assert ulp <= 1, f"combined av_vels off by {ulp} ulp"
        print(f"gloo slab model ok: {world} ranks, {nx}x{ny}, {iters} steps, av_vels max {ulp} ulp")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

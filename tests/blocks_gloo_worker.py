"""CPU model of the 2-D block partition for tests/test_blocks_gloo.py: px * py ranks over gloo.

Each rank owns the block lbm-graphcore_b200/blocks.py assigns it and keeps everything else POISONED with NaN, so its
cells only come out right if the halo plan is complete: per step it sweeps the rows of its block with the oracle's fused
timestep (columns outside the block compute garbage that is thrown away) and then sends exactly what the CUDA sweep kernel
stores into its neighbours -- three planes of each edge row / column, one population of each corner cell.  Rank 0 gathers the
blocks and compares with the single-domain oracle, bit for bit.

    RANK=r WORLD_SIZE=n MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tests/blocks_gloo_worker.py <nx> <ny> <px> <py> <iters> <seed>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gloo_worker import load_by_path  # noqa: E402


def main():
    nx, ny, px, py, iters, seed = (int(v) for v in sys.argv[1:7])
    import torch
    import torch.distributed as dist
    from oracle import oracle

    blocks = load_by_path("lbm_blocks", "lbm-graphcore_b200", "blocks.py")
    problem = load_by_path("lbm_problem", "lbm-graphcore_b200", "problem.py")
    dist.init_process_group(backend="gloo", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == px * py

    p, full, obst = problem.perturbed_problem(nx, ny, iters, seed, solid_frac=0.08)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    me = blocks.block_of(rank, px, py, nx, ny)
    r0, n, c0, m = me["row0"], me["nyl"], me["col0"], me["nxl"]
    rows, cols = slice(r0, r0 + n), slice(c0, c0 + m)
    top, bottom, east, west = r0 + n - 1, r0, c0 + m - 1, c0

    a = np.full((ny, nx, 9), np.nan, dtype=np.float32)
    b = np.full((ny, nx, 9), np.nan, dtype=np.float32)
    a[rows, cols] = full[rows, cols]

    def exchange(buf):
        """What the sweep kernel pushes, as (destination rank, global cells, planes, values); then the matching receives."""
        sends = []
        for (dx, dy), planes, src_rows, src_cols in (
                ((0, 1), blocks.NORTH_PLANES, [top], range(c0, c0 + m)), ((0, -1), blocks.SOUTH_PLANES, [bottom], range(c0, c0 + m)),
                ((1, 0), blocks.EAST_PLANES, range(r0, r0 + n), [east]), ((-1, 0), blocks.WEST_PLANES, range(r0, r0 + n), [west]),
                ((1, 1), (blocks.CORNER_PLANE[(1, 1)],), [top], [east]), ((-1, 1), (blocks.CORNER_PLANE[(-1, 1)],), [top], [west]),
                ((-1, -1), (blocks.CORNER_PLANE[(-1, -1)],), [bottom], [west]), ((1, -1), (blocks.CORNER_PLANE[(1, -1)],), [bottom], [east])):
            ys, xs = np.array(list(src_rows)), np.array(list(src_cols))
            vals = buf[np.ix_(ys, xs, list(planes))]
            sends.append((blocks.neighbour(rank, px, py, dx, dy), (dx, dy), ys, xs, planes, np.ascontiguousarray(vals)))
        # tags by direction; a rank may be its own neighbour (px or py == 1): handled locally
        reqs, recvs = [], []
        for k, (dst, d, ys, xs, planes, vals) in enumerate(sends):
            if dst != rank:
                reqs.append(dist.isend(torch.from_numpy(vals), dst, tag=k))
        for k, (dst, (dx, dy), ys, xs, planes, vals) in enumerate(sends):
            src = blocks.neighbour(rank, px, py, -dx, -dy)  # who sends me its direction-k data
            if src == rank:
                got = vals  # the block is its own neighbour in that direction: what it sent is what it gets
                other = me
            else:
                other = blocks.block_of(src, px, py, nx, ny)
                shape = (len(ys) if dx != 0 and dy == 0 else 1, len(xs) if dy != 0 and dx == 0 else 1, len(planes))
                if dx != 0 and dy == 0:
                    shape = (other["nyl"], 1, len(planes))
                elif dy != 0 and dx == 0:
                    shape = (1, other["nxl"], len(planes))
                t = torch.empty(shape, dtype=torch.float32)
                reqs.append(dist.irecv(t, src, tag=k))
                got = t
            recvs.append(((dx, dy), other, planes, got))
        for q in reqs:
            q.wait()
        for (dx, dy), o, planes, got in recvs:
            got = got.numpy() if hasattr(got, "numpy") else got
            # the sender's edge cells, in global coordinates: they are my ghost cells
            ys = [o["row0"] + o["nyl"] - 1] if dy == 1 else ([o["row0"]] if dy == -1 else list(range(o["row0"], o["row0"] + o["nyl"])))
            xs = [o["col0"] + o["nxl"] - 1] if dx == 1 else ([o["col0"]] if dx == -1 else list(range(o["col0"], o["col0"] + o["nxl"])))
            buf[np.ix_(np.array(ys), np.array(xs), list(planes))] = got

    # one-off guarded accelerate on the owners of row ny-2 (NaN cells fail the guard and stay untouched)
    if ny >= 2 and r0 <= ny - 2 < r0 + n:
        oracle.accelerate_flow(op, a, obst)
    exchange(a)
    for t in range(iters):
        oracle.timestep_rows(op, a, b, obst, r0, r0 + n, oracle.ACCUM_DOUBLE, 1)
        keep = b[rows, cols].copy()
        b[...] = np.nan  # what was computed outside the block is garbage; stale data must never be needed
        b[rows, cols] = keep
        exchange(b)
        a, b = b, a

    parts = [None] * world
    dist.all_gather_object(parts, (r0, c0, a[rows, cols].copy()))
    if rank == 0:
        got = np.empty((ny, nx, 9), dtype=np.float32)
        for br0, bc0, blk in parts:
            got[br0:br0 + blk.shape[0], bc0:bc0 + blk.shape[1]] = blk
        want, _ = oracle.run(op, full, obst, iters, accum=oracle.ACCUM_DOUBLE)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "block model differs from the single-domain oracle"
        print(f"gloo block model ok: {px} x {py} ranks, {nx}x{ny}, {iters} steps")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

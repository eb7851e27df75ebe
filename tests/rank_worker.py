"""Worker for tests/test_multi_gpu.py: one process per GPU under torchrun (the launch bench.py gets).

Each rank owns one row slab, maps its ring neighbours' lattices with CUDA IPC, uploads ITS rows only, runs
the time loop (the sweep kernel stores its halo rows straight into the neighbour GPU's memory) and writes
its rows + per-step tot_u partials to <outdir>/rank<r>.npz for the parent test to compare with the oracle.

    torchrun --nproc-per-node N tests/rank_worker.py <outdir> <nx> <ny> <iters> <seed> [<pieces>]
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problem(lbm, nx, ny, iters, seed):
    """Same generator as tests/test_gpu_parity.py::_random_problem (seeded, identical on every rank)."""
    return lbm.perturbed_problem(nx, ny, iters, seed)


def main():
    outdir, nx, ny, iters, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    pieces = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    grid = tuple(int(v) for v in sys.argv[7].split("x")) if len(sys.argv) > 7 else None  # PXxPY: one BLOCK per rank
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="gloo", init_method="env://")
    lbm = importlib.import_module("lbm-graphcore_b200")
    from importlib import import_module
    slabs = import_module("lbm-graphcore_b200.slabs")

    p, cells, obst = problem(lbm, nx, ny, iters, seed)
    s = lbm.Solver(p, rank=rank, world=world, device=local, grid=grid)
    slabs.connect_ring(s, dist)  # (all-gather of the IPC blobs + lbm_ipc_connect: the same for slabs and blocks)
    r0, n, c0, m = s.row_begin, s.row_count, s.col_begin, s.col_count
    if grid is None:
        assert (r0, n) == slabs.slab_rows(rank, world, ny) and (c0, m) == (0, nx)
    else:
        assert world == grid[0] * grid[1] and s.partition == grid
    s.upload(cells[r0:r0 + n, c0:c0 + m], obst[r0:r0 + n, c0:c0 + m])
    dist.barrier()  # every rank's lattice is in place before anyone pushes halo rows
    done = 0
    for i in range(pieces):
        k = (iters - done) if i == pieces - 1 else max(1, iters // pieces)
        s.run(k)
        done += k
    got, _ = s.download(av_vels=False)
    tot = s.download_tot_u()
    av = slabs.combine_av_vels(dist, tot, s.free_cells)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), cells=got, row_begin=r0, col_begin=c0, tot_u=tot, free_cells=s.free_cells,
             av_vels=av, variant=s.kernel_variant)
    slabs.close_ring(s, dist)  # barrier -> unmap the neighbours -> barrier -> free
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

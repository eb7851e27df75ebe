"""Row-slab decomposition across ranks: the host-side plan of the multi-GPU path (pure Python/numpy).

The reference cuts the grid into blocks with periodic 1-cell halos (main/include/StructuredGridUtils.hpp:472-561,
`Halos::forSliceWithWraparound` :805-851) and copies all nine populations of every halo cell
(main/LbmAoS.cpp:151-166).  Here the grid is cut into ROW slabs (x-wrap stays local, y-wrap makes the ranks a
ring) and only the populations that actually cross an edge travel:

    top owned row    -> planes NORTH_PLANES (2 N, 5 NE, 6 NW) -> south ghost row of the rank above
    bottom owned row -> planes SOUTH_PLANES (4 S, 7 SW, 8 SE) -> north ghost row of the rank below

On the GPUs the sweep kernel itself performs those stores into peer memory (lbm_kernels.cuh, the `gn` / `gs` stores of sweep_tile_row);
this module holds what the host has to agree on -- who owns which rows, who the ring neighbours are, how the
IPC handles are exchanged and in which order the per-rank av_vels partials are combined -- so that bench.py,
the tests and the gloo CPU model all use one definition.  `dist` is `torch.distributed` (any backend that
can all_gather_object: gloo on CPU, nccl+gloo on the GPU box); nothing here imports torch.
"""
from __future__ import annotations

import numpy as np

NORTH_PLANES = (2, 5, 6)  # populations moving to y+1: leave through a slab's top row
SOUTH_PLANES = (4, 7, 8)  # populations moving to y-1: leave through a slab's bottom row


def slab_rows(rank: int, world: int, ny: int) -> tuple[int, int]:
    """(first global row, row count) of `rank`; identical to lbm_create_rank (lbm_capi.cu)."""
    if not (0 <= rank < world) or ny < world:
        raise ValueError(f"cannot cut {ny} rows into {world} slabs (rank {rank})")
    r0 = rank * ny // world
    return r0, (rank + 1) * ny // world - r0


def ring_neighbours(rank: int, world: int) -> tuple[int, int]:
    """(south, north) neighbour ranks; the y-wrap closes the ring (rank 0's south is world-1)."""
    return (rank + world - 1) % world, (rank + 1) % world


def owner_of_row(row: int, world: int, ny: int) -> int:
    """Rank that owns global row `row` (the driven row ny-2 lives on exactly one rank)."""
    for r in range(world):
        r0, n = slab_rows(r, world, ny)
        if r0 <= row < r0 + n:
            return r
    raise ValueError(f"row {row} outside 0..{ny - 1}")


def connect_ring(solver, dist) -> None:
    """Exchange the CUDA IPC blobs of all ranks and map the neighbours' lattices (lbm_ipc_export ->
    all-gather -> lbm_ipc_connect).  Collective: every rank must call it."""
    world = dist.get_world_size()
    blobs = [None] * world
    dist.all_gather_object(blobs, solver.ipc_export())
    solver.ipc_connect(b"".join(blobs))
    dist.barrier()


def close_ring(solver, dist) -> None:
    """Two-phase teardown of a connected rank-mode solver (protocol in include/lbm_b200.h): all ranks have finished
    their work -> every rank unmaps its neighbours' slabs -> barrier -> free.  Collective: every rank must call it."""
    if dist is None or dist.get_world_size() == 1:
        solver.close()
        return
    dist.barrier()           # nobody is still running or pushing halos
    solver.ipc_disconnect()  # waits for this rank's stream, closes its peer mappings
    dist.barrier()           # every mapping of every slab is closed
    solver.close()


def barrier_before_reupload(dist) -> None:
    """A rank may reset its halo flags (lbm_upload / lbm_init_synthetic) only when no neighbour is still running."""
    if dist is not None and dist.get_world_size() > 1:
        dist.barrier()


def combine_tot_u(dist, tot_u: np.ndarray, free_cells: int) -> tuple[np.ndarray, int]:
    """All ranks' per-step sums of sqrt(u_sq) added in RANK ORDER (the documented, deterministic order of
    the last reduction stage) and the global free-cell count."""
    if dist is None:
        return np.asarray(tot_u, dtype=np.float64), int(free_cells)
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(tot_u, dtype=np.float64), int(free_cells)))
    total = np.zeros_like(parts[0][0])
    for t, _ in parts:  # rank order
        total = total + t
    return total, sum(f for _, f in parts)


def av_vels_from_tot_u(tot_u: np.ndarray, free_cells: int) -> np.ndarray:
    """main/LastChance.cpp:266: the fp32 quotient of the step total by the fp32 free-cell count."""
    return (np.asarray(tot_u, dtype=np.float64).astype(np.float32) / np.float32(free_cells)).astype(np.float32)


def combine_av_vels(dist, tot_u: np.ndarray, free_cells: int) -> np.ndarray:
    total, free_total = combine_tot_u(dist, tot_u, free_cells)
    return av_vels_from_tot_u(total, free_total)

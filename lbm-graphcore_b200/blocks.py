"""2-D block decomposition: the host-side plan of lbm_create_grid (pure Python/numpy, no CUDA).

The reference cuts its grid for several devices into a grid of blocks (`partitionForIpus`,
main/include/StructuredGridUtils.hpp:501-527: 4 -> 2 x 2, 8 -> 4 x 2 or 2 x 4) and gives every block a one-cell halo
with wrap-around (`Halos::forSliceWithWraparound`, :805-851), copying all nine populations of every halo cell.  Here only
the populations that cross an edge travel, pushed by the block that computed them:

    top row    -> NORTH_PLANES (2 N, 5 NE, 6 NW) -> south ghost row of the block above      bottom row -> SOUTH_PLANES -> below
    east column -> EAST_PLANES (1 E, 5 NE, 8 SE) -> west ghost column of the block beside   west column -> WEST_PLANES
    the four corner cells -> one diagonal population each -> the corner ghost of the diagonal block

`lbm_partition.hpp` holds the same arithmetic for the C++ side; tests/test_blocks_gloo.py runs this plan over gloo with
every rank holding nothing but its block and checks the result against the single-domain oracle.
"""
from __future__ import annotations

NORTH_PLANES, SOUTH_PLANES = (2, 5, 6), (4, 7, 8)
EAST_PLANES, WEST_PLANES = (1, 5, 8), (3, 6, 7)
# diagonal neighbour (dx, dy) -> the one population that moves that way
CORNER_PLANE = {(1, 1): 5, (-1, 1): 6, (-1, -1): 7, (1, -1): 8}


def range_of(n: int, parts: int, i: int) -> tuple[int, int]:
    """(begin, count) of part i of n cells cut into `parts` (lbm::partition::range_of)."""
    b = i * n // parts
    return b, (i + 1) * n // parts - b


def grid_for_gpus(nx: int, ny: int, gpus: int) -> tuple[int, int]:
    """(px, py) the reference would choose (partitionForIpus, StructuredGridUtils.hpp:479-527)."""
    row_imb, col_imb = (ny % gpus) / ny, (nx % gpus) / nx
    px, py = 1, gpus
    if gpus == 2:
        px, py = (1, 2) if row_imb < col_imb else (2, 1)
    elif gpus == 4:
        px, py = 2, 2
    elif gpus == 8:
        px, py = (2, 4) if row_imb < col_imb else (4, 2)
    elif gpus == 16:
        px, py = 4, 4
    return min(px, nx), min(py, ny)


def block_of(index: int, px: int, py: int, nx: int, ny: int) -> dict:
    """Block `index` = by * px + bx: its position, columns and rows."""
    bx, by = index % px, index // px
    c0, m = range_of(nx, px, bx)
    r0, n = range_of(ny, py, by)
    return {"bx": bx, "by": by, "col0": c0, "nxl": m, "row0": r0, "nyl": n}


def neighbour(index: int, px: int, py: int, dx: int, dy: int) -> int:
    """Index of the block dx to the east and dy to the north (both directions periodic)."""
    bx, by = index % px, index // px
    return ((by + dy) % py) * px + (bx + dx) % px

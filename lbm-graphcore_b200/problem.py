"""Problem description for the D2Q9-BGK solver: pure numpy, no CUDA, no shared library.

Mirrors lbm::Params / lbm::Obstacles / lbm::Cells of the reference
(main/include/LbmParams.hpp, main/include/LatticeBoltzmannUtils.hpp of thorbenlouw/lbm-graphcore)
plus the synthetic-channel generator of SURVEY.md 8(d).  bench.py's reference arm loads this file on
its own (by path) so that nothing of the CUDA product is on that path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

NSPEEDS = 9

@dataclass(frozen=True)
class Params:
    """lbm::Params (main/include/LbmParams.hpp:16-65)."""

    nx: int
    ny: int
    max_iters: int
    reynolds_dim: int
    density: float
    accel: float
    omega: float

    @staticmethod
    def from_file(path: str) -> "Params":
        """Params::fromFile (LbmParams.hpp:28-58): seven lines, four unsigned then three floats."""
        with open(path, "r") as f:
            lines = [f.readline() for _ in range(7)]
        if any(line == "" for line in lines):
            raise ValueError(f"Could not read parameters from {path}")
        ints = [int(lines[i].split()[0]) for i in range(4)]
        floats = [float(np.float32(lines[i].split()[0])) for i in range(4, 7)]
        if any(v < 0 for v in ints):
            raise ValueError(f"negative size in {path}")
        return Params(ints[0], ints[1], ints[2], ints[3], *floats)

    def reynolds_number(self, average_velocity: float) -> float:
        """lbm::reynoldsNumber (LatticeBoltzmannUtils.hpp:202-205), fp32 arithmetic."""
        f = np.float32
        viscosity = f(1.0) / f(6.0) * (f(2.0) / f(self.omega) - f(1.0))
        return float(f(average_velocity) * f(self.reynolds_dim) / viscosity)


def load_obstacles(nx: int, ny: int, path: str) -> np.ndarray:
    """Obstacles::fromFile (LbmParams.hpp:92-123): lines "x y 1"; returns uint8[ny, nx].

    Like main/LastChance.cpp:476-480 (and unlike the C++ loader) coordinates are range-checked;
    duplicate lines, which the shipped files contain, are idempotent."""
    obst = np.zeros((ny, nx), dtype=np.uint8)
    with open(path, "r") as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            if len(parts) != 3:
                raise ValueError("expected 3 values per line in obstacle file")
            x, y, blocked = int(parts[0]), int(parts[1]), int(parts[2])
            if blocked != 1:
                raise ValueError("Malformed line: obstacle must be 1")
            if not (0 <= x < nx and 0 <= y < ny):
                raise ValueError("obstacle coordinate out of range")
            obst[y, x] = 1
    return obst


def initial_cells(p: Params) -> np.ndarray:
    """Cells::initialise (LatticeBoltzmannUtils.hpp:137-157): float32[ny, nx, 9]."""
    f = np.float32
    w0 = f(p.density) * f(4.0) / f(9.0)
    w1 = f(p.density) / f(9.0)
    w2 = f(p.density) / f(36.0)
    cells = np.empty((p.ny, p.nx, NSPEEDS), dtype=np.float32)
    cells[..., 0] = w0
    cells[..., 1:5] = w1
    cells[..., 5:9] = w2
    return cells


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_obstacles(nx: int, ny: int, seed: int, solid_prob: float, row_begin: int = 0,
                        row_count: int | None = None) -> np.ndarray:
    """Host twin of lbm_init_synthetic's mask: wall rows 0 and ny-1, Bernoulli(solid_prob)
    elsewhere except the driven row ny-2, from the top 24 bits of splitmix64(seed + y*nx + x)."""
    if row_count is None:
        row_count = ny - row_begin
    ys = np.arange(row_begin, row_begin + row_count, dtype=np.uint64)[:, None]
    xs = np.arange(nx, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) + ys * np.uint64(nx) + xs)
    threshold = np.uint64(int(float(np.float32(solid_prob)) * 16777216.0))
    solid = (h >> np.uint64(40)) < threshold
    yy = np.broadcast_to(ys, solid.shape)
    solid = np.where(yy == np.uint64(ny - 2), False, solid)
    solid = np.where((yy == np.uint64(0)) | (yy == np.uint64(ny - 1)), True, solid)
    return solid.astype(np.uint8)




def perturbed_problem(nx: int, ny: int, iters: int, seed: int, solid_frac: float = 0.05):
    """A seeded random test problem whose every halo row carries information from step 0 on: 5 % random solids and
    every population perturbed by up to 10 % (shipped parameters otherwise).  Identical on every rank for one seed.
    Used by the parity tests (tests/rank_worker.py) and by bench.py's driver-visible multi-GPU parity record."""
    rng = np.random.default_rng(seed)
    p = Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    obst = (rng.random((ny, nx)) < solid_frac).astype(np.uint8)
    cells = (initial_cells(p) * (1.0 + 0.1 * rng.random((ny, nx, 9))).astype(np.float32)).astype(np.float32)
    return p, cells, obst

"""CPU model of lbm_pipeline's banding (lbm_capi.cu, solve_band): K timesteps of a band of rows depend only on the band plus
K rows either side of the INITIAL state.  Each band is solved here exactly the way a band solver does it -- as its own small
periodic problem of band + 2 x margin rows, with the driven row wherever global row ny-2 falls in it (or nowhere) -- with the
CPU oracle; the bands' central rows must reproduce the single-domain oracle bit for bit, and their per-step sums its av_vels.
The GPU twin is tests/test_solve_host_gpu.py."""
import ctypes

import numpy as np
import pytest

from conftest import ulp_distance
from gloo_worker import load_by_path

problem = load_by_path("lbm_problem_pm", "lbm-graphcore_b200", "problem.py")


def margin_for(iters):
    """(M + 1) a multiple of the 8-row tile, M >= iters: lbm_capi.cu plan_bands."""
    return (max(iters, 1) + 1 + 7) // 8 * 8 - 1


def solve_band(oracle, p, cells, obst, r0, rows, iters, margin):
    """One band solver: rows [r0 - M, r0 + rows + M) of the grid (periodic) as a periodic problem of its own."""
    ny, nx = obst.shape
    total = rows + 2 * margin
    idx = np.arange(r0 - margin, r0 + rows + margin) % ny
    bc, bo = np.ascontiguousarray(cells[idx]), np.ascontiguousarray(obst[idx])
    first = (r0 - margin) % ny
    arow = (ny - 2 - first) % ny            # where the driven row sits in the band, if it does
    a, b = bc.copy(), np.empty_like(bc)
    tot = np.zeros(iters)
    # oracle parameters with ny = the band height put the driven row at total - 2: shift the band so that it IS there, or, when
    # the band does not contain it, run with accel = 0 (w1 = w2 = 0: both accelerates become no-ops)
    if arow < total:
        shift = (total - 2) - arow
        a, bo2 = np.roll(a, shift, axis=0), np.roll(bo, shift, axis=0)
        op = oracle.make_params(nx, total, iters, 10, p.density, p.accel, p.omega)
    else:
        shift, bo2 = 0, bo
        op = oracle.make_params(nx, total, iters, 10, p.density, 0.0, p.omega)
    b = np.empty_like(a)
    oracle.accelerate_flow(op, a, bo2)
    central = (np.arange(margin, margin + rows) + shift) % total
    for t in range(iters):
        oracle.timestep_rows(op, a, b, bo2, 0, total, oracle.ACCUM_DOUBLE, 1)
        a, b = b, a
    out = a[central]
    # per-step sums over the central rows need the pre-collision speeds of every step: recompute them step by step
    a2 = np.roll(bc.copy(), shift, axis=0)
    b2 = np.empty_like(a2)
    oracle.accelerate_flow(op, a2, bo2)
    for t in range(iters):
        lo, hi = int(central.min()), int(central.max()) + 1
        if hi - lo == rows:
            tot[t] = oracle.timestep_rows(op, a2, b2, bo2, lo, hi, oracle.ACCUM_DOUBLE, 1)
            oracle.timestep_rows(op, a2, b2, bo2, 0, lo, oracle.ACCUM_DOUBLE, 1)
            oracle.timestep_rows(op, a2, b2, bo2, hi, total, oracle.ACCUM_DOUBLE, 1)
        else:  # the shift wrapped the central rows around the band's end: two pieces
            wrapped = np.sort(central)
            cut = int(np.where(np.diff(wrapped) > 1)[0][0]) + 1
            (l0, h0), (l1, h1) = (int(wrapped[0]), int(wrapped[cut - 1]) + 1), (int(wrapped[cut]), int(wrapped[-1]) + 1)
            tot[t] = (oracle.timestep_rows(op, a2, b2, bo2, l0, h0, oracle.ACCUM_DOUBLE, 1) +
                      oracle.timestep_rows(op, a2, b2, bo2, l1, h1, oracle.ACCUM_DOUBLE, 1))
            oracle.timestep_rows(op, a2, b2, bo2, h0, l1, oracle.ACCUM_DOUBLE, 1)
        a2, b2 = b2, a2
    return out, tot, int((obst[r0:r0 + rows] == 0).sum())


@pytest.mark.parametrize("nx,ny,iters,band", [(24, 64, 5, 16), (16, 96, 7, 32), (40, 48, 3, 8), (12, 40, 9, 8)])
def test_bands_with_margins_reproduce_the_single_domain_run(oracle, nx, ny, iters, band):
    p, cells, obst = problem.perturbed_problem(nx, ny, iters, seed=nx * ny + iters, solid_frac=0.08)
    op = oracle.make_params(nx, ny, iters, 10, p.density, p.accel, p.omega)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    m = margin_for(iters)
    assert m >= iters and (m + 1) % 8 == 0 and band + 2 * m <= ny
    got = np.empty_like(want_cells)
    tot, free = np.zeros(iters), 0
    for r0 in range(0, ny, band):  # band order: the order the pipeline adds the bands' sums in
        rows = min(band, ny - r0)
        out, t, f = solve_band(oracle, p, cells, obst, r0, rows, iters, m)
        got[r0:r0 + rows] = out
        tot, free = tot + t, free + f
    assert np.array_equal(got.view(np.uint32), want_cells.view(np.uint32))
    av = (tot.astype(np.float32) / np.float32(free)).astype(np.float32)
    assert ulp_distance(av, want_av).max() <= 1


def test_a_margin_shorter_than_the_job_is_not_enough(oracle):
    """The converse: with fewer margin rows than timesteps the band's periodic wrap does reach the band."""
    nx, ny, iters = 16, 64, 9
    p, cells, obst = problem.perturbed_problem(nx, ny, iters, seed=3, solid_frac=0.0)
    op = oracle.make_params(nx, ny, iters, 10, p.density, p.accel, p.omega)
    want_cells, _ = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    out, _, _ = solve_band(oracle, p, cells, obst, 16, 16, iters, margin=4)
    assert not np.array_equal(out.view(np.uint32), want_cells[16:32].view(np.uint32))


def test_a_band_taller_than_the_grid_holds_the_driven_row_twice(oracle):
    """Why plan_bands never lets band + 2 x margin exceed ny: such a band contains two images of global row ny - 2 and a band
    solver drives one.  (lbm_pipeline_plan shrinks the band instead: tests/test_abi.py.)"""
    nx, ny, iters = 12, 40, 9
    p, cells, obst = problem.perturbed_problem(nx, ny, iters, seed=11, solid_frac=0.0)
    op = oracle.make_params(nx, ny, iters, 10, p.density, p.accel, p.omega)
    want_cells, _ = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    out, _, _ = solve_band(oracle, p, cells, obst, 0, ny, iters, margin_for(iters))
    assert not np.array_equal(out.view(np.uint32), want_cells.view(np.uint32))

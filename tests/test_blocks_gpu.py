"""2-D block partition with ghost columns and corner exchange (the reference's partitionForIpus /
Halos, main/include/StructuredGridUtils.hpp:501-527,805-851): whatever the block grid, the lattice must be bit-identical to the
single-domain CPU oracle.  Here all blocks live on ONE GPU (they run back to back on one stream and exchange their
halos through the same stores and epoch flags as blocks on different GPUs); tests/test_multi_gpu.py repeats it across GPUs."""
import numpy as np
import pytest

from conftest import ulp_distance

pytestmark = pytest.mark.gpu


def _check(oracle, lbm, nx, ny, px, py, iters=40, pieces=(), seed=0, devices=None):
    p, cells, obst = lbm.perturbed_problem(nx, ny, iters, seed or (nx * 131 + ny * 7 + px * 3 + py))
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p, grid=(px, py), devices=devices or [0] * (px * py)) as s:
        assert s.partition == (px, py) and len(s.blocks) == px * py
        assert sum(b["row_count"] * b["col_count"] for b in s.blocks) == nx * ny
        s.upload(cells, obst)
        assert s.free_cells == int((obst == 0).sum())
        done = 0
        for k in pieces:
            s.run(k)
            done += k
        s.run(iters - done)
        got_cells, got_av = s.download()
        fields = s.download_fields()
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1
    ux, uy, u, pressure = oracle.final_state_fields(op, want_cells, obst)
    assert np.array_equal(fields[..., 3].view(np.uint32), pressure.view(np.uint32))


@pytest.mark.parametrize("px,py", [(1, 1), (2, 1), (1, 2), (2, 2), (4, 2), (2, 4), (3, 3), (8, 1), (1, 8)])
def test_block_grids_bit_exact(oracle, lbm, px, py):
    """Block widths that are multiples of the 128-cell tile: the float4 kernel with ghost columns."""
    _check(oracle, lbm, 128 * px * 2, 24 * py, px, py)


@pytest.mark.parametrize("nx,ny,px,py", [(300, 50, 2, 2),      # blocks 150 wide: a full tile + a ragged one (hybrid kernel)
                                         (200, 33, 2, 3),      # blocks 100 wide: scalar kernel; uneven rows 11 / 11 / 11
                                         (257, 19, 2, 2),      # 128 | 129 columns: one block vectorised, one hybrid
                                         (130, 13, 4, 1),      # 32 / 33 / 32 / 33 columns
                                         (5, 7, 5, 7),         # one cell per block: every neighbour is a ghost, every cell a corner
                                         (6, 4, 3, 2),
                                         (512, 9, 1, 3),       # three-row blocks: the south and north edge tile rows coincide
                                         (1024, 64, 2, 2),
                                         (128, 16, 1, 1)])     # 1 x 1: the block is all eight of its own neighbours
def test_block_grids_ragged(oracle, lbm, nx, ny, px, py):
    _check(oracle, lbm, nx, ny, px, py, iters=25)


def test_block_grid_run_in_pieces_and_odd_counts(oracle, lbm):
    _check(oracle, lbm, 512, 96, 2, 4, iters=31, pieces=(1, 2, 9))


def test_block_grid_equals_row_slabs_on_the_synthetic_channel(lbm):
    """The synthetic channel generated per block (global coordinates in the hash) equals the one generated per slab."""
    p = lbm.Params(1024, 512, 30, 10, 0.1, 0.005, 1.85)
    out = []
    for kw in ({"devices": [0, 0, 0, 0]}, {"grid": (2, 2), "devices": [0, 0, 0, 0]}, {"grid": (4, 1), "devices": [0, 0, 0, 0]}):
        with lbm.Solver(p, **kw) as s:
            s.init_synthetic(1234, 0.02)
            free = s.free_cells
            s.run(30)
            cells, av = s.download()
        out.append((cells, av, free))
    for cells, av, free in out[1:]:
        assert free == out[0][2]
        assert np.array_equal(cells.view(np.uint32), out[0][0].view(np.uint32))
        assert ulp_distance(av, out[0][1]).max() <= 1


def test_block_grid_argument_errors(lbm):
    p = lbm.Params(64, 64, 1, 10, 0.1, 0.005, 1.85)
    with pytest.raises(lbm.LbmError):
        lbm.Solver(p, grid=(0, 1))
    with pytest.raises(lbm.LbmError, match="1..64"):
        lbm.Solver(p, grid=(65, 1), devices=[0] * 65)
    with pytest.raises(lbm.LbmError, match="columns"):
        lbm.Solver(lbm.Params(8, 64, 1, 10, 0.1, 0.005, 1.85), grid=(16, 1), devices=[0] * 16)
    with pytest.raises(lbm.LbmError, match="GPU"):
        lbm.Solver(p, grid=(1, 2), devices=[0, 99])


def test_lbm_runner_grid_and_stats(lbm, tmp_path):
    """lbm_runner --grid PXxPY (all blocks on GPU 0 here is not expressible through the CLI, so 1x1 with ghost columns) and
    --stats: same output files as the row-slab run, plus the partition report."""
    from conftest import obstacles_path, params_path
    import subprocess
    outs = []
    for extra in ([], ["--grid", "1x1", "--stats"]):
        d = tmp_path / ("grid" if extra else "plain")
        d.mkdir()
        r = subprocess.run([lbm.RUNNER_PATH, "--params", params_path("128x256"), "--obstacles", obstacles_path("128x256"), "--outdir", str(d),
                            "--iters", "2000", "--reruns", "0", *extra], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        if extra:
            assert "numGpus,width,height,px,py" in r.stdout and "Block 0: GPU 0, rows [0, 256), cols [0, 128)" in r.stdout
            assert "Free-cell load balance: 100.00 %" in r.stdout
        outs.append(d)
    assert (outs[0] / "final_state.dat").read_bytes() == (outs[1] / "final_state.dat").read_bytes()
    a, b = (np.loadtxt(o / "av_vels.dat", usecols=[1]).astype(np.float32) for o in outs)
    assert ulp_distance(a, b).max() <= 1

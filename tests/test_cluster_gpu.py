"""The cluster-resident kernel (sweep_cluster_kernel: both lattices in the distributed shared memory of thread-block
clusters, halo rows pushed through DSMEM inside a cluster and through a global halo row + epoch flag between clusters)
against the oracle: bit-exact lattice, av_vels within 1 fp32 ulp, whatever the cluster geometry."""
import numpy as np
import pytest

from conftest import ulp_distance

pytestmark = pytest.mark.gpu


def _check(oracle, lbm, nx, ny, iters, pieces=(), seed=0):
    p, cells, obst = lbm.perturbed_problem(nx, ny, iters, seed or (nx * 31 + ny))
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p) as s:
        assert s.kernel_variant == "vec4-cluster"
        s.upload(cells, obst)
        done = 0
        for k in pieces:
            s.run(k)
            done += k
        s.run(iters - done)
        got_cells, got_av = s.download()
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1


# CTAs G = ny / (cells / nx); the largest cluster size <= 16 dividing G is used
SHAPES = [(128, 128),   # 1024-cell CTAs: one cluster of 16, the y-wrap closes inside the cluster; 512: two clusters
          (128, 256),   # two / four clusters of 16
          (256, 256),   # four / eight clusters of 16, halo rows spread over several warps
          (256, 64),
          (128, 24),    # G = 3 / 6: one odd-sized cluster
          (256, 52),    # G = 13 / 26
          (128, 136),   # G = 17 (seventeen clusters of ONE CTA: every edge goes through the global halo) / 34
          (256, 8),
          (128, 8)]     # G = 1 / 2: a CTA that is its own neighbour on both sides
WIDE = [(512, 512),     # 2048-cell CTAs (four rows of 512, four cells per thread): 128 CTAs
        (512, 64),      # 1024-cell CTAs (two rows, two cells per thread): 32 CTAs
        (512, 296),     # 148 CTAs of 1024 cells: every SM
        (512, 12)]      # G = 6


@pytest.mark.parametrize("nx,ny", SHAPES + WIDE)
def test_cluster_kernel_bit_exact(oracle, lbm, nx, ny):
    _check(oracle, lbm, nx, ny, 45)


@pytest.mark.parametrize("cells,cpt", [("1024", "4"), ("1024", "2"), ("512", "4"), ("512", "2"), ("512", "1"), ("256", "1")])
@pytest.mark.parametrize("nx,ny", SHAPES)
def test_cluster_kernel_every_shape(oracle, lbm, monkeypatch, nx, ny, cells, cpt):
    """Every instantiation of the kernel: cells per CTA x cells per thread (float4 / float2 / scalar shared-memory access)."""
    if cells == "256" and nx != 128:
        pytest.skip("256-cell CTAs need two full rows: nx = 128 only")
    monkeypatch.setenv("LBM_B200_CLUSTER_CELLS", cells)
    monkeypatch.setenv("LBM_B200_CLUSTER_CPT", cpt)
    _check(oracle, lbm, nx, ny, 21)


@pytest.mark.parametrize("size", ["8", "4", "1"])
@pytest.mark.parametrize("nx,ny", [(128, 128), (256, 256)])
def test_cluster_kernel_with_smaller_clusters(oracle, lbm, monkeypatch, nx, ny, size):
    """LBM_B200_CLUSTER_SIZE forces smaller clusters, i.e. more inter-cluster (global halo + flag) edges."""
    monkeypatch.setenv("LBM_B200_CLUSTER_SIZE", size)
    _check(oracle, lbm, nx, ny, 33)


def test_cluster_kernel_run_in_pieces_and_long(oracle, lbm):
    """Several lbm_run calls (odd and even step counts: the lattice returns to HBM in between, the epoch flags carry on)
    and a longer run on the two-cluster case."""
    _check(oracle, lbm, 128, 256, 60, pieces=(1, 2, 7, 20))
    _check(oracle, lbm, 256, 256, 41, pieces=(5, 5))
    _check(oracle, lbm, 128, 256, 700)


def test_cluster_then_other_kernels_share_state(oracle, lbm, monkeypatch):
    """A state written back by the cluster kernel continues bit-exactly under the per-step kernels after save/load-free
    re-creation: run 20 steps clustered, download, upload into a per-step solver with first_accel already applied is not
    possible through the ABI -- so compare instead a clustered run against the SAME run with LBM_B200_CLUSTER=0."""
    iters = 50
    p, cells, obst = lbm.perturbed_problem(256, 128, iters, 99)
    out = {}
    for env in ("1", "0"):
        monkeypatch.setenv("LBM_B200_CLUSTER", env)
        with lbm.Solver(p) as s:
            assert ("cluster" in s.kernel_variant) == (env == "1")
            s.upload(cells, obst)
            s.run(iters)
            out[env] = s.download()
    assert np.array_equal(out["1"][0].view(np.uint32), out["0"][0].view(np.uint32))
    assert ulp_distance(out["1"][1], out["0"][1]).max() <= 1

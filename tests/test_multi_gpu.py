"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): row slabs on different B200s whose sweep
kernels store their halo rows straight into each other's memory over NVLink.  The lattice must be
bit-identical to the single-domain CPU oracle whatever the decomposition."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, ulp_distance
from rank_worker import problem

pytestmark = pytest.mark.gpu


def _need_gpus(lbm, n):
    have = lbm.device_count()
    if have < n:
        pytest.skip(f"needs {n} GPUs, {have} visible")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nx,ny,gpus", [(256, 64, 2), (100, 19, 2), (1024, 33, 2), (256, 96, 4), (384, 50, 8)])
def test_in_process_multi_gpu_bit_exact(oracle, lbm, nx, ny, gpus):
    """lbm_create(num_gpus = N): one process drives N GPUs (the reference's -n flag, main/LbmRunner.cpp:19)."""
    _need_gpus(lbm, gpus)
    iters = 40
    p, cells, obst = problem(lbm, nx, ny, iters, seed=nx + ny + gpus)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p, num_gpus=gpus) as s:
        s.upload(cells, obst)
        s.run(iters)
        got_cells, got_av = s.download()
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1


@pytest.mark.parametrize("nx,ny,world,pieces", [(256, 64, 2, 1), (130, 21, 2, 3), (2048, 512, 2, 1), (512, 100, 4, 2)])
def test_one_process_per_gpu_bit_exact(oracle, lbm, tmp_path, nx, ny, world, pieces):
    """torchrun, one rank per GPU, CUDA-IPC mapped neighbours (how bench.py --gpus N runs)."""
    _need_gpus(lbm, world)
    iters, seed = 30, 11
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "rank_worker.py"), str(tmp_path), str(nx), str(ny), str(iters), str(seed),
           str(pieces)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    p, cells, obst = problem(lbm, nx, ny, iters, seed)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    rows = 0
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        r0, got = int(z["row_begin"]), z["cells"]
        assert np.array_equal(got.view(np.uint32), want_cells[r0:r0 + got.shape[0]].view(np.uint32)), f"rank {rank}"
        assert ulp_distance(z["av_vels"], want_av).max() <= 1
        rows += got.shape[0]
    assert rows == ny


def test_lbm_runner_on_two_gpus_writes_the_same_files(lbm, tmp_path):
    """lbm_runner -n 2 (the reference's --num-ipus flag): same av_vels.dat and final_state.dat as on one GPU (the
    lattice is bit-identical; av_vels may differ in the last place because the slab partials are added in rank order)."""
    _need_gpus(lbm, 2)
    from conftest import obstacles_path, params_path
    outs = []
    for n in (1, 2):
        d = tmp_path / f"n{n}"
        d.mkdir()
        r = subprocess.run([lbm.RUNNER_PATH, "--params", params_path("128x256"), "--obstacles", obstacles_path("128x256"), "--outdir",
                            str(d), "--iters", "3000", "--reruns", "0", "-n", str(n), "-d"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("Slab ") == n
        outs.append(d)
    assert (outs[0] / "final_state.dat").read_bytes() == (outs[1] / "final_state.dat").read_bytes()
    a, b = (np.loadtxt(o / "av_vels.dat", usecols=[1]).astype(np.float32) for o in outs)
    assert ulp_distance(a, b).max() <= 1


@pytest.mark.parametrize("nx,ny,px,py", [(512, 64, 2, 1), (300, 50, 1, 2), (1024, 96, 2, 2), (640, 100, 2, 2), (1024, 128, 4, 2),
                                         (1024, 128, 2, 4), (777, 90, 4, 2)])
def test_block_partition_across_gpus_bit_exact(oracle, lbm, nx, ny, px, py):
    """lbm_create_grid with one block per GPU: ghost columns, ghost rows and corners travel over NVLink (2 x 2 on four GPUs,
    2 x 4 and 4 x 2 on eight; the reference's partitionForIpus layouts, main/include/StructuredGridUtils.hpp:501-527)."""
    _need_gpus(lbm, px * py)
    iters = 40
    p, cells, obst = problem(lbm, nx, ny, iters, seed=nx + ny + px * py)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p, grid=(px, py)) as s:
        assert sorted(b["device"] for b in s.blocks) == list(range(px * py))
        s.upload(cells, obst)
        s.run(17)
        s.run(iters - 17)
        got_cells, got_av = s.download()
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1


@pytest.mark.parametrize("nx,ny,px,py", [(512, 64, 2, 1), (300, 50, 1, 2), (640, 96, 2, 2), (1024, 128, 4, 2), (777, 90, 2, 4)])
def test_one_process_per_gpu_block_partition_bit_exact(oracle, lbm, tmp_path, nx, ny, px, py):
    """torchrun, one rank per GPU, each rank one BLOCK of a px x py partition (lbm_create_rank_grid): up to eight neighbours
    mapped with CUDA IPC, ghost rows / columns / corners stored into them by the sweep kernel."""
    world = px * py
    _need_gpus(lbm, world)
    iters, seed = 30, 23
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "rank_worker.py"), str(tmp_path), str(nx), str(ny), str(iters), str(seed), "2", f"{px}x{py}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    p, cells, obst = problem(lbm, nx, ny, iters, seed)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    covered = 0
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        r0, c0, got = int(z["row_begin"]), int(z["col_begin"]), z["cells"]
        want = want_cells[r0:r0 + got.shape[0], c0:c0 + got.shape[1]]
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32)), f"rank {rank}"
        assert ulp_distance(z["av_vels"], want_av).max() <= 1
        covered += got.shape[0] * got.shape[1]
    assert covered == nx * ny

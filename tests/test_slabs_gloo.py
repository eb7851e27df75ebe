"""CPU tests of the multi-rank host logic (lbm-graphcore_b200/slabs.py) -- world_size > 1 over gloo.
The GPU twin (kernels storing halo rows into peer memory) is tests/test_multi_gpu.py."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from gloo_worker import load_by_path

slabs = load_by_path("lbm_slabs_t", "lbm-graphcore_b200", "slabs.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_slab_rows_tile_the_grid_like_the_c_library():
    for ny in (2, 3, 19, 64, 16384, 32768 * 8):
        for world in (1, 2, 3, 4, 8):
            if ny < world:
                continue
            rows = [slabs.slab_rows(r, world, ny) for r in range(world)]
            assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
            for (r0, n), (r1, _) in zip(rows, rows[1:]):
                assert r0 + n == r1 and n >= 1
            assert max(n for _, n in rows) - min(n for _, n in rows) <= 1
            assert slabs.owner_of_row(ny - 2, world, ny) == max(r for r, (r0, _) in enumerate(rows) if r0 <= ny - 2)
    with pytest.raises(ValueError):
        slabs.slab_rows(0, 4, 3)


def test_ring_neighbours_wrap():
    assert slabs.ring_neighbours(0, 1) == (0, 0)
    assert slabs.ring_neighbours(0, 2) == (1, 1)
    assert slabs.ring_neighbours(0, 8) == (7, 1) and slabs.ring_neighbours(7, 8) == (6, 0)
    assert set(slabs.NORTH_PLANES) | set(slabs.SOUTH_PLANES) == {2, 5, 6, 4, 7, 8}


@pytest.mark.parametrize("world,nx,ny,iters", [(2, 24, 10, 25), (2, 7, 2, 12), (3, 16, 11, 20)])
def test_halo_plan_reproduces_single_domain_run_over_gloo(world, nx, ny, iters):
    """world ranks, each holding only its slab (everything else NaN), exchanging exactly the planes the CUDA
    kernel pushes to its neighbours; rank 0 asserts bit-equality with the single-domain oracle."""
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "gloo_worker.py"), str(nx), str(ny),
                                       str(iters), "5"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    assert all(pr.returncode == 0 for pr in procs), "\n".join(o[-3000:] for o in outs)
    assert "gloo slab model ok" in outs[0]


def test_two_phase_teardown_order_over_gloo(tmp_path):
    """slabs.close_ring (the protocol of include/lbm_b200.h): no rank drops its mappings while another may still be running,
    and no rank frees its slab while another still has it mapped -- whatever the skew between the ranks."""
    import json
    world, port = 3, _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "teardown_gloo_worker.py"), str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for pr in procs:
        out, _ = pr.communicate(timeout=120)
        assert pr.returncode == 0, out[-2000:]
    ev = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(world)]
    assert max(e["still_running_until"] for e in ev) <= min(e["disconnect"] for e in ev)
    assert max(e["disconnect"] for e in ev) <= min(e["close"] for e in ev)


def test_close_ring_without_a_group_just_closes():
    class S:
        closed = False

        def close(self):
            self.closed = True

        def ipc_disconnect(self):
            raise AssertionError("a single-process solver has nothing to disconnect")

    s = S()
    slabs.close_ring(s, None)
    assert s.closed

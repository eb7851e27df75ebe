"""CPU tests of the 2-D block partition's host-side plan (lbm-graphcore_b200/blocks.py): the arithmetic agrees with the C++
side, and the halo plan -- which populations of which cells go to which of the eight neighbours -- reproduces a single-domain
run when px * py ranks exchange exactly that over gloo.  The GPU twin is tests/test_blocks_gpu.py / test_multi_gpu.py."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from gloo_worker import load_by_path

blocks = load_by_path("lbm_blocks_t", "lbm-graphcore_b200", "blocks.py")
TOOL = os.path.join(ROOT, "lbm-graphcore_b200", "bin", "lbm_partition")


def test_blocks_tile_the_grid_and_agree_with_the_cxx_arithmetic():
    for nx, ny, px, py in ((1024, 1024, 2, 2), (777, 90, 4, 2), (5, 7, 5, 7), (16384, 16384, 2, 4)):
        bl = [blocks.block_of(i, px, py, nx, ny) for i in range(px * py)]
        assert sum(b["nxl"] * b["nyl"] for b in bl) == nx * ny
        assert max(b["nxl"] for b in bl) - min(b["nxl"] for b in bl) <= 1
        out = subprocess.run([TOOL, "stats", "--width", str(nx), "--height", str(ny), "--grid", f"{px}x{py}"], capture_output=True, text=True)
        head, row = (l.split(",") for l in out.stdout.strip().splitlines())
        rec = dict(zip(head, row))
        assert int(rec["largestBlockSize"]) == max(b["nxl"] * b["nyl"] for b in bl)
        assert int(rec["smallestBlockSize"]) == min(b["nxl"] * b["nyl"] for b in bl)
    for n in (2, 4, 8, 16, 3):
        for nx, ny in ((1024, 1024), (1001, 1024), (1024, 1001)):
            out = subprocess.run([TOOL, "stats", "--width", str(nx), "--height", str(ny), "--num-gpus", str(n), "--grid", "auto"],
                                 capture_output=True, text=True)
            head, row = (l.split(",") for l in out.stdout.strip().splitlines())
            rec = dict(zip(head, row))
            assert (int(rec["px"]), int(rec["py"])) == blocks.grid_for_gpus(nx, ny, n)


def test_neighbours_wrap_in_both_directions():
    assert blocks.neighbour(0, 2, 2, 1, 0) == 1 and blocks.neighbour(0, 2, 2, -1, 0) == 1
    assert blocks.neighbour(0, 2, 2, 0, 1) == 2 and blocks.neighbour(0, 2, 2, 1, 1) == 3 and blocks.neighbour(0, 2, 2, -1, -1) == 3
    assert blocks.neighbour(0, 1, 1, 1, 1) == 0
    assert blocks.neighbour(5, 4, 2, 1, 1) == 2 and blocks.neighbour(3, 4, 2, 1, 0) == 0
    assert set(blocks.EAST_PLANES) | set(blocks.WEST_PLANES) == {1, 5, 8, 3, 6, 7}
    assert blocks.CORNER_PLANE == {(1, 1): 5, (-1, 1): 6, (-1, -1): 7, (1, -1): 8}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nx,ny,px,py,iters", [(14, 10, 2, 2, 18), (9, 6, 2, 1, 12), (7, 8, 1, 2, 12), (5, 4, 2, 2, 9)])
def test_block_halo_plan_reproduces_single_domain_run_over_gloo(nx, ny, px, py, iters):
    world, port = px * py, _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "blocks_gloo_worker.py"), str(nx), str(ny), str(px),
                                       str(py), str(iters), "5"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
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
    assert "gloo block model ok" in outs[0]

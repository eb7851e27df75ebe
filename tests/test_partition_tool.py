"""lbm_partition (the B200 counterpart of the reference's tile_mapping_stats / viz_tile_map, main/TileMappingStats.cpp:54-101,
main/VisualiseTileMapping.cpp:66-205) and the partition arithmetic it shares with the solver (lbm_partition.hpp).  Host only."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT, obstacles_path

TOOL = os.path.join(ROOT, "lbm-graphcore_b200", "bin", "lbm_partition")


def run(*args):
    return subprocess.run([TOOL, *args], capture_output=True, text=True, timeout=120)


def rows(out):
    lines = out.strip().splitlines()
    head = lines[0].split(",")
    return [dict(zip(head, l.split(","))) for l in lines[1:]]


def test_stats_row_for_the_headline_grid():
    r = run("stats", "--width", "16384", "--height", "16384", "--num-gpus", "8")
    assert r.returncode == 0, r.stderr
    (row,) = rows(r.stdout)
    assert (row["numGpus"], row["px"], row["py"]) == ("8", "1", "8")  # row slabs unless --grid is given
    assert float(row["loadBalance"]) == 100.0 and float(row["maxSpeedup"]) == 8.0
    assert int(row["largestBlockSize"]) == 16384 * 2048 == int(row["smallestBlockSize"])
    # one CTA per 128 x 8 tile of the 2048 + 2 storage rows; 3 planes of the top and of the bottom row leave per step
    assert int(row["maxTilesPerGpu"]) == 128 * ((2048 + 2 + 7) // 8)
    assert int(row["haloBytesPerStep"]) == 2 * 3 * 16384 * 4
    assert float(row["wavesPerStep"]) == pytest.approx(int(row["maxTilesPerGpu"]) / (148 * 4), rel=1e-3)


def test_auto_grid_follows_the_reference_rule():
    """partitionForIpus (main/include/StructuredGridUtils.hpp:479-527): 4 -> 2 x 2; 8 -> 4 rows x 2 cols when the rows divide
    more evenly than the columns, else 2 rows x 4 cols; 2 likewise; 16 -> 4 x 4; other counts stay row slabs."""
    def grid(w, h, n):
        (row,) = rows(run("stats", "--width", str(w), "--height", str(h), "--num-gpus", str(n), "--grid", "auto").stdout)
        return int(row["px"]), int(row["py"])
    assert grid(1024, 1024, 4) == (2, 2) and grid(1024, 1024, 16) == (4, 4)
    assert grid(1024, 1024, 8) == (4, 2)        # equal imbalance: numRows = 2, numCols = 4
    assert grid(1001, 1024, 8) == (2, 4)        # rows divide evenly, columns do not: numRows = 4
    assert grid(1024, 1001, 8) == (4, 2)
    assert grid(1001, 1024, 2) == (1, 2) and grid(1024, 1001, 2) == (2, 1)
    assert grid(1024, 1024, 3) == (1, 3)


def test_uneven_cut_and_free_cell_balance():
    r = run("stats", "--width", "1024", "--height", "1024", "--num-gpus", "3", "--obstacles", obstacles_path("1024x1024"))
    assert r.returncode == 0, r.stderr
    (row,) = rows(r.stdout)
    assert int(row["largestBlockSize"]) == 342 * 1024 and int(row["smallestBlockSize"]) == 341 * 1024
    assert float(row["loadBalance"]) == pytest.approx(1024 * 1024 / 3 / (342 * 1024) * 100, rel=1e-3)
    assert int(row["freeCells"]) == 1043462
    assert 90 < float(row["freeCellLoadBalance"]) <= 100
    g = rows(run("stats", "--width", "1024", "--height", "1024", "--grid", "2x2").stdout)[0]
    assert int(g["haloBytesPerStep"]) == 2 * 3 * 512 * 4 + 2 * 3 * 512 * 4 + 16  # rows + ghost columns + four corners


def test_sampling_mode_and_errors():
    r = run("stats", "--sample", "--num-samples", "7", "--min-width", "100", "--max-width", "5000", "--min-height", "100",
            "--max-height", "5000", "--num-gpus", "4", "--grid", "auto")
    assert r.returncode == 0 and len(rows(r.stdout)) == 7
    for row in rows(r.stdout):
        assert 100 <= int(row["width"]) <= 5000 and 100 <= int(row["height"]) <= 5000 and 0 < float(row["loadBalance"]) <= 100
    assert run().returncode != 0 and "Show how a grid" in run().stderr
    assert run("stats", "--width", "16").returncode != 0
    assert run("map", "--width", "16", "--height", "16").returncode != 0  # no output file
    r = run("stats", "--width", "4", "--height", "4", "--grid", "8x1")
    assert r.returncode != 0 and "Couldn't fit" in r.stderr


def test_map_is_a_valid_png_of_the_partition(tmp_path):
    out = tmp_path / "map.png"
    r = run("map", "--width", "256", "--height", "256", "--num-gpus", "4", "--grid", "auto", "--obstacles", obstacles_path("256x256"),
            "-o", str(out))
    assert r.returncode == 0 and "256 rows by 256 cols" in r.stdout, r.stderr
    data = out.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(kind + body)
        chunks.append((kind, body))
        pos += 12 + n
    assert [k for k, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, colour = struct.unpack(">IIBB", chunks[0][1][:10])
    assert (w, h, depth, colour) == (256, 256, 8, 2)
    raw = np.frombuffer(zlib.decompress(chunks[1][1]), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert np.all(raw[:, 0] == 0)
    img = raw[:, 1:].reshape(h, w, 3)
    assert np.all(img[0] == 0) and np.all(img[-1] == 0) and np.all(img[:, 0] == 0)  # the box walls of the shipped input are solid
    assert np.all(img[127, 5:100] == 255) and np.all(img[5:100, 128] == 255)         # block borders of the 2 x 2 cut
    inner = {tuple(img[y, x]) for y, x in ((64, 64), (64, 192), (192, 64), (192, 192))}
    assert len(inner) == 4  # four blocks, four colours

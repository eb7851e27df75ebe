#!/usr/bin/env python
"""Time lbm_solve_host on the benchmark channel (development; LBM_B200_SOLVE_TRACE=1 prints the per-band phases)."""
import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lbm = importlib.import_module("lbm-graphcore_b200")
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nb = ny * nx * 36
hc, ho, out = lbm.host_alloc(nb, 0), lbm.host_alloc(ny * nx, 0), lbm.host_alloc(nb, 0)
cells = np.ctypeslib.as_array(ctypes.cast(hc, ctypes.POINTER(ctypes.c_float)), shape=(ny, nx, 9))
obst = np.ctypeslib.as_array(ctypes.cast(ho, ctypes.POINTER(ctypes.c_uint8)), shape=(ny, nx))
cells[...] = lbm.initial_cells(lbm.Params(1, 1, 1, 1, 0.1, 0.005, 1.85))[0, 0]
for r in range(0, ny, 1024):
    obst[r:r + 1024] = lbm.synthetic_obstacles(nx, ny, 1234, 0.01, r, min(1024, ny - r))
np.ctypeslib.as_array(ctypes.cast(out, ctypes.POINTER(ctypes.c_uint8)), shape=(nb,))[::4096] = 0  # touch
av = np.zeros(K, dtype=np.float32)
pipe = lbm.Pipeline(lbm.Params(nx, ny, K, 10, 0.1, 0.005, 1.85), K, 0)
print(pipe.info, flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    pipe.solve_raw(hc, ho, out, av.ctypes.data)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt:.3f} s wall, {nx * ny * K / dt / 1e6:.0f} MLUPS e2e, {(2 * nb + nx * ny) / dt / 1e9:.1f} GB/s over PCIe, av_last {av[-1]:.6e}", flush=True)
pipe.close()

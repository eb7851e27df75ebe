#!/usr/bin/env python
"""Generate tests/golden/oracle_<case>.npz from the CPU oracle -- and PIN the oracle while doing so.

For each shipped configuration (params/input_*.params + obstacles_*.dat of the reference, copied as
data into tests/golden/params/) this script

  1. runs the restated oracle (oracle/lbm_oracle.c, -ffp-contract=off) for the full maxIters,
  2. when oracle/_ref/compareLbm_strict exists (the UNMODIFIED reference main/LastChance.cpp built by
     `make -C oracle ref` in a container that has /root/reference), runs it too and asserts that its
     av_vels.dat and final_state.dat are byte-identical to the oracle's  (cases listed in PIN_CASES),
  3. stores what the GPU parity tests need, small enough to commit:
       av_f32sum   float32[maxIters]  av_vels with the reference's literal fp32 running sum
       av_f64sum   float32[maxIters]  av_vels with an fp64 accumulator (what the GPU computes)
       cells_sha256                   SHA-256 of the final AoS float32[ny][nx][9] lattice bytes
       pressure    float32[ny', nx']  final pressure field, subsampled by `stride`
       ux          float32[ny', nx']  final u_x, same subsampling
       cells_row   float32[nx, 9]     the final driven row ny-2 (carries the F8 extra-accelerate quirk)
       free_cells, stride, reynolds

Usage:  python tests/golden/make_oracle_fixtures.py [case ...]     (default: all four)
"""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

CASES = ["128x128", "128x256", "256x256", "1024x1024"]
PIN_CASES = ["128x128", "128x256", "256x256"]  # 1024x1024 takes ~10 min per program; pinned by the smaller ones
STRIDE = {"128x128": 1, "128x256": 1, "256x256": 2, "1024x1024": 8}
GOLD = os.path.join(ROOT, "tests", "golden")


def generate(case: str) -> None:
    pfile = os.path.join(GOLD, "params", f"input_{case}.params")
    ofile = os.path.join(GOLD, "params", f"obstacles_{case}.dat")
    p = oracle.load_params(pfile)
    obst = oracle.load_obstacles(ofile, p.nx, p.ny)
    cells0 = oracle.init_cells(p)
    t0 = time.time()
    final, av_f, av_d = oracle.run_both(p, cells0, obst, p.max_iters)
    dt = time.time() - t0
    print(f"[{case}] oracle: {p.max_iters} steps in {dt:.1f} s "
          f"({p.nx * p.ny * p.max_iters / dt / 1e6:.1f} MLUPS, 1 thread)", flush=True)

    if case in PIN_CASES and os.path.exists(oracle.REF_STRICT):
        with tempfile.TemporaryDirectory() as tmp:
            oracle.write_av_vels(os.path.join(tmp, "o_av_vels.dat"), av_f)
            oracle.write_final_state(os.path.join(tmp, "o_final_state.dat"), p, final, obst, transposed_flag=True)
            subprocess.run([oracle.REF_STRICT, pfile, ofile], cwd=tmp, check=True, capture_output=True)
            for name in ("av_vels.dat", "final_state.dat"):
                a = open(os.path.join(tmp, name), "rb").read()
                b = open(os.path.join(tmp, "o_" + name), "rb").read()
                assert a == b, f"{case}: oracle {name} differs from the unmodified reference program"
        print(f"[{case}] PINNED: byte-identical av_vels.dat and final_state.dat vs oracle/_ref/compareLbm_strict",
              flush=True)

    ux, uy, u, pressure = oracle.final_state_fields(p, final, obst)
    s = STRIDE[case]
    np.savez_compressed(
        os.path.join(GOLD, f"oracle_{case}.npz"),
        av_f32sum=av_f, av_f64sum=av_d,
        cells_sha256=np.array(hashlib.sha256(final.tobytes()).hexdigest()),
        pressure=pressure[::s, ::s].astype(np.float32), ux=ux[::s, ::s].astype(np.float32),
        cells_row=final[p.ny - 2].copy(), free_cells=np.int64((obst == 0).sum()), stride=np.int32(s),
        reynolds=np.float32(oracle.reynolds(p, oracle.av_velocity(p, final, obst))),
    )
    print(f"[{case}] wrote oracle_{case}.npz", flush=True)


if __name__ == "__main__":
    oracle.build()
    for c in (sys.argv[1:] or CASES):
        generate(c)

"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel of the
library runs at least once -- float4 and scalar sweeps, single slab and three slabs with flag-synchronised ghost rows,
upload/download transposition, mask packing, the synthetic channel, device-side final_state fields -- and every
result is still compared with the oracle.      compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lbm = importlib.import_module("lbm-graphcore_b200")
from oracle import oracle  # noqa: E402


def case(nx, ny, iters, devices=None, seed=0):
    rng = np.random.default_rng(seed)
    p = lbm.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    obst = (rng.random((ny, nx)) < 0.08).astype(np.uint8)
    cells = (lbm.initial_cells(p) * (1.0 + 0.1 * rng.random((ny, nx, 9))).astype(np.float32)).astype(np.float32)
    want, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    kw = {"devices": devices} if devices else {}
    with lbm.Solver(p, **kw) as s:
        s.upload(cells, obst)
        s.run(iters // 2)
        s.run(iters - iters // 2)
        got, av = s.download()
        fields = s.download_fields()
        variant = s.kernel_variant
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (nx, ny, devices)
    ux, uy, u, pr = oracle.final_state_fields(op, got, obst)
    assert np.array_equal(fields.view(np.uint32), np.stack([ux, uy, u, pr], -1).astype(np.float32).view(np.uint32))
    print(f"ok {nx}x{ny} x{iters} slabs={len(devices) if devices else 1} kernel={variant}", flush=True)


case(256, 20, 6)
case(100, 13, 6)
case(129, 7, 4)
case(300, 11, 5)
case(256, 30, 6, devices=[0, 0, 0])
case(77, 19, 6, devices=[0, 0])
p = lbm.Params(384, 40, 5, 10, 0.1, 0.005, 1.85)
with lbm.Solver(p) as s:
    s.init_synthetic(1234, 0.02)
    s.run(5)
    cells, av = s.download()
    assert np.all(np.isfinite(av)) and np.all(np.isfinite(cells))
print("ok synthetic", flush=True)

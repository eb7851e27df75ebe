"""lbm-graphcore_b200 -- Python face of the B200-native D2Q9-BGK solver.

The product is ``liblbm_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/lbm_b200.h``) plus the C++ ``lbm_runner`` driver.  This module is a thin
ctypes binding used by the tests, ``bench.py`` and ``__graft_entry__``; it mirrors the
reference's operator surface (main/LbmRunner.cpp:75-144 of thorbenlouw/lbm-graphcore):

    solver = Solver(params, num_gpus=1)   # Engine(...) + engine.load(device)
    solver.upload(cells, obstacles)       # engine.run(0)   ">>cells", ">>obstacles"
    solver.run(params.max_iters)          # engine.run(1)   the whole time loop
    cells, av_vels = solver.download()    # engine.run(2)   "<<cells", "<<av_vel"
    solver.elapsed_ms                     # readTimer

There is no CPU fallback: if the shared library is missing this module raises at import,
and every compute call raises ``LbmError`` when no CUDA device is usable.

The directory name contains a hyphen, so import it with
``importlib.import_module("lbm-graphcore_b200")``.
"""
from __future__ import annotations

import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "liblbm_b200.so")  # override: kernel tuning builds
HOST_LIB_PATH = os.path.join(_HERE, "liblbm_host.so")
RUNNER_PATH = os.path.join(_HERE, "bin", "lbm_runner")

# names every build of liblbm_b200.so must export (kept in sync with include/lbm_b200.h by
# tests/test_abi.py, which parses the header)
ABI_SYMBOLS = (
    "lbm_create", "lbm_upload", "lbm_run", "lbm_download", "lbm_elapsed_ms", "lbm_destroy",
    "lbm_last_error", "lbm_abi_version", "lbm_device_count", "lbm_steps_done", "lbm_free_cells",
    "lbm_launch_count", "lbm_kernel_variant", "lbm_create_on_devices", "lbm_init_synthetic", "lbm_rows",
    "lbm_download_tot_u", "lbm_host_alloc", "lbm_host_free", "lbm_host_alloc_near", "lbm_bind_thread_near", "lbm_create_rank", "lbm_ipc_blob_bytes",
    "lbm_ipc_export", "lbm_ipc_connect", "lbm_ipc_disconnect", "lbm_save", "lbm_load", "lbm_load_grid", "lbm_problem", "lbm_create_grid", "lbm_create_rank_grid", "lbm_partition", "lbm_block_info", "lbm_solve_host", "lbm_pipeline_create", "lbm_pipeline_solve",
    "lbm_pipeline_info", "lbm_pipeline_destroy", "lbm_pipeline_create_window", "lbm_pipeline_solve_window", "lbm_pipeline_plan", "lbm_selftest_math", "lbm_download_fields", "lbm_slab_count", "lbm_slab_info",
)


class LbmError(RuntimeError):
    """A failing call into liblbm_b200.so (carries the library's error code and message)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"liblbm_b200 error {code}: {message}")
        self.code = code


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU implementation to fall back to."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c_int, c_float, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
    P = ctypes.POINTER
    lib.lbm_create.argtypes = [P(c_void_p), c_int, c_int, c_int, c_float, c_float, c_float, c_int]
    lib.lbm_create_on_devices.argtypes = [P(c_void_p), c_int, c_int, c_int, c_float, c_float, c_float, c_int,
                                          P(c_int)]
    lib.lbm_create_rank.argtypes = [P(c_void_p), c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_int,
                                    c_int]
    lib.lbm_upload.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.lbm_init_synthetic.argtypes = [c_void_p, ctypes.c_uint64, c_float]
    lib.lbm_run.argtypes = [c_void_p, c_int]
    lib.lbm_download.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.lbm_download_tot_u.argtypes = [c_void_p, c_void_p]
    lib.lbm_download_fields.argtypes = [c_void_p, c_void_p]
    lib.lbm_slab_count.argtypes = [c_void_p, P(c_int)]
    lib.lbm_slab_info.argtypes = [c_void_p, c_int, P(c_int), P(c_int), P(c_int), P(ctypes.c_longlong)]
    lib.lbm_elapsed_ms.argtypes = [c_void_p, P(ctypes.c_double)]
    lib.lbm_destroy.argtypes = [c_void_p]
    lib.lbm_destroy.restype = None
    lib.lbm_last_error.restype = ctypes.c_char_p
    lib.lbm_device_count.argtypes = [P(c_int)]
    lib.lbm_steps_done.argtypes = [c_void_p, P(c_int)]
    lib.lbm_free_cells.argtypes = [c_void_p, P(ctypes.c_longlong)]
    lib.lbm_launch_count.argtypes = [c_void_p, P(ctypes.c_longlong)]
    lib.lbm_kernel_variant.argtypes = [c_void_p]
    lib.lbm_kernel_variant.restype = ctypes.c_char_p
    lib.lbm_rows.argtypes = [c_void_p, P(c_int), P(c_int)]
    lib.lbm_host_alloc.argtypes = [ctypes.c_size_t]
    lib.lbm_host_alloc.restype = c_void_p
    lib.lbm_host_free.argtypes = [c_void_p]
    lib.lbm_host_free.restype = None
    lib.lbm_host_alloc_near.argtypes = [ctypes.c_size_t, c_int]
    lib.lbm_host_alloc_near.restype = c_void_p
    lib.lbm_bind_thread_near.argtypes = [c_int]
    lib.lbm_ipc_blob_bytes.restype = ctypes.c_size_t
    lib.lbm_ipc_export.argtypes = [c_void_p, c_void_p]
    lib.lbm_ipc_connect.argtypes = [c_void_p, c_void_p]
    lib.lbm_ipc_disconnect.argtypes = [c_void_p]
    lib.lbm_create_grid.argtypes = [P(c_void_p), c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_int, P(c_int)]
    lib.lbm_create_rank_grid.argtypes = [P(c_void_p), c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_int, c_int, c_int]
    lib.lbm_partition.argtypes = [c_void_p, P(c_int), P(c_int)]
    lib.lbm_block_info.argtypes = [c_void_p, c_int, P(c_int), P(c_int), P(c_int), P(c_int), P(c_int), P(ctypes.c_longlong)]
    lib.lbm_solve_host.argtypes = [c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   P(ctypes.c_double)]
    lib.lbm_pipeline_create.argtypes = [P(c_void_p), c_int, c_int, c_int, c_float, c_float, c_float, c_int]
    lib.lbm_pipeline_solve.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lbm_pipeline_create_window.argtypes = [P(c_void_p), c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int]
    lib.lbm_pipeline_solve_window.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(ctypes.c_longlong)]
    lib.lbm_pipeline_info.argtypes = [c_void_p, P(c_int), P(c_int), P(c_int), P(ctypes.c_double)]
    lib.lbm_pipeline_plan.argtypes = [c_int, c_int, c_int, c_int, P(c_int), P(c_int), P(c_int)]
    lib.lbm_pipeline_destroy.argtypes = [c_void_p]
    lib.lbm_pipeline_destroy.restype = None
    lib.lbm_save.argtypes = [c_void_p, ctypes.c_char_p]
    lib.lbm_load.argtypes = [P(c_void_p), ctypes.c_char_p, c_int]
    lib.lbm_load_grid.argtypes = [P(c_void_p), ctypes.c_char_p, c_int, c_int, P(c_int)]
    lib.lbm_problem.argtypes = [c_void_p, P(c_int), P(c_int), P(c_int), P(c_float), P(c_float), P(c_float)]
    lib.lbm_selftest_math.argtypes = [c_int, ctypes.c_ulonglong, P(ctypes.c_ulonglong)]
    return lib


lib = _load()


def _check(rc: int) -> None:
    if rc != 0:
        raise LbmError(rc, lib.lbm_last_error().decode("utf-8", "replace"))


def selftest_math(device: int = 0, pairs: int = 1 << 32) -> dict:
    """Compare the kernels' division / sqrt helpers with CUDA's IEEE library calls on the GPU."""
    c = (ctypes.c_ulonglong * 7)()
    _check(lib.lbm_selftest_math(device, pairs, c))
    return {"div9_bad": c[0], "sqrt_bad": c[1], "div_bad": c[2], "div36_bad": c[3], "div9_n": c[4], "sqrt_n": c[5],
            "div_n": c[6]}


def host_alloc(nbytes: int, device: int | None = None) -> int:
    """Pinned host memory (address, or 0/None on failure); near GPU `device`'s NUMA node when given. Free with lib.lbm_host_free."""
    return lib.lbm_host_alloc(nbytes) if device is None else lib.lbm_host_alloc_near(nbytes, device)


def pipeline_plan(nx: int, ny: int, iters: int, window_rows: int = 0) -> dict:
    """lbm_pipeline_plan: how a job would be cut into bands (needs no GPU)."""
    b, r, m = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(lib.lbm_pipeline_plan(nx, ny, window_rows, iters, ctypes.byref(b), ctypes.byref(r), ctypes.byref(m)))
    return {"bands": b.value, "band_rows": r.value, "margin_rows": m.value}


class Pipeline:
    """lbm_pipeline: whole jobs (upload, `iters` timesteps, download) of one problem shape, band-pipelined over PCIe."""

    def __init__(self, params: Params, iters: int | None = None, device: int = 0, window: tuple[int, int] | None = None):
        """window = (row_begin, row_count): one rank's slab of the grid (lbm_pipeline_create_window)."""
        p = params
        self.params, self.iters, self.window = p, (p.max_iters if iters is None else iters), window
        self._p = ctypes.c_void_p()
        if window is None:
            _check(lib.lbm_pipeline_create(ctypes.byref(self._p), p.nx, p.ny, self.iters, p.density, p.accel, p.omega, device))
        else:
            _check(lib.lbm_pipeline_create_window(ctypes.byref(self._p), p.nx, p.ny, window[0], window[1], self.iters, p.density,
                                                  p.accel, p.omega, device))

    def solve_window_raw(self, cells_ptr: int, obstacles_ptr: int, cells_out_ptr: int):
        """One rank's job: returns (tot_u float64[iters], free_cells) of the window."""
        tot = np.zeros(max(self.iters, 1), dtype=np.float64)
        free = ctypes.c_longlong()
        _check(lib.lbm_pipeline_solve_window(self._p, cells_ptr, obstacles_ptr, cells_out_ptr, tot.ctypes.data, ctypes.byref(free)))
        return tot[: self.iters], free.value

    def solve_window(self, cells: np.ndarray, obstacles: np.ndarray):
        """cells / obstacles: the WHOLE grid (convenience for tests): the window plus its margins is cut out here."""
        r0, n = self.window
        m, ny = self.info["margin_rows"], self.params.ny
        idx = np.arange(r0 - m, r0 + n + m) % ny
        win_c = np.ascontiguousarray(cells[idx], dtype=np.float32)
        win_o = np.ascontiguousarray(obstacles[idx]).astype(np.uint8, copy=False)
        out = np.empty((n, self.params.nx, NSPEEDS), dtype=np.float32)
        tot, free = self.solve_window_raw(win_c.ctypes.data, win_o.ctypes.data, out.ctypes.data)
        return out, tot, free

    def solve_raw(self, cells_in_ptr: int, obstacles_ptr: int, cells_out_ptr: int, av_vels_ptr: int | None) -> None:
        _check(lib.lbm_pipeline_solve(self._p, cells_in_ptr, obstacles_ptr, cells_out_ptr, av_vels_ptr))

    def solve(self, cells: np.ndarray, obstacles: np.ndarray):
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        obstacles = np.ascontiguousarray(obstacles).astype(np.uint8, copy=False)
        out = np.empty_like(cells)
        av = np.zeros(max(self.iters, 1), dtype=np.float32)
        self.solve_raw(cells.ctypes.data, obstacles.ctypes.data, out.ctypes.data, av.ctypes.data)
        return out, av[: self.iters]

    @property
    def info(self) -> dict:
        b, r, m, ms = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        _check(lib.lbm_pipeline_info(self._p, ctypes.byref(b), ctypes.byref(r), ctypes.byref(m), ctypes.byref(ms)))
        return {"bands": b.value, "band_rows": r.value, "margin_rows": m.value, "last_ms": ms.value}

    def close(self) -> None:
        if self._p:
            lib.lbm_pipeline_destroy(self._p)
            self._p = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_host(params: Params, cells: np.ndarray, obstacles: np.ndarray, iters: int | None = None, device: int = 0):
    """lbm_solve_host: the whole job (upload, `iters` timesteps, download) as one pipelined call.  Returns
    (cells float32[ny, nx, 9], av_vels float32[iters], elapsed_ms)."""
    p = params
    iters = p.max_iters if iters is None else iters
    cells = np.ascontiguousarray(cells, dtype=np.float32)
    obstacles = np.ascontiguousarray(obstacles).astype(np.uint8, copy=False)
    out = np.empty_like(cells)
    av = np.zeros(max(iters, 1), dtype=np.float32)
    ms = ctypes.c_double()
    _check(lib.lbm_solve_host(p.nx, p.ny, iters, p.density, p.accel, p.omega, device, cells.ctypes.data, obstacles.ctypes.data,
                              out.ctypes.data, av.ctypes.data, ctypes.byref(ms)))
    return out, av[:iters], ms.value


def device_count() -> int:
    """Number of CUDA devices the library can see (0 when there is none)."""
    n = ctypes.c_int(0)
    rc = lib.lbm_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


from .problem import (NSPEEDS, Params, initial_cells, load_obstacles, perturbed_problem,  # noqa: E402,F401
                      synthetic_obstacles)


# ---------------------------------------------------------------------------------------------------
# The solver handle
# ---------------------------------------------------------------------------------------------------
class Solver:
    """One ``lbm_handle``.  ``devices`` (a list that may repeat a GPU) selects
    ``lbm_create_on_devices``; ``rank``/``world`` select the one-process-per-GPU mode."""

    def __init__(self, params: Params, num_gpus: int = 1, devices: list[int] | None = None,
                 rank: int | None = None, world: int | None = None, device: int = 0,
                 grid: tuple[int, int] | None = None):
        self.params = params
        self._h = ctypes.c_void_p()
        p = params
        args = (p.nx, p.ny, p.max_iters, p.density, p.accel, p.omega)
        if grid is not None and rank is not None:  # one process per GPU, this rank owns block `rank` of the px x py grid
            _check(lib.lbm_create_rank_grid(ctypes.byref(self._h), *args, rank, grid[0], grid[1], device))
        elif grid is not None:  # (px, py) blocks; devices = GPU of block by * px + bx (default: one GPU each)
            px, py = grid
            arr = (ctypes.c_int * (px * py))(*devices) if devices is not None else None
            _check(lib.lbm_create_grid(ctypes.byref(self._h), *args, px, py, arr))
        elif rank is not None:
            _check(lib.lbm_create_rank(ctypes.byref(self._h), *args, rank, world, device))
        elif devices is not None:
            arr = (ctypes.c_int * len(devices))(*devices)
            _check(lib.lbm_create_on_devices(ctypes.byref(self._h), *args, len(devices), arr))
        else:
            _check(lib.lbm_create(ctypes.byref(self._h), *args, num_gpus))
        rb, rc = ctypes.c_int(), ctypes.c_int()
        _check(lib.lbm_rows(self._h, ctypes.byref(rb), ctypes.byref(rc)))
        self.row_begin, self.row_count = rb.value, rc.value
        # the columns a caller passes per row: all of them, except for a rank that owns one block of a block partition
        self.col_begin, self.col_count = 0, p.nx
        if grid is not None and rank is not None:
            b = self.blocks[0]
            self.col_begin, self.col_count = b["col_begin"], b["col_count"]

    @classmethod
    def load(cls, path: str, num_gpus: int = 1, reynolds_dim: int = 10, grid: tuple[int, int] | None = None,
             devices: list[int] | None = None) -> "Solver":
        """lbm_load / lbm_load_grid: a solver restored from a file written by save() (the reference's --exe)."""
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        if grid is not None:
            arr = (ctypes.c_int * (grid[0] * grid[1]))(*devices) if devices is not None else None
            _check(lib.lbm_load_grid(ctypes.byref(self._h), os.fsencode(path), grid[0], grid[1], arr))
        else:
            _check(lib.lbm_load(ctypes.byref(self._h), os.fsencode(path), num_gpus))
        nx, ny, it = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        d, a, o = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        _check(lib.lbm_problem(self._h, ctypes.byref(nx), ctypes.byref(ny), ctypes.byref(it), ctypes.byref(d),
                               ctypes.byref(a), ctypes.byref(o)))
        self.params = Params(nx.value, ny.value, it.value, reynolds_dim, d.value, a.value, o.value)
        self.row_begin, self.row_count = 0, ny.value
        self.col_begin, self.col_count = 0, nx.value
        return self

    def save(self, path: str) -> None:
        """lbm_save: problem + current state, resumable with Solver.load."""
        _check(lib.lbm_save(self._h, os.fsencode(path)))

    # -- lifetime -----------------------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            lib.lbm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference's three programs -------------------------------------------------------------
    def upload(self, cells: np.ndarray, obstacles: np.ndarray) -> None:
        """engine.run(0).  cells float32[rows, nx, 9], obstacles uint8/bool[rows, nx] for the rows this
        handle owns (all of them unless in rank mode)."""
        p = self.params
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        obstacles = np.ascontiguousarray(obstacles).astype(np.uint8, copy=False)
        if cells.size != self.row_count * self.col_count * NSPEEDS or obstacles.size != self.row_count * self.col_count:
            raise ValueError("cells/obstacles do not match the rows (and columns) owned by this solver")
        _check(lib.lbm_upload(self._h, cells.ctypes.data, obstacles.ctypes.data))

    def upload_raw(self, cells_ptr: int, obstacles_ptr: int) -> None:
        _check(lib.lbm_upload(self._h, cells_ptr, obstacles_ptr))

    def init_synthetic(self, seed: int, solid_prob: float) -> None:
        _check(lib.lbm_init_synthetic(self._h, seed, solid_prob))

    def run(self, iters: int) -> None:
        """engine.run(1): `iters` timesteps (the first call also does the one-off accelerate_flow)."""
        _check(lib.lbm_run(self._h, iters))

    def download(self, cells: bool = True, av_vels: bool = True):
        """engine.run(2): returns (cells float32[rows, nx, 9] | None, av_vels float32[steps] | None)."""
        p = self.params
        c = np.empty((self.row_count, self.col_count, NSPEEDS), dtype=np.float32) if cells else None
        a = np.zeros(max(p.max_iters, 1), dtype=np.float32) if av_vels else None
        _check(lib.lbm_download(self._h, c.ctypes.data if cells else None, a.ctypes.data if av_vels else None))
        return c, (a[: self.steps_done] if av_vels else None)

    def download_raw(self, cells_ptr: int, av_vels_ptr: int) -> None:
        _check(lib.lbm_download(self._h, cells_ptr, av_vels_ptr))

    def download_fields(self) -> np.ndarray:
        """u_x, u_y, u, pressure per cell, computed on the device: float32[rows, nx, 4]."""
        f = np.empty((self.row_count, self.col_count, 4), dtype=np.float32)
        _check(lib.lbm_download_fields(self._h, f.ctypes.data))
        return f

    def download_tot_u(self) -> np.ndarray:
        t = np.zeros(max(self.steps_done, 1), dtype=np.float64)
        _check(lib.lbm_download_tot_u(self._h, t.ctypes.data))
        return t[: self.steps_done]

    # -- introspection ------------------------------------------------------------------------------
    @property
    def elapsed_ms(self) -> float:
        ms = ctypes.c_double()
        _check(lib.lbm_elapsed_ms(self._h, ctypes.byref(ms)))
        return ms.value

    @property
    def steps_done(self) -> int:
        n = ctypes.c_int()
        _check(lib.lbm_steps_done(self._h, ctypes.byref(n)))
        return n.value

    @property
    def free_cells(self) -> int:
        n = ctypes.c_longlong()
        _check(lib.lbm_free_cells(self._h, ctypes.byref(n)))
        return n.value

    @property
    def launch_count(self) -> int:
        n = ctypes.c_longlong()
        _check(lib.lbm_launch_count(self._h, ctypes.byref(n)))
        return n.value

    @property
    def slabs(self) -> list[dict]:
        """Per row slab: GPU, global rows, non-obstacle cells (the partition diagnostics of lbm_slab_info)."""
        n = ctypes.c_int()
        _check(lib.lbm_slab_count(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            d, r0, rc, fc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
            _check(lib.lbm_slab_info(self._h, i, ctypes.byref(d), ctypes.byref(r0), ctypes.byref(rc), ctypes.byref(fc)))
            out.append({"device": d.value, "row_begin": r0.value, "row_count": rc.value, "free_cells": fc.value})
        return out

    @property
    def blocks(self) -> list[dict]:
        """Per block of the partition: GPU, global rows and columns, non-obstacle cells (lbm_block_info)."""
        n = ctypes.c_int()
        _check(lib.lbm_slab_count(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            d, r0, rc, c0, cc, fc = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(),
                                     ctypes.c_longlong())
            _check(lib.lbm_block_info(self._h, i, ctypes.byref(d), ctypes.byref(r0), ctypes.byref(rc), ctypes.byref(c0),
                                      ctypes.byref(cc), ctypes.byref(fc)))
            out.append({"device": d.value, "row_begin": r0.value, "row_count": rc.value, "col_begin": c0.value,
                        "col_count": cc.value, "free_cells": fc.value})
        return out

    @property
    def partition(self) -> tuple[int, int]:
        px, py = ctypes.c_int(), ctypes.c_int()
        _check(lib.lbm_partition(self._h, ctypes.byref(px), ctypes.byref(py)))
        return px.value, py.value

    @property
    def kernel_variant(self) -> str:
        return lib.lbm_kernel_variant(self._h).decode()

    # -- one process per GPU ------------------------------------------------------------------------
    def ipc_export(self) -> bytes:
        n = lib.lbm_ipc_blob_bytes()
        buf = ctypes.create_string_buffer(n)
        _check(lib.lbm_ipc_export(self._h, buf))
        return buf.raw

    def ipc_connect(self, blobs: bytes) -> None:
        buf = ctypes.create_string_buffer(blobs, len(blobs))
        _check(lib.lbm_ipc_connect(self._h, buf))

    def ipc_disconnect(self) -> None:
        """First half of the rank-mode teardown: close the mappings of the neighbours' slabs (then barrier, then close)."""
        _check(lib.lbm_ipc_disconnect(self._h))

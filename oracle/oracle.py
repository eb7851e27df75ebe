"""ctypes face of oracle/liblbm_oracle.so -- the CPU restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Import this from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py, never from the product package.
See lbm_oracle.h for the reference citations of every function.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblbm_oracle.so")
CLI_PATH = os.path.join(_HERE, "oracle_lbm")
REF_STRICT = os.path.join(_HERE, "_ref", "compareLbm_strict")  # unmodified LastChance.cpp, -ffp-contract=off
REF_FMA = os.path.join(_HERE, "_ref", "compareLbm_fma")        # unmodified LastChance.cpp, reference flags

ACCUM_FLOAT, ACCUM_DOUBLE = 0, 1


class OracleParams(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("max_iters", ctypes.c_int),
                ("reynolds_dim", ctypes.c_int), ("density", ctypes.c_float), ("accel", ctypes.c_float),
                ("omega", ctypes.c_float)]


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or not os.path.exists(CLI_PATH):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        PP = ctypes.POINTER(OracleParams)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.oracle_load_params.argtypes = [ctypes.c_char_p, PP]
        L.oracle_load_obstacles.argtypes = [ctypes.c_char_p, ci, ci, vp]
        L.oracle_load_obstacles.restype = ctypes.c_long
        L.oracle_count_free_cells.argtypes = [vp, ci, ci]
        L.oracle_count_free_cells.restype = ctypes.c_long
        L.oracle_init_cells.argtypes = [PP, vp]
        L.oracle_init_cells.restype = None
        L.oracle_rebound_cell.argtypes = [vp, vp]
        L.oracle_rebound_cell.restype = None
        L.oracle_collide_cell.argtypes = [vp, ctypes.c_float, vp]
        L.oracle_collide_cell.restype = ctypes.c_float
        L.oracle_accelerate_flow.argtypes = [PP, vp, vp]
        L.oracle_accelerate_flow.restype = None
        L.oracle_timestep_rows.argtypes = [PP, vp, vp, vp, ci, ci, ci, ci]
        L.oracle_timestep_rows.restype = ctypes.c_double
        L.oracle_run.argtypes = [PP, vp, vp, vp, ci, vp, ci, ci, ci]
        L.oracle_run.restype = vp
        L.oracle_run_both.argtypes = [PP, vp, vp, vp, ci, vp, vp, ci]
        L.oracle_run_both.restype = vp
        L.oracle_av_velocity.argtypes = [PP, vp, vp]
        L.oracle_av_velocity.restype = ctypes.c_float
        L.oracle_reynolds.argtypes = [PP, ctypes.c_float]
        L.oracle_reynolds.restype = ctypes.c_float
        L.oracle_total_density.argtypes = [PP, vp]
        L.oracle_total_density.restype = ctypes.c_float
        L.oracle_write_final_state.argtypes = [ctypes.c_char_p, PP, vp, vp, ci]
        L.oracle_write_av_vels.argtypes = [ctypes.c_char_p, vp, ci]
        _lib = L
    return _lib


def make_params(nx, ny, max_iters, reynolds_dim, density, accel, omega) -> OracleParams:
    return OracleParams(int(nx), int(ny), int(max_iters), int(reynolds_dim), float(density), float(accel),
                        float(omega))


def load_params(path: str) -> OracleParams:
    p = OracleParams()
    rc = lib().oracle_load_params(path.encode(), ctypes.byref(p))
    if rc:
        raise ValueError(f"oracle_load_params({path}) failed with {rc}")
    return p


def load_obstacles(path: str, nx: int, ny: int) -> np.ndarray:
    obst = np.zeros((ny, nx), dtype=np.uint8)
    rc = lib().oracle_load_obstacles(path.encode(), nx, ny, obst.ctypes.data)
    if rc < 0:
        raise ValueError(f"oracle_load_obstacles({path}) failed with {rc}")
    return obst


def init_cells(p: OracleParams) -> np.ndarray:
    cells = np.empty((p.ny, p.nx, 9), dtype=np.float32)
    lib().oracle_init_cells(ctypes.byref(p), cells.ctypes.data)
    return cells


def rebound_cell(s) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.float32)
    out = np.empty(9, dtype=np.float32)
    lib().oracle_rebound_cell(s.ctypes.data, out.ctypes.data)
    return out


def collide_cell(s, omega: float):
    s = np.ascontiguousarray(s, dtype=np.float32)
    out = np.empty(9, dtype=np.float32)
    u_sq = lib().oracle_collide_cell(s.ctypes.data, float(omega), out.ctypes.data)
    return out, u_sq


def accelerate_flow(p: OracleParams, cells: np.ndarray, obst: np.ndarray) -> None:
    assert cells.dtype == np.float32 and cells.flags.c_contiguous
    obst = np.ascontiguousarray(obst, dtype=np.uint8)
    lib().oracle_accelerate_flow(ctypes.byref(p), cells.ctypes.data, obst.ctypes.data)


def timestep_rows(p: OracleParams, src: np.ndarray, dst: np.ndarray, obst: np.ndarray, row_begin: int,
                  row_end: int, accum: int = ACCUM_DOUBLE, threads: int = 1) -> float:
    assert src.dtype == np.float32 and dst.dtype == np.float32 and obst.dtype == np.uint8
    return lib().oracle_timestep_rows(ctypes.byref(p), src.ctypes.data, dst.ctypes.data, obst.ctypes.data,
                                      row_begin, row_end, accum, threads)


def run(p: OracleParams, cells: np.ndarray, obst: np.ndarray, iters: int, accum: int = ACCUM_DOUBLE,
        threads: int = 1, first_accel: bool = True):
    """Runs `iters` steps from `cells` (not modified).  Returns (final cells, av_vels float32[iters])."""
    a = np.ascontiguousarray(cells, dtype=np.float32).copy()
    b = np.empty_like(a)
    obst = np.ascontiguousarray(obst, dtype=np.uint8)
    av = np.zeros(max(iters, 1), dtype=np.float32)
    fin = lib().oracle_run(ctypes.byref(p), a.ctypes.data, b.ctypes.data, obst.ctypes.data, iters,
                           av.ctypes.data, accum, threads, 1 if first_accel else 0)
    final = a if fin == a.ctypes.data else b
    return final, av[:iters]


def run_both(p: OracleParams, cells: np.ndarray, obst: np.ndarray, iters: int, first_accel: bool = True):
    """Like run() but returns (final cells, av_vels with fp32 accumulation, av_vels with fp64 accumulation)."""
    a = np.ascontiguousarray(cells, dtype=np.float32).copy()
    b = np.empty_like(a)
    obst = np.ascontiguousarray(obst, dtype=np.uint8)
    avf = np.zeros(max(iters, 1), dtype=np.float32)
    avd = np.zeros(max(iters, 1), dtype=np.float32)
    fin = lib().oracle_run_both(ctypes.byref(p), a.ctypes.data, b.ctypes.data, obst.ctypes.data, iters,
                                avf.ctypes.data, avd.ctypes.data, 1 if first_accel else 0)
    final = a if fin == a.ctypes.data else b
    return final, avf[:iters], avd[:iters]


def av_velocity(p: OracleParams, cells: np.ndarray, obst: np.ndarray) -> float:
    return lib().oracle_av_velocity(ctypes.byref(p), cells.ctypes.data, obst.ctypes.data)


def reynolds(p: OracleParams, av: float) -> float:
    return lib().oracle_reynolds(ctypes.byref(p), float(av))


def total_density(p: OracleParams, cells: np.ndarray) -> float:
    return lib().oracle_total_density(ctypes.byref(p), cells.ctypes.data)


def write_final_state(path: str, p: OracleParams, cells: np.ndarray, obst: np.ndarray,
                      transposed_flag: bool = False) -> None:
    rc = lib().oracle_write_final_state(path.encode(), ctypes.byref(p), cells.ctypes.data, obst.ctypes.data,
                                        1 if transposed_flag else 0)
    if rc:
        raise OSError(f"cannot write {path}")


def write_av_vels(path: str, av_vels: np.ndarray) -> None:
    av = np.ascontiguousarray(av_vels, dtype=np.float32)
    rc = lib().oracle_write_av_vels(path.encode(), av.ctypes.data, int(av.size))
    if rc:
        raise OSError(f"cannot write {path}")


def final_state_fields(p: OracleParams, cells: np.ndarray, obst: np.ndarray):
    """numpy restatement of the per-cell fields of write_values (main/LastChance.cpp:572-611):
    returns (u_x, u_y, u, pressure) float32[ny, nx]."""
    f = np.float32
    s = cells.astype(np.float32)
    rho = np.zeros(s.shape[:2], dtype=np.float32)
    for k in range(9):
        rho = rho + s[..., k]
    with np.errstate(divide="ignore", invalid="ignore"):
        ux = (s[..., 1] + s[..., 5] + s[..., 8] - (s[..., 3] + s[..., 6] + s[..., 7])) / rho
        uy = (s[..., 2] + s[..., 5] + s[..., 6] - (s[..., 4] + s[..., 7] + s[..., 8])) / rho
        u = np.sqrt(ux * ux + uy * uy)
    pressure = rho * (f(1.0) / f(3.0))
    solid = obst.astype(bool)
    ux = np.where(solid, f(0), ux)
    uy = np.where(solid, f(0), uy)
    u = np.where(solid, f(0), u)
    pressure = np.where(solid, f(p.density) * (f(1.0) / f(3.0)), pressure)
    return ux, uy, u, pressure

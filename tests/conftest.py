"""Shared fixtures.  `-m "not gpu"` runs here on CPU (oracle vs goldens, host logic, ABI exports);
`-m gpu` runs on a B200 and compares the CUDA path, called through the C ABI, with the oracle."""
import gzip
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["128x128", "128x256", "256x256", "1024x1024"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU-side check")


def params_path(case):
    return os.path.join(GOLDEN, "params", f"input_{case}.params")


def obstacles_path(case):
    return os.path.join(GOLDEN, "params", f"obstacles_{case}.dat")


def golden_text(name, tmpdir):
    """Decompress tests/golden/check/<name>.gz (a data file of the reference's check/ directory)."""
    out = os.path.join(str(tmpdir), name)
    with gzip.open(os.path.join(GOLDEN, "check", name + ".gz"), "rb") as f, open(out, "wb") as g:
        shutil.copyfileobj(f, g)
    return out


def golden_av_vels(case):
    with gzip.open(os.path.join(GOLDEN, "check", f"{case}.av_vels.dat.gz"), "rt") as f:
        return np.loadtxt(f, usecols=[1])


def golden_final_state(case):
    with gzip.open(os.path.join(GOLDEN, "check", f"{case}.final_state.dat.gz"), "rt") as f:
        return np.loadtxt(f, usecols=[0, 1, 5])


def oracle_fixture(case):
    return np.load(os.path.join(GOLDEN, f"oracle_{case}.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def lbm():
    """The product package (ctypes binding of liblbm_b200.so); fails loudly if the .so is missing."""
    return importlib.import_module("lbm-graphcore_b200")


def ulp_distance(a, b):
    """Distance in units of float32 ulps between two float32 arrays (same-sign finite values)."""
    ai = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ai - bi)

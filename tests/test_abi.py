"""The drop-in boundary on a machine WITHOUT a GPU: liblbm_b200.so loads, exports every symbol that
include/lbm_b200.h declares (and nothing undeclared with the lbm_ prefix), reports errors through codes
instead of exceptions, and refuses to compute -- there is no CPU path to fall back to."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "lbm_b200.h")
HOST_HEADER = os.path.join(ROOT, "include", "lbm_host.h")


def declared_functions(path):
    text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", text)))


def exported_functions(lib_path, prefix):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    return sorted(line.split()[-1] for line in out.splitlines() if " T " in line and line.split()[-1].startswith(prefix))


def test_header_and_library_agree(lbm):
    declared = declared_functions(HEADER)
    assert len(declared) >= 20
    assert declared == sorted(lbm.ABI_SYMBOLS), "lbm-graphcore_b200/__init__.py ABI_SYMBOLS is out of sync with the header"
    assert exported_functions(lbm.LIB_PATH, "lbm_") == declared
    for name in declared:
        assert hasattr(lbm.lib, name)
    assert lbm.lib.lbm_abi_version() == int(re.search(r"#define LBM_B200_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_host_library_exports_its_header(lbm):
    declared = declared_functions(HOST_HEADER)
    assert declared == exported_functions(lbm.HOST_LIB_PATH, "lbm_host_")


def test_header_is_plain_c(tmp_path):
    """The boundary must be consumable from C (cgo / JNI / ctypes style bindings): compile it with gcc -std=c99."""
    src = tmp_path / "use.c"
    src.write_text('#include "lbm_b200.h"\n#include "lbm_host.h"\nint main(void) { lbm_handle *h = 0; (void)h; return LBM_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                    str(tmp_path / "use.o")], check=True, capture_output=True)


def test_no_gpu_means_loud_failure_not_a_cpu_fallback(lbm):
    if lbm.device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    p = lbm.Params(128, 128, 10, 10, 0.1, 0.005, 1.85)
    with pytest.raises(lbm.LbmError) as e:
        lbm.Solver(p)
    assert e.value.code == -2 and "CUDA" in str(e.value)
    n = ctypes.c_int(7)
    assert lbm.lib.lbm_device_count(ctypes.byref(n)) == -2 and n.value == 0
    c = (ctypes.c_ulonglong * 7)()
    assert lbm.lib.lbm_selftest_math(0, 16, c) == -2
    # the driver binary fails the same way
    gold = os.path.join(ROOT, "tests", "golden", "params")
    r = subprocess.run([lbm.RUNNER_PATH, "--params", os.path.join(gold, "input_128x128.params"), "--obstacles",
                        os.path.join(gold, "obstacles_128x128.dat")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr


def test_argument_errors_do_not_need_a_gpu(lbm):
    h = ctypes.c_void_p()
    assert lbm.lib.lbm_create(ctypes.byref(h), 0, 5, 1, 0.1, 0.005, 1.85, 1) != 0 and not h.value
    assert lbm.lib.lbm_create(ctypes.byref(h), 128, 128, 1, 0.1, 0.005, 1.85, 0) == -1
    assert b"num_gpus" in lbm.lib.lbm_last_error()
    assert lbm.lib.lbm_run(None, 1) == -1 and lbm.lib.lbm_upload(None, None, None) == -1
    assert lbm.lib.lbm_download(None, None, None) == -1
    lbm.lib.lbm_destroy(None)  # harmless
    assert lbm.lib.lbm_ipc_blob_bytes() >= 64


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the product package or the C ABI sources may reference it."""
    pkg = os.path.join(ROOT, "lbm-graphcore_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle|#\s*include\s*[<\"][^>\"]*oracle|(CDLL|dlopen)\([^)]*oracle|-llbm_oracle",
                             text, flags=re.M):
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders
    ldd = subprocess.run(["ldd", os.path.join(pkg, "liblbm_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "torch" not in ldd


def test_runner_cli_errors_before_any_gpu_work(lbm, tmp_path):
    """Flag handling of the LbmRunner-style driver (main/LbmRunner.cpp:16-48): both input files are mandatory, unknown
    flags print the usage, unreadable or malformed inputs abort with the reference's messages -- all before CUDA."""
    run = lambda *a: subprocess.run([lbm.RUNNER_PATH, *a], capture_output=True, text=True)
    gold = os.path.join(ROOT, "tests", "golden", "params")
    r = run()
    assert r.returncode != 0 and "--params" in r.stderr and "--obstacles" in r.stderr
    r = run("--params", os.path.join(gold, "input_128x128.params"))
    assert r.returncode != 0 and "--obstacles" in r.stderr
    r = run("--bogus")
    assert r.returncode != 0 and "Runs a Lattice Boltzmann simulation" in r.stderr
    r = run("--params", str(tmp_path / "nope.params"), "--obstacles", os.path.join(gold, "obstacles_128x128.dat"))
    assert r.returncode != 0 and "Could not parse parameters file" in r.stderr
    bad = tmp_path / "bad.params"
    bad.write_text("128\n128\n")
    r = run("--params", str(bad), "--obstacles", os.path.join(gold, "obstacles_128x128.dat"))
    assert r.returncode != 0 and "Could not parse parameters file" in r.stderr
    r = run("--params", os.path.join(gold, "input_128x128.params"), "--obstacles", str(tmp_path / "nope.dat"))
    assert r.returncode != 0 and "Could not parse obstacles file" in r.stderr
    r = run("--synthetic", "128x", "--iters", "5")
    assert r.returncode != 0
    # the reference's flag spellings are accepted: --num-ipus, --device (ignored), --exe (a prepared-solver file; one
    # that does not exist yet means "prepare on the fly")
    r = run("--params=" + os.path.join(gold, "input_128x128.params"), "--obstacles=" + os.path.join(gold, "obstacles_128x128.dat"),
            "--num-ipus", "1", "--device", "ipu", "--exe", str(tmp_path / "graph.exe"), "--outdir", str(tmp_path), "--reruns", "0")
    if lbm.device_count() == 0:
        assert r.returncode != 0 and "no CPU path" in r.stderr
    else:
        assert r.returncode == 0 and "==done==" in r.stdout


def test_runner_accepts_the_cxxopts_spellings(lbm, tmp_path):
    """What cxxopts 2.2 (the reference's parser, main/LbmRunner.cpp:14-33) accepts must parse here too: grouped short
    flags (-dn 2), --flag=true/false, --name=value; GNU-style -n2 on top.  Parsing is judged by how far the run gets:
    a parse error prints the usage, a good parse reaches the device check (or runs, on a GPU box)."""
    gold = os.path.join(ROOT, "tests", "golden", "params")
    files = ["--params", os.path.join(gold, "input_128x128.params"), "--obstacles", os.path.join(gold, "obstacles_128x128.dat"),
             "--outdir", str(tmp_path), "--reruns", "0", "--iters", "10"]
    run = lambda *a: subprocess.run([lbm.RUNNER_PATH, *files, *a], capture_output=True, text=True)
    usage = "Runs a Lattice Boltzmann simulation"
    for good in (["-dn", "1"], ["-n1"], ["-d", "-n", "1"], ["--debug=true"], ["--debug=false", "--num-gpus=1"], ["--debug=T"],
                 ["-o", str(tmp_path / "state.lbm")], ["--stats"], ["--no-output=1"]):
        r = run(*good)
        assert usage not in r.stderr, (good, r.stderr)
        if lbm.device_count() == 0:
            assert r.returncode != 0 and "no CPU path" in r.stderr
        else:
            assert r.returncode == 0 and "==done==" in r.stdout, (good, r.stderr)
            assert ("Capturing profile information" in r.stdout) == (good[0] in ("-dn", "-d", "--debug=true", "--debug=T"))
    for bad in (["-nd", "1"], ["--debug=maybe"], ["-n"], ["--num-gpus"], ["-x"], ["stray"], ["--prepare"], ["-n", "two"]):
        r = run(*bad)
        assert r.returncode != 0 and usage in r.stderr, bad


@pytest.mark.gpu
def test_numa_helpers_are_best_effort(lbm):
    """lbm_host_alloc_near / lbm_bind_thread_near: pinned memory near a GPU and the GPU-local CPU set, from sysfs.  On a flat
    or virtualised topology they degrade to a plain pinned allocation / a refusal code -- never a crash."""
    import ctypes
    import numpy as np
    ptr = lbm.host_alloc(1 << 20, 0)
    assert ptr
    a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(1 << 20,))
    a[:] = 7
    assert int(a.sum()) == 7 << 20
    lbm.lib.lbm_host_free(ptr)
    before = os.sched_getaffinity(0)
    try:
        rc = lbm.lib.lbm_bind_thread_near(0)
        assert rc >= 0 or rc == -3  # the NUMA node, or LBM_E_STATE when sysfs has no local_cpulist / the call is refused
        assert lbm.lib.lbm_bind_thread_near(999) < 0
    finally:
        os.sched_setaffinity(0, before)  # the rest of the session keeps all its cores


def test_new_entry_points_validate_before_touching_cuda(lbm, tmp_path):
    """lbm_load, lbm_create_grid, lbm_pipeline_create and lbm_solve_host reject bad input with LBM_E_ARG (-1) and a message
    without needing a device; with good input and no device they fail with LBM_E_CUDA (-2), never silently."""
    import ctypes
    h = ctypes.c_void_p()
    bad = tmp_path / "bad.lbm"
    bad.write_bytes(b"x" * 100)
    assert lbm.lib.lbm_load(ctypes.byref(h), str(bad).encode(), 1) == -1 and b"not a prepared-solver file" in lbm.lib.lbm_last_error()
    assert lbm.lib.lbm_load(ctypes.byref(h), str(tmp_path / "none.lbm").encode(), 1) == -1 and not h.value
    assert lbm.lib.lbm_create_grid(ctypes.byref(h), 64, 64, 1, 0.1, 0.005, 1.85, 0, 2, None) == -1
    assert lbm.lib.lbm_create_grid(ctypes.byref(h), 8, 64, 1, 0.1, 0.005, 1.85, 16, 1, None) == -1 and b"columns" in lbm.lib.lbm_last_error()
    p = ctypes.c_void_p()
    assert lbm.lib.lbm_pipeline_create(ctypes.byref(p), 0, 64, 1, 0.1, 0.005, 1.85, 0) == -1 and not p.value
    assert lbm.lib.lbm_pipeline_solve(None, None, None, None, None) == -1
    lbm.lib.lbm_pipeline_destroy(None)  # harmless
    assert lbm.lib.lbm_solve_host(64, 64, 1, 0.1, 0.005, 1.85, 0, None, None, None, None, None) == -1
    assert lbm.lib.lbm_save(None, b"x") == -1 and lbm.lib.lbm_ipc_disconnect(None) == -1
    if lbm.device_count() == 0:
        assert lbm.lib.lbm_create_grid(ctypes.byref(h), 64, 64, 1, 0.1, 0.005, 1.85, 1, 1, None) == -2
        assert lbm.lib.lbm_pipeline_create(ctypes.byref(p), 64, 64, 1, 0.1, 0.005, 1.85, 0) == -2
        assert b"CUDA" in lbm.lib.lbm_last_error() or b"device" in lbm.lib.lbm_last_error()


def test_pipeline_plan_without_a_gpu(lbm, monkeypatch):
    """lbm_pipeline_plan: the band cut of a whole job / of one rank's window.  The margin covers the steps and ends on a tile
    row; bands are whole tile rows and at least 8 margins tall; jobs too long or grids too small are not banded."""
    monkeypatch.delenv("LBM_B200_BAND_ROWS", raising=False)
    monkeypatch.delenv("LBM_B200_NO_BANDS", raising=False)
    assert lbm.pipeline_plan(16384, 16384, 20) == {"bands": 16, "band_rows": 1024, "margin_rows": 23}
    assert lbm.pipeline_plan(16384, 16384, 200) == {"bands": 1, "band_rows": 16384, "margin_rows": 0}  # 8 x 207 rows > a band
    assert lbm.pipeline_plan(2048, 2048, 20)["bands"] == 1                                            # fewer than four bands
    for nx, ny, iters in ((16384, 16384, 1), (16384, 16384, 7), (8192, 32768, 50), (32768, 32768, 20)):
        plan = lbm.pipeline_plan(nx, ny, iters)
        if plan["bands"] > 1:
            m, b = plan["margin_rows"], plan["band_rows"]
            assert m >= iters and (m + 1) % 8 == 0 and b % 8 == 0 and b >= 8 * m and plan["bands"] == -(-ny // b) >= 4
    # one rank's slab of the 16384^2 grid on 2 / 4 / 8 GPUs: always banded, at least four bands where the margins allow
    assert lbm.pipeline_plan(16384, 16384, 20, window_rows=8192) == {"bands": 8, "band_rows": 1024, "margin_rows": 23}
    assert lbm.pipeline_plan(16384, 16384, 20, window_rows=4096) == {"bands": 4, "band_rows": 1024, "margin_rows": 23}
    assert lbm.pipeline_plan(16384, 16384, 20, window_rows=2048) == {"bands": 4, "band_rows": 512, "margin_rows": 23}
    assert lbm.pipeline_plan(16384, 16384, 100, window_rows=2048) == {"bands": 3, "band_rows": 824, "margin_rows": 103}
    assert lbm.pipeline_plan(1024, 1024, 300, window_rows=128)["bands"] == 1  # a window is one band when nothing else fits
    with pytest.raises(lbm.LbmError, match="multiple of 8"):
        lbm.pipeline_plan(256, 256, 3, window_rows=100)


def test_window_plan_never_holds_the_driven_row_twice(lbm):
    """A band solver drives one row, so band + 2 x margin must fit in the grid (tests/test_pipeline_model.py shows what happens
    otherwise): short grids get shorter bands, and a grid with no room for even one tile row of band is refused."""
    for nx, ny, rows, iters in [(16, 64, 64, 20), (16, 64, 32, 20), (128, 128, 128, 50), (384, 256, 128, 9), (64, 96, 48, 30)]:
        plan = lbm.pipeline_plan(nx, ny, iters, window_rows=rows)
        assert plan["margin_rows"] >= iters and (plan["margin_rows"] + 1) % 8 == 0
        assert plan["band_rows"] % 8 == 0 and plan["band_rows"] >= 8
        assert plan["band_rows"] + 2 * plan["margin_rows"] <= ny, plan
        assert plan["bands"] == -(-rows // plan["band_rows"])
    assert lbm.pipeline_plan(384, 256, 9, window_rows=128) == {"bands": 2, "band_rows": 120, "margin_rows": 15}  # unchanged
    with pytest.raises(lbm.LbmError, match="too short"):
        lbm.pipeline_plan(16, 40, 20, window_rows=40)

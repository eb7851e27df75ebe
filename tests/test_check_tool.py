"""tools/check.py is the reference's acceptance test (check/check.py, Python 2 only) restated for Python 3.
Where /root/reference is mounted the original is translated mechanically (print statements only) and run
side by side on the same files; everywhere, the tool's verdicts are checked on constructed cases."""
import io
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_text, oracle_fixture
from tools import check as check_tool

REF_CHECK = "/root/reference/check/check.py"


def write_case(tmp_path, av, pressure, nx, ny, tag):
    a, f = tmp_path / f"{tag}.av_vels.dat", tmp_path / f"{tag}.final_state.dat"
    with open(a, "w") as fh:
        for i, v in enumerate(av):
            fh.write(f"{i}:\t{v:.12E}\n")
    with open(f, "w") as fh:
        for jj in range(ny):
            for ii in range(nx):
                fh.write(f"{ii} {jj} 0.0 0.0 0.0 {pressure[jj, ii]:.12E} 0\n")
    return str(a), str(f)


def run_tool(args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check.py")] + args, capture_output=True, text=True)


def test_verdicts(tmp_path):
    rng = np.random.default_rng(0)
    av = 0.01 + 0.001 * rng.random(50)
    pr = 0.0333 + 0.0001 * rng.random((6, 5))
    ra, rf = write_case(tmp_path, av, pr, 5, 6, "ref")
    for scale, want in ((1.0, 0), (1.005, 0), (1.02, 1)):
        sa, sf = write_case(tmp_path, av * scale, pr, 5, 6, f"sim{scale}")
        r = run_tool(["--ref-av-vels-file", ra, "--ref-final-state-file", rf, "--av-vels-file", sa, "--final-state-file", sf])
        assert r.returncode == want, r.stdout
        assert ("Both tests passed!" in r.stdout) == (want == 0)
    # tolerance flag, pressure failure, NaN, different step count
    sa, sf = write_case(tmp_path, av * 1.02, pr, 5, 6, "loose")
    assert run_tool(["--tolerance", "5", "--ref-av-vels-file", ra, "--ref-final-state-file", rf, "--av-vels-file", sa,
                     "--final-state-file", sf]).returncode == 0
    sa, sf = write_case(tmp_path, av, pr * 1.03, 5, 6, "press")
    r = run_tool(["--ref-av-vels-file", ra, "--ref-final-state-file", rf, "--av-vels-file", sa, "--final-state-file", sf])
    assert r.returncode == 1 and "final state failed check" in r.stdout and "av_vels failed" not in r.stdout
    bad = av.copy()
    bad[3] = np.nan
    sa, sf = write_case(tmp_path, bad, pr, 5, 6, "nan")
    assert run_tool(["--ref-av-vels-file", ra, "--ref-final-state-file", rf, "--av-vels-file", sa, "--final-state-file", sf]).returncode == 1
    sa, sf = write_case(tmp_path, av[:-1], pr, 5, 6, "short")
    r = run_tool(["--ref-av-vels-file", ra, "--ref-final-state-file", rf, "--av-vels-file", sa, "--final-state-file", sf])
    assert r.returncode == 1 and "Different number of steps" in r.stdout


@pytest.mark.skipif(not os.path.exists(REF_CHECK), reason="the reference tree is not mounted on this machine")
def test_same_numbers_as_the_original_script(tmp_path):
    """Mechanical py3 translation of check/check.py (strip the version gate, print X -> print(X)) vs tools/check.py
    on the reference's 128x128 goldens and the oracle's output for them."""
    src = open(REF_CHECK).read()
    out_lines = []
    gate = False
    for line in src.splitlines():
        if line.startswith("if sys.version_info"):  # the Python-2.7-only gate (check/check.py:5-10)
            gate = True
            continue
        if gate and (line.startswith(" ") or line.startswith("\t")):
            continue
        gate = False
        m = re.match(r"^(\s*)print(?:\s+(.*))?$", line)  # a print STATEMENT (not print_diffs(...))
        if m:
            line = f"{m.group(1)}print({m.group(2) or ''})"
        out_lines.append(line)
    translated = tmp_path / "check_py3.py"
    translated.write_text("\n".join(out_lines) + "\n")
    fx = oracle_fixture("128x128")
    ref_av, ref_fs = golden_text("128x128.av_vels.dat", tmp_path), golden_text("128x128.final_state.dat", tmp_path)
    nx = ny = 128
    sa, sf = write_case(tmp_path, fx["av_f32sum"].astype(np.float64), fx["pressure"].astype(np.float64).reshape(ny, nx), nx, ny, "oracle")
    args = ["--ref-av-vels-file", ref_av, "--ref-final-state-file", ref_fs, "--av-vels-file", sa, "--final-state-file", sf]
    orig = subprocess.run([sys.executable, str(translated)] + args, capture_output=True, text=True)
    if orig.returncode not in (0, 1) or "Traceback" in orig.stderr:
        pytest.skip("mechanical translation of the reference script did not run: " + orig.stderr[-300:])
    ours = run_tool(args)
    assert ours.returncode == orig.returncode == 0
    nums = lambda s: [float(x) for x in re.findall(r"[-+]?\d\.\d+E[-+]\d+", s)]
    assert nums(ours.stdout) == nums(orig.stdout) and len(nums(ours.stdout)) >= 6

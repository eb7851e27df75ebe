#!/usr/bin/env python3
"""check.py for Python 3 -- the reference's acceptance test, restated.

The reference ships check/check.py (thorbenlouw/lbm-graphcore), a Python 2.7-only script
(print statements and an explicit version gate, check/check.py:5-10).  This file keeps its
command line (check/check.py:17-57), its arithmetic (:62-100) and its verdict (:134-147)
and runs under Python 3.  tests/test_check_tool.py translates the original mechanically
(print -> print()) whenever /root/reference is present and asserts both give the same
numbers and exit code on the same files.

    python tools/check.py --ref-av-vels-file REF_AV --ref-final-state-file REF_FS \
                          --av-vels-file AV --final-state-file FS [--tolerance PCT]

Compared: column 1 of av_vels.dat and column 5 (pressure) of final_state.dat, as the
maximum relative deviation in percent, `100 * (ref - sim) / sim`; non-finite or larger
than the tolerance (default 1 %) fails.  Exit code 0 = both passed, 1 = failed.
"""
import argparse
import sys

import numpy as np


def parse(argv=None):
    p = argparse.ArgumentParser(description="Testing script for HPC LBM coursework", fromfile_prefix_chars="@",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--tolerance", nargs=1, default=[1], type=float,
                   help="Percentage tolerance to match against reference results")
    p.add_argument("--ref-av-vels-file", nargs=1, required=True, help="reference av_vels results file")
    p.add_argument("--ref-final-state-file", nargs=1, required=True, help="reference final_state results file")
    p.add_argument("--av-vels-file", nargs=1, required=True, help="calculated av_vels results file")
    p.add_argument("--final-state-file", nargs=1, required=True, help="calculated final_state results file")
    return p.parse_args(argv)


def load_dat_files(av_vels_filename, final_state_filename):
    """check/check.py:62-68: av_vels column 1; final_state columns x, y, pressure."""
    av_vels = np.loadtxt(av_vels_filename, usecols=[1])
    final_state = np.loadtxt(final_state_filename, usecols=[0, 1, 5])
    return av_vels, final_state


def get_diff_values(ref_vals, sim_vals):
    """check/check.py:84-100."""
    with np.errstate(divide="ignore", invalid="ignore"):
        diff = ref_vals - sim_vals
        diff_pcnt = 100.0 * (diff / (ref_vals - diff))
    k = int(np.argmax(np.abs(diff_pcnt)))
    return {"max_diff_step": k, "max_diff": diff[k], "max_diff_pcnt": diff_pcnt[k], "sim_val": sim_vals[k],
            "ref_val": ref_vals[k], "total": float(np.sum(np.abs(diff)))}


def compare(av_ref, fs_ref, av_sim, fs_sim, tolerance=1.0, out=sys.stdout):
    """Returns (exit_code, av_vels_diffs, final_state_diffs) exactly as the reference decides them."""
    if np.any(fs_ref[:, 0:2] != fs_sim[:, 0:2]):  # :75-77
        print("Final state files coordinates were not the same", file=out)
        return 1, None, None
    if av_ref.size != av_sim.size:  # :80-82
        print("Different number of steps in av_vels files", file=out)
        return 1, None, None
    av = get_diff_values(av_ref, av_sim)
    print("Total difference in av_vels : {total:.12E}".format(**av), file=out)
    print("Biggest difference (at step {max_diff_step:d}) : {max_diff:.12E}".format(**av), file=out)
    print("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**av), file=out)
    print(file=out)
    fs = get_diff_values(fs_ref[:, 2], fs_sim[:, 2])
    loc = fs["max_diff_step"]
    fs["jj"], fs["ii"] = int(fs_sim[loc, 0]), int(fs_sim[loc, 1])
    print("Total difference in final_state : {total:.12E}".format(**fs), file=out)
    print("Biggest difference (at coord ({jj:d},{ii:d})) : {max_diff:.12E}".format(**fs), file=out)
    print("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**fs), file=out)
    print(file=out)
    fs_failed = (not np.isfinite(fs["max_diff_pcnt"])) or abs(fs["max_diff_pcnt"]) > tolerance  # :134
    av_failed = (not np.isfinite(av["max_diff_pcnt"])) or abs(av["max_diff_pcnt"]) > tolerance  # :135
    if fs_failed:
        print("final state failed check", file=out)
    if av_failed:
        print("av_vels failed check", file=out)
    if fs_failed or av_failed:
        return 1, av, fs
    print("Both tests passed!", file=out)
    return 0, av, fs


def main(argv=None):
    a = parse(argv)
    av_ref, fs_ref = load_dat_files(a.ref_av_vels_file[0], a.ref_final_state_file[0])
    av_sim, fs_sim = load_dat_files(a.av_vels_file[0], a.final_state_file[0])
    code, _, _ = compare(av_ref, fs_ref, av_sim, fs_sim, a.tolerance[0])
    return code


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Result checker for the d2q9-bgk solver — Python 3.

Same command line, same arithmetic, same messages and exit codes as the reference's
Python-2.7-only checker (ag14774/OpenCL-Lattice-Boltzmann check/check.py:1-147), which
cannot run on an image that has only Python 3:

  * flags --tolerance (percent, default 1), --ref-av-vels-file, --ref-final-state-file,
    --av-vels-file, --final-state-file; ``@file`` argument files (check.py:16-57);
  * av_vels: column 1; final_state: columns 0, 1 (coordinates) and 5 (pressure) (:62-66);
  * coordinates and step counts must agree (:75-82);
  * diff = ref - sim, diff_pcnt = 100 * diff / (ref - diff), worst = argmax |diff_pcnt| (:84-101);
  * fails when the worst percentage of either file is not finite or exceeds the
    tolerance; exit status 1 on failure, 0 with "Both tests passed!" otherwise (:131-147).

The comparison is also importable: ``compare(ref, sim)`` and ``check_files(...)``.
"""
import argparse
import sys

import numpy as np


def build_parser():
    parser = argparse.ArgumentParser(
        description="Testing script for HPC LBM coursework",
        fromfile_prefix_chars="@",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
    )
    parser.add_argument("--tolerance", nargs=1, default=[1], type=float,
                        help="Percentage tolerance to match against reference results")
    parser.add_argument("--ref-av-vels-file", nargs=1, required=True, help="reference av_vels results file")
    parser.add_argument("--ref-final-state-file", nargs=1, required=True,
                        help="reference final_state results file")
    parser.add_argument("--av-vels-file", nargs=1, required=True, help="calculated av_vels results file")
    parser.add_argument("--final-state-file", nargs=1, required=True, help="calculated final_state results file")
    return parser


def load_av_vels(filename):
    """Second column of "step:\\tvalue" lines."""
    with open(filename, "r") as fp:
        return np.atleast_1d(np.loadtxt(fp, usecols=[1]))


def load_final_state(filename):
    """Columns x, y, pressure of "x y u_x u_y u pressure obstacle" lines."""
    with open(filename, "r") as fp:
        return np.atleast_2d(np.loadtxt(fp, usecols=[0, 1, 5]))


def compare(ref_vals, sim_vals):
    """The reference's difference measure; returns a dict with the worst entry."""
    ref_vals = np.asarray(ref_vals, dtype=np.float64)
    sim_vals = np.asarray(sim_vals, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        diff = ref_vals - sim_vals
        diff_pcnt = 100.0 * (diff / (ref_vals - diff))
    worst = int(np.argmax(np.abs(diff_pcnt)))
    return {
        "max_diff_step": worst,
        "max_diff": diff[worst],
        "max_diff_pcnt": diff_pcnt[worst],
        "sim_val": sim_vals[worst],
        "ref_val": ref_vals[worst],
        "total": np.sum(np.abs(diff)),
    }


def failed(diffs, tolerance):
    pct = diffs["max_diff_pcnt"]
    return (not np.isfinite(pct)) or (abs(pct) > tolerance)


def check_files(ref_av_vels_file, ref_final_state_file, av_vels_file, final_state_file, tolerance=1.0, out=sys.stdout):
    """Runs the whole check; returns the process exit status (0 pass, 1 fail)."""
    av_vels_ref = load_av_vels(ref_av_vels_file)
    final_state_ref = load_final_state(ref_final_state_file)
    av_vels_sim = load_av_vels(av_vels_file)
    final_state_sim = load_final_state(final_state_file)

    if final_state_ref.shape != final_state_sim.shape or np.any(final_state_ref[:, 0:2] != final_state_sim[:, 0:2]):
        print("Final state files coordinates were not the same", file=out)
        return 1
    if av_vels_ref.size != av_vels_sim.size:
        print("Different number of steps in av_vels files", file=out)
        return 1

    av = compare(av_vels_ref, av_vels_sim)
    print("Total difference in av_vels : {total:.12E}".format(**av), file=out)
    print("Biggest difference (at step {max_diff_step:d}) : {max_diff:.12E}".format(**av), file=out)
    print("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**av), file=out)
    print(file=out)

    fs = compare(final_state_ref[:, 2], final_state_sim[:, 2])
    where = fs["max_diff_step"]
    fs["jj"] = int(final_state_sim[where, 0])
    fs["ii"] = int(final_state_sim[where, 1])
    print("Total difference in final_state : {total:.12E}".format(**fs), file=out)
    print("Biggest difference (at coord ({jj:d},{ii:d})) : {max_diff:.12E}".format(**fs), file=out)
    print("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**fs), file=out)
    print(file=out)

    final_state_failed = failed(fs, tolerance)
    av_vels_failed = failed(av, tolerance)
    if final_state_failed:
        print("final state failed check", file=out)
    if av_vels_failed:
        print("av_vels failed check", file=out)
    if final_state_failed or av_vels_failed:
        return 1
    print("Both tests passed!", file=out)
    return 0


def main(argv=None):
    args = build_parser().parse_args(argv)
    return check_files(args.ref_av_vels_file[0], args.ref_final_state_file[0], args.av_vels_file[0],
                       args.final_state_file[0], tolerance=args.tolerance[0])


if __name__ == "__main__":
    sys.exit(main())

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

Extension (off by default, so the default run is the reference's check exactly):
  * --velocity-tolerance PCT also compares the u_x, u_y and |u| columns (2, 3, 4) of final_state,
    which the reference never looks at.  Velocities pass through zero, so the reference's
    per-value relative measure is meaningless for them; the measure here is fp32-aware:
    100 * |ref - sim| / max(|ref|, 1e-3 * max|ref u|), i.e. relative to the value, but never to
    less than a thousandth of the field's largest speed.

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
    parser.add_argument("--velocity-tolerance", nargs=1, default=[None], type=float,
                        help="extension: also check the u_x, u_y, |u| columns of final_state against this "
                             "percentage (relative to max(|ref|, 0.1 %% of the largest reference speed))")
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


def load_velocities(filename):
    """Columns u_x, u_y, |u| of "x y u_x u_y u pressure obstacle" lines."""
    with open(filename, "r") as fp:
        return np.atleast_2d(np.loadtxt(fp, usecols=[2, 3, 4]))


def compare_velocities(ref_u, sim_u):
    """fp32-aware measure for the velocity columns: percent of max(|ref|, 1e-3 * max|ref speed|)."""
    ref_u = np.asarray(ref_u, dtype=np.float64)
    sim_u = np.asarray(sim_u, dtype=np.float64)
    floor = 1e-3 * float(np.max(np.abs(ref_u[:, 2]))) if ref_u.size else 0.0
    diff = ref_u - sim_u
    with np.errstate(divide="ignore", invalid="ignore"):
        pcnt = 100.0 * np.abs(diff) / np.maximum(np.abs(ref_u), floor)
    pcnt = np.where(np.isnan(pcnt) & (diff == 0.0), 0.0, pcnt)   # 0/0: identical zero fields
    flat = int(np.argmax(np.where(np.isfinite(pcnt), pcnt, np.inf)))
    row, col = divmod(flat, 3)
    return {"row": row, "column": ("u_x", "u_y", "u")[col], "max_diff": diff[row, col], "max_diff_pcnt": pcnt[row, col],
            "sim_val": sim_u[row, col], "ref_val": ref_u[row, col], "total": float(np.sum(np.abs(diff)))}


def failed(diffs, tolerance):
    pct = diffs["max_diff_pcnt"]
    return (not np.isfinite(pct)) or (abs(pct) > tolerance)


def check_files(ref_av_vels_file, ref_final_state_file, av_vels_file, final_state_file, tolerance=1.0, out=sys.stdout,
                velocity_tolerance=None):
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

    velocity_failed = False
    if velocity_tolerance is not None:   # extension: the columns the reference ignores
        uv = compare_velocities(load_velocities(ref_final_state_file), load_velocities(final_state_file))
        uv["jj"] = int(final_state_sim[uv["row"], 0])
        uv["ii"] = int(final_state_sim[uv["row"], 1])
        print("Total difference in final_state velocities : {total:.12E}".format(**uv), file=out)
        print("Biggest difference ({column} at coord ({jj:d},{ii:d})) : {max_diff:.12E}".format(**uv), file=out)
        print("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**uv), file=out)
        print(file=out)
        velocity_failed = failed(uv, velocity_tolerance)

    final_state_failed = failed(fs, tolerance)
    av_vels_failed = failed(av, tolerance)
    if final_state_failed:
        print("final state failed check", file=out)
    if av_vels_failed:
        print("av_vels failed check", file=out)
    if velocity_failed:
        print("final state velocities failed check", file=out)
    if final_state_failed or av_vels_failed or velocity_failed:
        return 1
    print("Both tests passed!", file=out)
    return 0


def main(argv=None):
    args = build_parser().parse_args(argv)
    return check_files(args.ref_av_vels_file[0], args.ref_final_state_file[0], args.av_vels_file[0],
                       args.final_state_file[0], tolerance=args.tolerance[0],
                       velocity_tolerance=args.velocity_tolerance[0])


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Regenerates the two final_state goldens the reference repository stripped as large blobs
(/root/reference/.MISSING_LARGE_BLOBS: check/256x256.final_state.dat, check/1024x1024.final_state.dat).

They are produced by the fp64 oracle (oracle/lbm_oracle.c, the original serial equations), which
reproduces every golden the reference DOES ship bit for digit (tests/test_oracle_goldens.py: av_vels to
5e-11 %, final-state pressure identical to all 12 printed digits) — including the av_vels goldens of
these two decks, which the run below re-checks before writing anything.

    python check/regenerate_missing_goldens.py 256x256      # -> check/256x256.final_state.dat (text, committed)
    python check/regenerate_missing_goldens.py 1024x1024    # -> check/1024x1024.final_state.pressure.npz (committed)

The 1024x1024 text file would be 90 MB, so only its checked column (pressure, float64) is committed;
`python check/regenerate_missing_goldens.py --expand 1024x1024` writes check/1024x1024.final_state.dat
(git-ignored) from it, with the unchecked velocity columns as 0.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CHECK = os.path.join(ROOT, "check")


def write_text(path, nx, ny, u_x, u_y, u, pressure, obstacles):
    with open(path, "w") as fp:
        for ii in range(ny):
            rows = ["%d %d %.12E %.12E %.12E %.12E %d\n" % (jj, ii, u_x[ii, jj], u_y[ii, jj], u[ii, jj],
                                                           pressure[ii, jj], obstacles[ii, jj]) for jj in range(nx)]
            fp.write("".join(rows))


def expand(name):
    from opencl_lattice_boltzmann_b200 import decks
    p, _, obstacles = decks.load_deck(*decks.deck_paths(name))
    pressure = np.load(os.path.join(CHECK, f"{name}.final_state.pressure.npz"))["pressure"]
    zero = np.zeros_like(pressure)
    write_text(os.path.join(CHECK, f"{name}.final_state.dat"), p.nx, p.ny, zero, zero, zero, pressure, obstacles)


def regenerate(name):
    import oracle_lib
    from helpers import pct_diff
    from opencl_lattice_boltzmann_b200 import decks
    p, _, obstacles = decks.load_deck(*decks.deck_paths(name))
    toks = open(decks.deck_paths(name)[0]).read().split()
    p64 = decks.Params(nx=p.nx, ny=p.ny, maxIters=p.maxIters, reynolds_dim=p.reynolds_dim,
                       density=float(toks[4]), accel=float(toks[5]), omega=float(toks[6]))
    c64 = np.empty((9, p.ny, p.nx), dtype=np.float64)
    c64[0] = p64.density * 4.0 / 9.0
    c64[1:5] = p64.density / 9.0
    c64[5:9] = p64.density / 36.0
    t0 = time.time()
    cells, av, _ = oracle_lib.run_f64(p64, c64, obstacles, p.maxIters)
    ref_av = np.loadtxt(os.path.join(CHECK, f"{name}.av_vels.dat"), usecols=[1])
    worst, step = pct_diff(ref_av, av)
    print(f"{name}: {p.maxIters} fp64 steps in {time.time() - t0:.0f}s; av_vels vs shipped golden: worst {worst:.3g}% at step {step}")
    if not abs(worst) < 1e-8:
        raise SystemExit("fp64 oracle does not reproduce the shipped av_vels golden; not writing anything")
    u_x, u_y, u, pressure = oracle_lib.final_state_f64(p64, cells, obstacles)
    if p.nx * p.ny <= 256 * 256:
        write_text(os.path.join(CHECK, f"{name}.final_state.dat"), p.nx, p.ny, u_x, u_y, u, pressure, obstacles)
    else:
        # keep the 12 printed digits only, as the text golden would
        pr = np.array([float("%.12E" % v) for v in pressure.ravel()]).reshape(pressure.shape)
        np.savez_compressed(os.path.join(CHECK, f"{name}.final_state.pressure.npz"), pressure=pr)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--expand":
        expand(sys.argv[2])
    elif len(sys.argv) == 2:
        regenerate(sys.argv[1])
    else:
        raise SystemExit(__doc__)

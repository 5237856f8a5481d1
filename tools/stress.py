"""Stress loop for rare races: repeats small persistent / per-step runs against the oracle and
prints where the lattice differs when it does."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    persistent = int(sys.argv[2]) if len(sys.argv) > 2 else 1      # -1: the automatic choice (tile kernel on these shapes)
    shapes = [(256, 64), (128, 128), (384, 24), (512, 96), (1024, 64), (384, 384), (200, 333)]
    refs = {}
    bad = 0
    for it in range(iters):
        nx, ny = shapes[it % len(shapes)]
        nsteps = 7 + (it % 3) + (40 if persistent < 0 else 0)       # several rounds of the tile kernel
        key = (nx, ny, nsteps)
        p, cells, obstacles = helpers.random_case(nx, ny, seed=nx * 1000 + ny)
        if key not in refs:
            refs[key] = oracle_lib.run_f32(p, cells, obstacles, nsteps)
        ref_cells, ref_av = refs[key]
        with lbm.cabi.Simulation(p, options={"persistent": persistent}) as sim:
            sim.upload(cells, obstacles)
            sim.run(nsteps)
            sim.sync()
            got, av, info = sim.download_cells(), sim.download_av_vels(nsteps), sim.info()
        diff = helpers.bits(got) != helpers.bits(ref_cells)
        av_bad = not np.allclose(av, ref_av, rtol=2e-6, atol=0)
        if diff.any() or av_bad:
            bad += 1
            k, y, x = np.nonzero(diff)
            print(f"iter {it} {nx}x{ny} steps {nsteps} {info['kernel_name']}: {diff.sum()} cells differ; planes {sorted(set(k.tolist()))} "
                  f"rows {sorted(set(y.tolist()))[:12]} x-range {x.min() if x.size else None}..{x.max() if x.size else None} av_bad={av_bad}",
                  flush=True)
            if av_bad:
                print("   av got", av, "\n   av ref", ref_av, flush=True)
    print(f"stress done: {bad} bad of {iters}", flush=True)


if __name__ == "__main__":
    main()

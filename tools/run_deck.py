"""Runs a reference deck (or a prefix of it) through the C-ABI once — the command profiled by ncu for
the L2-resident grids (profiles/*_deck_*).  Prints device time, us/step, MLUPS."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("deck")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--persistent", type=int, default=-1)
    args = ap.parse_args()
    p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(args.deck))
    n = args.steps or p.maxIters
    opts = {"persistent": args.persistent}
    if args.chunk:
        opts["chunk_steps"] = args.chunk
    with lbm.cabi.Simulation(p, options=opts) as sim:
        sim.upload(cells, obstacles)
        ms = sim.run_timed(n)
        info = sim.info()
    print(f"{args.deck} {n} steps {info['kernel_name']}: {ms:.3f} ms, {ms * 1e3 / n:.3f} us/step, "
          f"{p.nx * p.ny * n / ms / 1e3:.0f} MLUPS")


if __name__ == "__main__":
    main()

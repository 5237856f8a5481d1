"""Where a round of tile_kernel spends its time (development tool): SM clock stamps of tile 0 / thread 0 over
64 rounds (option tile_debug), averaged per phase.

    python tools/tile_timing.py [deck] [--variants k1,k2,k4,k4:16x8,...] [--steps N]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("deck", nargs="?", default="128x128")
    ap.add_argument("--variants", default="auto,k1,k2,k4,k6,k4:16x8")
    ap.add_argument("--steps", type=int, default=2000)
    args = ap.parse_args()
    p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(args.deck))
    for var in args.variants.split(","):
        opts = {"tile": 1, "tile_debug": 1}
        for part in ([] if var == "auto" else var.split(":")):
            if part[0] == "k":
                opts["tile_steps"] = int(part[1:])
            else:
                w, h = part.split("x")
                opts.update({"tile_w": int(w), "tile_h": int(h)})
        with lbm.cabi.Simulation(p, options=opts) as sim:
            sim.upload(cells, obstacles)
            sim.run(args.steps)
            sim.sync()
            ms = sim.run_timed(args.steps)
            t = sim.tile_timing().astype(np.float64)
            info = sim.info()
        k = int(info["kernel_name"].split("K=")[1].split(",")[0])
        ok = t[:, 0] > 0
        t = t[ok]
        if len(t) < 4:
            print(var, "no timing rows")
            continue
        ghz = 1.965
        rnd = np.diff(t[:, 0]).mean() / ghz / 1e3
        poll = (t[:, 1] - t[:, 0]).mean() / ghz / 1e3
        hbar = (t[:, 2] - t[:, 1]).mean() / ghz / 1e3
        steps = [(t[:, 3 + i] - t[:, 2 + i]).mean() / ghz / 1e3 for i in range(min(k, 12))]
        own_last = (t[:, 15] - t[:, 2 + k - 1]).mean() / ghz / 1e3
        print(f"{args.deck} {var:10s} {info['kernel_name']:42s} {ms * 1e3 / args.steps:6.3f} us/step | round {rnd:6.3f} us = "
              f"flag wait {poll:5.3f} + halo load {hbar:5.3f} + steps [{', '.join(f'{s:5.3f}' for s in steps)}] "
              f"(own part of last step {own_last:5.3f})", flush=True)


if __name__ == "__main__":
    main()

"""Times the four reference decks at full length through the C-ABI (device loop time), for the
BASELINE.md table: persistent vs per-step-launch kernels, cells per thread, block size."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402
import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decks", default="128x128,128x256,256x256,1024x1024")
    ap.add_argument("--variants", default="p1:v4:t256,p1:v2:t256,p1:v1:t256,p1:v4:t128,p1:v1:t128,p0:v4:t256")
    args = ap.parse_args()
    for name in args.decks.split(","):
        p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(name))
        gold = helpers.golden_av_vels(name)
        for var in args.variants.split(","):
            pv, vv, tv = var.split(":")[:3]
            opts = {"persistent": int(pv[1:]), "cells_per_thread": int(vv[1:]), "threads_per_block": int(tv[1:])}
            for extra in var.split(":")[3:]:
                key = {"g": "global_barrier", "k": "packed", "s": "threads_per_sm"}[extra[0]]
                opts[key] = int(extra[1:])
            with lbm.cabi.Simulation(p, options=opts) as sim:
                sim.upload(cells, obstacles)
                sim.run(min(1000, p.maxIters))      # warm-up
                sim.sync()
                sim.upload(cells, obstacles)
                ms = sim.run_timed(p.maxIters)
                av = sim.download_av_vels(p.maxIters)
                info = sim.info()
            worst, step = helpers.pct_diff(gold, av)
            mlups = p.nx * p.ny * p.maxIters / (ms * 1e-3) / 1e6
            print(f"{name:10s} {var:14s} {info['kernel_name']:34s} {ms/1e3:8.4f} s  {ms*1e3/p.maxIters:7.3f} us/step  "
                  f"{mlups:10.0f} MLUPS  {mlups*72/1e3:8.1f} GB/s  av_vels worst {worst:+.3f}%", flush=True)


if __name__ == "__main__":
    main()

"""Kernel-variant sweep on one GPU (development tool; not the bench).  Prints MLUPS and the
fraction of the measured HBM copy bandwidth for each (cells_per_thread, threads_per_block, streaming)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variants", default="4:256:0,4:256:1,4:256:2,4:256:3,4:256:4,4:128:0,4:512:0,4:128:2,2:256:0,1:256:0")
    args = ap.parse_args()
    peak = 6546.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    t0 = time.time()
    p, cells, obstacles = lbm.decks.synthetic_channel(args.nx, args.ny)
    print(f"deck built in {time.time()-t0:.1f}s", flush=True)
    sim = lbm.cabi.Simulation(p)
    t0 = time.time()
    sim.upload(cells, obstacles)
    print(f"upload {time.time()-t0:.2f}s", flush=True)
    for var in args.variants.split(","):
        if var.startswith("f2"):          # f2:<warps>:<packed>:<seg_rows>[:<prefetch>[:<kernel 0|1|2>[:<l2_ahead>[:<mode>[:<long_rows>]]]]]
            _, w, pk, sr, *rest = var.split(":")
            pf = int(rest[0]) if rest else 1
            sim.set_option("fuse2_tma", int(rest[1]) if len(rest) > 1 else 1)
            sim.set_option("fuse2_l2_ahead", int(rest[2]) if len(rest) > 2 else 0)
            sim.set_option("fuse2_mode", int(rest[3]) if len(rest) > 3 else 1)
            sim.set_option("fuse2_long", int(rest[4]) if len(rest) > 4 else 0)     # 0: uniform segments
            for k, v in (("persistent", 0), ("cells_per_thread", 4), ("fuse2", 1), ("fuse2_warps", int(w)),
                         ("packed", int(pk)), ("fuse2_rows", int(sr)), ("fuse2_prefetch", pf)):
                sim.set_option(k, v)
            sim.run(args.warmup + (args.warmup & 1))
            sim.sync()
            ms = sim.run_timed(args.steps)
            mlups = args.nx * args.ny * args.steps / (ms * 1e-3) / 1e6
            print(f"{var} {sim.info()['kernel_name']} prefetch={pf} l2_ahead={int(rest[2]) if len(rest) > 2 else 0}: {ms/args.steps:.4f} ms/step  {mlups:,.0f} MLUPS  "
                  f"{mlups*72/1e3:,.0f} GB/s algorithmic  {mlups*72/1e3/peak*100:.1f}% of measured HBM copy", flush=True)
            sim.set_option("fuse2", 0)
            continue
        parts = [int(x) for x in var.split(":")]
        V, tpb, st = parts[:3]
        sim.set_option("persistent", parts[3] if len(parts) > 3 else 0)
        sim.set_option("threads_per_sm", parts[4] if len(parts) > 4 else 1024)
        sim.set_option("packed", parts[5] if len(parts) > 5 else 0)
        sim.set_option("cells_per_thread", V)
        sim.set_option("threads_per_block", tpb)
        sim.set_option("streaming", st)
        sim.run(args.warmup)
        sim.sync()
        ms = sim.run_timed(args.steps)
        mlups = args.nx * args.ny * args.steps / (ms * 1e-3) / 1e6
        print(f"{sim.info()['kernel_name']} V={V} tpb={tpb} hint={st} tps={parts[4] if len(parts) > 4 else 1024}: {ms/args.steps:.4f} ms/step  {mlups:,.0f} MLUPS  "
              f"{mlups*72/1e3:,.0f} GB/s  {mlups*72/1e3/peak*100:.1f}% of measured HBM copy", flush=True)
    print(sim.info())
    sim.close()


if __name__ == "__main__":
    main()

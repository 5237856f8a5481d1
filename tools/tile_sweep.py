"""Small decks: us/step of tile_kernel over steps-per-hand-off and tile shapes (development tool).

    python tools/tile_sweep.py [128x128,128x256,256x256] [--steps N]

Every run is checked against the reference's golden av_vels with check.py's measure."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402
import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("decks", nargs="?", default="128x128,128x256,256x256")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--variants", default="auto,k1,k2,k3,k4,k6,k8,k4:16x8,k4:32x4,k4:128x1,k8:16x16,k6:16x8,p1")
    args = ap.parse_args()
    for name in args.decks.split(","):
        p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(name))
        n = args.steps or p.maxIters
        gold = helpers.golden_av_vels(name)[:n]
        for var in args.variants.split(","):
            opts = {}
            if var == "p1":
                opts = {"persistent": 1}
            elif var != "auto":
                for part in var.split(":"):
                    if part[0] == "k":
                        opts["tile_steps"] = int(part[1:])
                    else:
                        w, h = part.split("x")
                        opts.update({"tile_w": int(w), "tile_h": int(h)})
                opts["tile"] = 1
            try:
                with lbm.cabi.Simulation(p, options=opts) as sim:
                    sim.upload(cells, obstacles)
                    sim.run(min(2000, n))
                    sim.sync()
                    sim.upload(cells, obstacles)
                    ms = min(sim.run_timed(n) for _ in range(1))
                    av = sim.download_av_vels(n)
                    info = sim.info()
                worst, step = helpers.pct_diff(gold, av)
                print(f"{name:10s} {var:12s} {info['kernel_name']:44s} {ms * 1e3 / n:7.3f} us/step "
                      f"{p.nx * p.ny * n / ms / 1e3:9.0f} MLUPS  av worst {worst:+.3f}%", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{name:10s} {var:12s} failed: {e}", flush=True)


if __name__ == "__main__":
    main()

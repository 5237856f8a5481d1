"""One deck, several option sets: us/step, MLUPS and check.py's av_vels measure for each (development tool).

    python tools/deck_variants.py 1024x1024 "" "persistent=0,fuse2=1,fuse2_rows=4" "persistent=0,fuse2=0" [--steps N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402
import helpers  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--steps")]
steps_arg = [a for a in sys.argv[1:] if a.startswith("--steps=")]
name, variants = args[0], args[1:] or [""]
p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(name))
n = int(steps_arg[0].split("=")[1]) if steps_arg else p.maxIters
gold = helpers.golden_av_vels(name)[:n]
for var in variants:
    opts = {k: int(v) for k, v in (kv.split("=") for kv in var.split(",") if kv)}
    try:
        with lbm.cabi.Simulation(p, options=opts) as sim:
            sim.upload(cells, obstacles)
            sim.run(min(2000, n))
            sim.sync()
            sim.upload(cells, obstacles)
            ms = sim.run_timed(n)
            av = sim.download_av_vels(n)
            info = sim.info()
        worst, _ = helpers.pct_diff(gold, av)
        print(f"{name:10s} {var or '(default)':44s} {info['kernel_name']:44s} {ms * 1e3 / n:8.3f} us/step "
              f"{p.nx * p.ny * n / ms / 1e3:9.0f} MLUPS  av worst {worst:+.3f}%", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{name:10s} {var:44s} failed: {e}", flush=True)

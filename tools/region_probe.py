"""Times each call of the reference's timed region (upload, loop, sync, downloads) for one deck (development tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "1024x1024"
p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(name))
n = p.maxIters
with lbm.cabi.Simulation(p, devices=[0]) as sim:
    for rep in range(3):
        t = [time.perf_counter()]
        sim.upload(cells, obstacles); t.append(time.perf_counter())
        sim.run(n); t.append(time.perf_counter())
        sim.sync(); t.append(time.perf_counter())
        sim.download_cells(); t.append(time.perf_counter())
        sim.download_av_vels(n); t.append(time.perf_counter())
        names = ["upload", "run (enqueue)", "sync", "download_cells", "download_av_vels"]
        print(name, sim.info()["kernel_name"], " ".join(f"{a}={b - c:.4f}s" for a, b, c in zip(names, t[1:], t[:-1])),
              f"total={t[-1] - t[0]:.4f}s", flush=True)

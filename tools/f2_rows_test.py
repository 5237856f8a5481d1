import os, sys
sys.path.insert(0, "/root/repo")
import opencl_lattice_boltzmann_b200 as lbm
for n in (4096, 6144, 8192, 16384):
    p, cells, obstacles = lbm.decks.synthetic_channel(n, n)
    steps = max(20, min(2000, int(4e9 / (n * n)) // 2 * 2))
    out = []
    for rows in (32, 64, 128, 256):
        with lbm.cabi.Simulation(p, options={"fuse2": 1, "persistent": 0, "fuse2_rows": rows}) as sim:
            sim.upload(cells, obstacles)
            sim.run(steps // 4 * 2 + 2); sim.sync()
            ms = sim.run_timed(steps)
        out.append(f"rows={rows}: {n*n*steps/ms/1e3:8.0f}")
    print(n, " | ".join(out), flush=True)

"""Two-step kernel: MLUPS over segment lengths for a few slab shapes (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402

shapes = [(16384, 2048), (16384, 4096), (16384, 8192), (8192, 8192), (4096, 4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1].split(",")]
for nx, ny in shapes:
    p, cells, obstacles = lbm.decks.synthetic_channel(nx, ny)
    steps = max(20, min(2000, int(4e9 / (nx * ny)) // 2 * 2))
    out = []
    for rows in (16, 32, 64, 128):
        with lbm.cabi.Simulation(p, options={"fuse2": 1, "persistent": 0, "fuse2_rows": rows}) as sim:
            sim.upload(cells, obstacles)
            sim.run(steps // 4 * 2 + 2)
            sim.sync()
            ms = sim.run_timed(steps)
        out.append(f"rows={rows}: {nx * ny * steps / ms / 1e3:8.0f}")
    with lbm.cabi.Simulation(p, options={"fuse2": 0, "persistent": 0}) as sim:
        sim.upload(cells, obstacles)
        sim.run(steps // 4 * 2 + 2)
        sim.sync()
        ms = sim.run_timed(steps)
    print(f"{nx}x{ny}", " | ".join(out), f"| one-step: {nx * ny * steps / ms / 1e3:8.0f}", flush=True)

"""Two-step kernel: MLUPS over row-segment tilings for a few slab shapes (development tool).

    python tools/f2_rows_sweep.py [16384x2048,16384x4096] [uniform rows ...] [long/short ...]

Each shape is run with the automatic tiling, with every uniform segment length and every long/short
pair given (default: 16 32 64 128 and 64/16 64/32 128/32 96/32), and with the one-step kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402

shapes = [(16384, 2048), (16384, 4096), (16384, 8192), (16384, 16384)]
tilings = ["auto", "16", "32", "64", "128", "64/16", "64/32", "128/32", "96/32", "128/16"]
args = sys.argv[1:]
if args and "x" in args[0]:
    shapes = [tuple(int(v) for v in s.split("x")) for s in args[0].split(",")]
    args = args[1:]
if args:
    tilings = args


def timed(p, cells, obstacles, steps, opts):
    with lbm.cabi.Simulation(p, options=opts) as sim:
        sim.upload(cells, obstacles)
        sim.run(steps // 4 * 2 + 2)
        sim.sync()
        best = min(sim.run_timed(steps) for _ in range(3))
        name = sim.info()["kernel_name"]
    return p.nx * p.ny * steps / best / 1e3, name


for nx, ny in shapes:
    p, cells, obstacles = lbm.decks.synthetic_channel(nx, ny)
    steps = max(20, min(2000, int(6e9 / (nx * ny)) // 2 * 2))
    out = []
    for t in tilings:
        opts = {"fuse2": 1, "persistent": 0}
        for kv in filter(None, os.environ.get("F2_OPTS", "").split(",")):   # e.g. F2_OPTS=fuse2_st_cs=1
            key, val = kv.split("=")
            opts[key] = int(val)
        if t != "auto":
            if "/" in t:     # long/short[/number of long segments per strip]
                parts = [int(v) for v in t.split("/")]
                opts.update({"fuse2_long": parts[0], "fuse2_rows": parts[1]})
                if len(parts) > 2:
                    opts["fuse2_nlong"] = parts[2]
            else:
                opts.update({"fuse2_rows": int(t), "fuse2_long": 0})
        try:
            mlups, name = timed(p, cells, obstacles, steps, opts)
            out.append(f"{t}: {mlups:8.0f}" + (f" [{name.split('rows=')[1].rstrip('>')}]" if t == "auto" else ""))
        except Exception as e:  # noqa: BLE001
            out.append(f"{t}: failed ({e})")
    one, _ = timed(p, cells, obstacles, steps, {"fuse2": 0, "persistent": 0})
    print(f"{nx}x{ny} ({steps} steps)", " | ".join(out), f"| one-step: {one:8.0f}", flush=True)

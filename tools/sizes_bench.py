"""MLUPS over grid sizes (square synthetic channels) with the library's automatic kernel choice and
with the one-step kernel forced — shows where the two-step / persistent kernels pay."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencl_lattice_boltzmann_b200 as lbm  # noqa: E402


EXTRA = {k: int(v) for k, v in (kv.split("=") for kv in filter(None, os.environ.get("LBM_OPTS", "").split(",")))}


def run(n, opts, steps):
    opts = {**EXTRA, **opts}      # e.g. LBM_OPTS=fuse2_tma=2 for the A/B predecessor of the two-step kernel
    p, cells, obstacles = lbm.decks.synthetic_channel(n, n)
    with lbm.cabi.Simulation(p, options=opts) as sim:
        sim.upload(cells, obstacles)
        sim.run(steps // 4 * 2 + 2)
        sim.sync()
        ms = sim.run_timed(steps)
        info = sim.info()
    return n * n * steps / ms / 1e3, info["kernel_name"]


def main():
    for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1536,2048,3072,4096,6144,8192,12288").split(",")]:
        steps = max(20, min(2000, int(4e9 / (n * n)) // 2 * 2))
        a, ka = run(n, {}, steps)
        b, kb = run(n, {"fuse2": 0, "persistent": 0}, steps)
        c, kc = run(n, {"fuse2": 1, "persistent": 0}, steps)
        line = f"{n:6d}^2 steps {steps:5d}: auto {a:9.0f} MLUPS [{ka}] | one-step {b:9.0f} | two-step {c:9.0f} [{kc}]"
        if 2.0 * 36.0 * n * n <= 200e6:      # small enough to try the persistent kernel too
            d, kd = run(n, {"persistent": 1}, steps)
            line += f" | persistent {d:9.0f}"
        print(line, flush=True)


if __name__ == "__main__":
    main()

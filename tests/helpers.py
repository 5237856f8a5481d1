"""Shared test helpers: seeded synthetic states and the reference checker's difference measure."""
import os

import numpy as np

from opencl_lattice_boltzmann_b200 import decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_DIR = os.path.join(ROOT, "check")


def random_case(nx, ny, seed=12345, obstacle_frac=0.06, walls=True, accel=0.005, omega=1.85, density=0.1,
                perturb=0.2):
    """A perturbed lattice with random obstacles: exercises every branch (rebound, accelerate mask,
    x/y wrap with open edges when walls=False)."""
    rng = np.random.default_rng(seed)
    p = decks.Params(nx=nx, ny=ny, maxIters=0, reynolds_dim=10, density=float(np.float32(density)),
                     accel=float(np.float32(accel)), omega=float(np.float32(omega)))
    obstacles = (rng.random((ny, nx)) < obstacle_frac).astype(np.int32)
    if walls:
        obstacles[0, :] = 1
        obstacles[-1, :] = 1
    if ny >= 2:
        obstacles[ny - 2, : max(1, nx // 2)] = 0  # keep part of the accelerate row fluid
    free = nx * ny - int(obstacles.sum())
    p.free_cells_inv = float(np.float32(1.0) / np.float32(max(free, 1)))
    cells = decks.initial_cells(p)
    cells = (cells * (1.0 + perturb * (rng.random(cells.shape) - 0.5))).astype(np.float32)
    return p, cells, obstacles


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def pct_diff(ref, sim):
    """check.py's measure: worst 100*(ref-sim)/sim (check/check.py compare())."""
    ref = np.asarray(ref, dtype=np.float64)
    sim = np.asarray(sim, dtype=np.float64)
    d = ref - sim
    with np.errstate(divide="ignore", invalid="ignore"):
        q = 100.0 * d / (ref - d)
    i = int(np.argmax(np.abs(q)))
    return float(q[i]), i


def golden_av_vels(name):
    return np.loadtxt(os.path.join(CHECK_DIR, f"{name}.av_vels.dat"), usecols=[1])


def golden_pressure(name):
    """Final-state pressure golden: the reference's text file, or for 1024x1024 (a 90 MB text file
    upstream stripped) the committed fp64 pressure column (check/regenerate_missing_goldens.py)."""
    npz = os.path.join(CHECK_DIR, f"{name}.final_state.pressure.npz")
    if os.path.exists(npz):
        return np.load(npz)["pressure"].ravel()
    path = os.path.join(CHECK_DIR, f"{name}.final_state.dat")
    return np.loadtxt(path, usecols=[5]) if os.path.exists(path) else None

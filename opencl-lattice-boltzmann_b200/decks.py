"""Deck I/O: the reference's input/output file formats, restated for the Python tooling.

The product's host is the C program under ``host/`` (the reference's language); this
module is the same surface for ``tests/`` and ``bench.py``: ``.params`` / obstacle-list
parsing with the reference's validation and messages (d2q9-bgk.c:457-591), the uniform
initial state (d2q9-bgk.c:529-550), the two output writers (d2q9-bgk.c:772-856) and the
synthetic channel deck of SURVEY.md §8(d).  Pure numpy, no device code.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

NSPEEDS = 9
FINALSTATEFILE = "final_state.dat"   # d2q9-bgk.c:69
AVVELSFILE = "av_vels.dat"           # d2q9-bgk.c:70


class DeckError(Exception):
    """Raised where the reference calls die() (d2q9-bgk.c:868-874)."""


@dataclasses.dataclass
class Params:
    """t_param, d2q9-bgk.c:81-92 (floats are fp32 as in the reference)."""

    nx: int
    ny: int
    maxIters: int
    reynolds_dim: int
    density: float
    accel: float
    omega: float
    free_cells_inv: float = 0.0

    def f32(self, name: str) -> np.float32:
        return np.float32(getattr(self, name))


def read_params(paramfile: str) -> Params:
    """d2q9-bgk.c:457-495: seven whitespace-separated values, four ints then three floats."""
    try:
        with open(paramfile, "r") as fp:
            tokens = fp.read().split()
    except OSError:
        raise DeckError(f"could not open input parameter file: {paramfile}")
    names = ["nx", "ny", "maxIters", "reynolds_dim", "density", "accel", "omega"]
    values = {}
    for i, name in enumerate(names):
        try:
            tok = tokens[i]
            values[name] = int(tok) if i < 4 else float(np.float32(tok))
        except (IndexError, ValueError):
            raise DeckError(f"could not read param file: {name}")
    return Params(**values)


def read_obstacles(obstaclefile: str, params: Params) -> np.ndarray:
    """d2q9-bgk.c:553-591: lines ``x y 1``; sets params.free_cells_inv (duplicate-safe count)."""
    nx, ny = params.nx, params.ny
    obstacles = np.zeros((ny, nx), dtype=np.int32)
    try:
        with open(obstaclefile, "r") as fp:
            tokens = fp.read().split()
    except OSError:
        raise DeckError(f"could not open input obstacles file: {obstaclefile}")
    if len(tokens) % 3 != 0:
        raise DeckError("expected 3 values per line in obstacle file")
    try:
        trip = np.array(tokens, dtype=np.int64).reshape(-1, 3)
    except ValueError:
        raise DeckError("expected 3 values per line in obstacle file")
    if trip.size:
        if np.any(trip[:, 0] < 0) or np.any(trip[:, 0] > nx - 1):
            raise DeckError("obstacle x-coord out of range")
        if np.any(trip[:, 1] < 0) or np.any(trip[:, 1] > ny - 1):
            raise DeckError("obstacle y-coord out of range")
        if np.any(trip[:, 2] != 1):
            raise DeckError("obstacle blocked value should be 1")
        obstacles[trip[:, 1], trip[:, 0]] = 1
    free_cells = nx * ny - int(obstacles.sum())
    params.free_cells_inv = float(np.float32(1.0) / np.float32(free_cells))  # d2q9-bgk.c:591
    return obstacles


def initial_cells(params: Params) -> np.ndarray:
    """d2q9-bgk.c:529-550: uniform w0/w1/w2 everywhere, obstacle cells included; SoA [9, ny, nx]."""
    d = np.float32(params.density)
    w0 = d * np.float32(4.0) / np.float32(9.0)
    w1 = d / np.float32(9.0)
    w2 = d / np.float32(36.0)
    cells = np.empty((NSPEEDS, params.ny, params.nx), dtype=np.float32)
    cells[0] = w0
    cells[1:5] = w1
    cells[5:9] = w2
    return cells


def load_deck(paramfile: str, obstaclefile: str):
    """The file half of initialise() (d2q9-bgk.c:444-597): (params, cells, obstacles)."""
    params = read_params(paramfile)
    obstacles = read_obstacles(obstaclefile, params)
    return params, initial_cells(params), obstacles


def deck_paths(name: str, root: str | None = None):
    """('128x128') -> (decks/input_128x128.params, decks/obstacles_128x128.dat)."""
    root = root or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "decks")
    return os.path.join(root, f"input_{name}.params"), os.path.join(root, f"obstacles_{name}.dat")


def final_state_fields(params: Params, cells: np.ndarray, obstacles: np.ndarray):
    """d2q9-bgk.c:789-831 in fp32: (u_x, u_y, u, pressure), each [ny, nx]."""
    c_sq = np.float32(1.0) / np.float32(3.0)
    cells = cells.reshape(NSPEEDS, params.ny, params.nx).astype(np.float32, copy=False)
    blocked = obstacles.reshape(params.ny, params.nx) != 0
    local_density = np.zeros((params.ny, params.nx), dtype=np.float32)
    for kk in range(NSPEEDS):
        local_density = local_density + cells[kk]
    with np.errstate(divide="ignore", invalid="ignore"):
        u_x = (cells[1] + cells[5] + cells[8] - cells[3] - cells[6] - cells[7]) / local_density
        u_y = (cells[2] + cells[5] + cells[6] - cells[4] - cells[7] - cells[8]) / local_density
    with np.errstate(invalid="ignore"):  # fp32 sum of squares, double sqrt, as d2q9-bgk.c:828
        u = np.sqrt(((u_x * u_x) + (u_y * u_y)).astype(np.float64)).astype(np.float32)
    pressure = local_density * c_sq
    u_x = np.where(blocked, np.float32(0), u_x)
    u_y = np.where(blocked, np.float32(0), u_y)
    u = np.where(blocked, np.float32(0), u)
    pressure = np.where(blocked, np.float32(params.density) * c_sq, pressure)
    return u_x, u_y, u, pressure


def _fmt_e12(a: np.ndarray) -> np.ndarray:
    return np.char.mod("%.12E", a.astype(np.float64).ravel())


def write_final_state(path: str, params: Params, cells: np.ndarray, obstacles: np.ndarray) -> None:
    """d2q9-bgk.c:835: ``"%d %d %.12E %.12E %.12E %.12E %d\\n"`` = x y u_x u_y u pressure obstacle."""
    u_x, u_y, u, pressure = final_state_fields(params, cells, obstacles)
    ii, jj = np.mgrid[0:params.ny, 0:params.nx]
    cols = [np.char.mod("%d", jj.ravel()), np.char.mod("%d", ii.ravel()), _fmt_e12(u_x), _fmt_e12(u_y),
            _fmt_e12(u), _fmt_e12(pressure), np.char.mod("%d", obstacles.reshape(-1))]
    lines = cols[0]
    for c in cols[1:]:
        lines = np.char.add(np.char.add(lines, " "), c)
    with open(path, "w") as fp:
        fp.write("\n".join(lines.tolist()))
        fp.write("\n")


def write_av_vels(path: str, av_vels: np.ndarray) -> None:
    """d2q9-bgk.c:850: ``"%d:\\t%.12E\\n"``."""
    with open(path, "w") as fp:
        for ii, v in enumerate(np.asarray(av_vels, dtype=np.float64)):
            fp.write("%d:\t%.12E\n" % (ii, v))


def write_values(params: Params, cells, obstacles, av_vels, outdir: str = ".") -> None:
    """write_values(), d2q9-bgk.c:772-856."""
    write_final_state(os.path.join(outdir, FINALSTATEFILE), params, cells, obstacles)
    write_av_vels(os.path.join(outdir, AVVELSFILE), av_vels)


def calc_reynolds(params: Params, av_velocity: float) -> float:
    """d2q9-bgk.c:747-752 given the final state's average velocity."""
    viscosity = np.float32(1.0) / np.float32(6.0) * (np.float32(2.0) / np.float32(params.omega) - np.float32(1.0))
    return float(np.float32(av_velocity) * np.float32(params.reynolds_dim) / viscosity)


def synthetic_channel(nx: int, ny: int, maxIters: int = 200, *, density=0.1, accel=0.005, omega=1.85,
                      reynolds_dim=10, block=64, spacing=1024, walls=True):
    """SURVEY.md §8(d) deck 5: channel walls on rows 0 and ny-1 plus solid ``block``² squares
    centred every ``spacing`` cells (≈0.4 % blocked).  Deterministic, no RNG.
    Returns (params, cells, obstacles)."""
    params = Params(nx=nx, ny=ny, maxIters=maxIters, reynolds_dim=reynolds_dim,
                    density=float(np.float32(density)), accel=float(np.float32(accel)),
                    omega=float(np.float32(omega)))
    obstacles = synthetic_channel_rows(nx, ny, 0, ny, block=block, spacing=spacing, walls=walls)
    free_cells = nx * ny - int(obstacles.sum(dtype=np.int64))
    params.free_cells_inv = float(np.float32(1.0) / np.float32(free_cells))
    return params, initial_cells(params), obstacles


def synthetic_channel_rows(nx: int, ny: int, y0: int, rows: int, *, block=64, spacing=1024, walls=True):
    """Rows [y0, y0+rows) of the synthetic channel's obstacle map (so a rank can build its slab)."""
    y = np.arange(y0, y0 + rows)
    x = np.arange(nx)
    half = block // 2
    centre = spacing // 2
    in_y = np.abs((y % spacing) - centre) < half if ny >= spacing else np.zeros(rows, dtype=bool)
    in_x = np.abs((x % spacing) - centre) < half if nx >= spacing else np.zeros(nx, dtype=bool)
    obstacles = (in_y[:, None] & in_x[None, :]).astype(np.int32)
    if walls:
        obstacles[(y == 0) | (y == ny - 1), :] = 1
    return obstacles


def synthetic_channel_free_cells(nx: int, ny: int, *, block=64, spacing=1024, walls=True) -> int:
    """Free-cell count of the synthetic channel without materialising the map."""
    total = 0
    chunk = 4096
    for y0 in range(0, ny, chunk):
        rows = min(chunk, ny - y0)
        total += rows * nx - int(synthetic_channel_rows(nx, ny, y0, rows, block=block, spacing=spacing,
                                                         walls=walls).sum(dtype=np.int64))
    return total


def perturbed_rows(cells: np.ndarray, y0: int) -> np.ndarray:
    """A smooth, cheap, deterministic perturbation of `cells` ([9, rows, nx], holding global rows
    [y0, y0+rows)) in place, so that every row and column evolves differently; a function of the GLOBAL
    cell coordinates only, so slabs built by different ranks agree with the whole grid built by one."""
    _, rows, nx = cells.shape
    x = np.arange(nx, dtype=np.float32)
    y = np.arange(y0, y0 + rows, dtype=np.float32)
    for k in range(NSPEEDS):
        sx = np.sin(x * np.float32(0.001 * (k + 1)))
        cy = np.cos(y * np.float32(0.0013 * (9 - k)))
        cells[k] *= (np.float32(1.0) + np.float32(0.02) * sx[None, :] * cy[:, None]).astype(np.float32)
    return cells


def bits_checksum(a: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum of the raw bits of a float32 array (position-weighted sum):
    equal checksums of two lattices = equal bits, cell for cell, for all practical purposes."""
    v = np.ascontiguousarray(a).reshape(-1).view(np.uint32).astype(np.uint64)
    w = (np.arange(v.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    return int((v * w).sum(dtype=np.uint64))

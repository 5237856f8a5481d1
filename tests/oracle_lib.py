"""ctypes loader for oracle/liboracle.so — the CHECKER, used by tests/, smoke() and bench.py's
cpu_baseline / reference-arm legs only (never by the product path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class OracleParams(C.Structure):
    _fields_ = [("density", C.c_float), ("accel", C.c_float), ("omega", C.c_float),
                ("free_cells_inv", C.c_float), ("nx", C.c_int), ("ny", C.c_int),
                ("maxIters", C.c_int), ("reynolds_dim", C.c_int)]


class OracleParams64(C.Structure):
    _fields_ = [("density", C.c_double), ("accel", C.c_double), ("omega", C.c_double),
                ("nx", C.c_int), ("ny", C.c_int), ("maxIters", C.c_int), ("reynolds_dim", C.c_int)]


def build_oracle() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def host_threads() -> int:
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arms undo that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as fp:
            return f" {flag} " in fp.read().replace("\n", " ") + " "
    except OSError:
        return False


def _cpu_has_avx2() -> bool:
    return _cpu_has("avx2") and _cpu_has("fma")


def _cpu_has_avx512() -> bool:
    return _cpu_has_avx2() and _cpu_has("avx512f") and _cpu_has("avx512vl")


def fastest_variant() -> str:
    """The widest build this CPU runs: 'avx512', 'avx2' or 'base'."""
    return "avx512" if _cpu_has_avx512() else "avx2" if _cpu_has_avx2() else "base"


def describe(variant: str) -> str:
    """How a build was made, for the bench records."""
    if variant == "fastest":
        variant = fastest_variant()
    common = "gcc -O3 -fopenmp -mfma -ffp-contract=off"
    return {"base": f"{common}, cell-by-cell row loop",
            "avx2": f"{common} -mavx2, vectorised row loop (8 columns per instruction)",
            "avx512": f"{common} -mavx512f, vectorised row loop (16 columns per instruction)"}[variant]


_lib_cache = {}


def load(variant: str = "base"):
    """variant: 'base' (-O3 -fopenmp, the north-star baseline flags, cell-by-cell row loop), 'avx2' / 'avx512'
    (the vectorisable row loop, bit-identical results), or 'fastest' (the widest this CPU runs)."""
    if variant == "fastest":
        variant = fastest_variant()
    if variant in _lib_cache:
        return _lib_cache[variant]
    name = {"base": "liboracle.so", "avx2": "liboracle_avx2.so", "avx512": "liboracle_avx512.so"}[variant]
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    P, P64 = C.POINTER(OracleParams), C.POINTER(OracleParams64)
    lib.oracle_f32_accelerate.argtypes = [P, fp, ip]
    lib.oracle_f32_accelerate.restype = None
    lib.oracle_f32_timestep.argtypes = [P, fp, fp, ip, C.c_int]
    lib.oracle_f32_timestep.restype = C.c_float
    lib.oracle_f32_run.argtypes = [P, fp, fp, ip, C.c_int, fp, C.c_int]
    lib.oracle_f32_run.restype = None
    lib.oracle_f32_av_velocity.argtypes = [P, fp, ip]
    lib.oracle_f32_av_velocity.restype = C.c_float
    lib.oracle_f32_reynolds.argtypes = [P, fp, ip]
    lib.oracle_f32_reynolds.restype = C.c_float
    lib.oracle_f32_total_density.argtypes = [P, fp]
    lib.oracle_f32_total_density.restype = C.c_float
    lib.oracle_f32_final_state.argtypes = [P, fp, ip, fp, fp, fp, fp]
    lib.oracle_f32_final_state.restype = None
    lib.oracle_f32_slab_accelerate.argtypes = [P, fp, ip, C.c_int, C.c_int]
    lib.oracle_f32_slab_accelerate.restype = None
    lib.oracle_f32_slab_timestep.argtypes = [P, fp, fp, ip, C.c_int, fp]
    lib.oracle_f32_slab_timestep.restype = None
    lib.oracle_f64_run.argtypes = [P64, dp, dp, ip, C.c_int, dp]
    lib.oracle_f64_run.restype = None
    lib.oracle_f64_av_velocity.argtypes = [P64, dp, ip]
    lib.oracle_f64_av_velocity.restype = C.c_double
    lib.oracle_f64_pressure.argtypes = [P64, dp, ip, dp]
    lib.oracle_f64_pressure.restype = None
    lib.oracle_f64_final_state.argtypes = [P64, dp, ip, dp, dp, dp, dp]
    lib.oracle_f64_final_state.restype = None
    lib.oracle_num_threads.argtypes = []
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_set_num_threads.argtypes = [C.c_int]
    lib.oracle_set_num_threads.restype = None
    _lib_cache[variant] = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def to_oracle_params(params) -> OracleParams:
    return OracleParams(params.density, params.accel, params.omega, params.free_cells_inv,
                        params.nx, params.ny, params.maxIters, params.reynolds_dim)


def run_f32(params, cells, obstacles, nsteps, reference_order=True, variant="base"):
    """Runs nsteps of the fp32 oracle; returns (final cells [9,ny,nx], av_vels[nsteps])."""
    lib = load(variant)
    op = to_oracle_params(params)
    cells = np.ascontiguousarray(cells, dtype=np.float32).copy()
    scratch = np.empty_like(cells)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.int32)
    av = np.zeros(max(nsteps, 1), dtype=np.float32)
    lib.oracle_f32_run(C.byref(op), _fp(cells), _fp(scratch), _ip(obstacles), nsteps, _fp(av),
                       1 if reference_order else 0)
    return cells, av[:nsteps]


def run_f64(params, cells, obstacles, nsteps):
    """Runs nsteps of the fp64 oracle from fp64-exact deck constants; returns (cells, av_vels, pressure)."""
    lib = load("base")
    op = OracleParams64(params.density, params.accel, params.omega, params.nx, params.ny,
                        params.maxIters, params.reynolds_dim)
    cells = np.ascontiguousarray(cells, dtype=np.float64).copy()
    scratch = np.empty_like(cells)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.int32)
    av = np.zeros(max(nsteps, 1), dtype=np.float64)
    lib.oracle_f64_run(C.byref(op), _dp(cells), _dp(scratch), _ip(obstacles), nsteps, _dp(av))
    pressure = np.empty((params.ny, params.nx), dtype=np.float64)
    lib.oracle_f64_pressure(C.byref(op), _dp(cells), _ip(obstacles), _dp(pressure))
    return cells, av[:nsteps], pressure


def final_state_f64(params, cells, obstacles):
    """(u_x, u_y, u, pressure) of an fp64 state, d2q9-bgk.c:789-831 in double."""
    lib = load("base")
    op = OracleParams64(params.density, params.accel, params.omega, params.nx, params.ny,
                        params.maxIters, params.reynolds_dim)
    cells = np.ascontiguousarray(cells, dtype=np.float64)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.int32)
    outs = [np.empty((params.ny, params.nx), dtype=np.float64) for _ in range(4)]
    lib.oracle_f64_final_state(C.byref(op), _dp(cells), _ip(obstacles), *[_dp(o) for o in outs])
    return outs


def final_state_f32(params, cells, obstacles):
    lib = load("base")
    op = to_oracle_params(params)
    cells = np.ascontiguousarray(cells, dtype=np.float32)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.int32)
    outs = [np.empty((params.ny, params.nx), dtype=np.float32) for _ in range(4)]
    lib.oracle_f32_final_state(C.byref(op), _fp(cells), _ip(obstacles), *[_fp(o) for o in outs])
    return outs


def av_velocity_f32(params, cells, obstacles) -> float:
    lib = load("base")
    op = to_oracle_params(params)
    return float(lib.oracle_f32_av_velocity(C.byref(op), _fp(np.ascontiguousarray(cells, dtype=np.float32)),
                                            _ip(np.ascontiguousarray(obstacles, dtype=np.int32))))


def reynolds_f32(params, cells, obstacles) -> float:
    lib = load("base")
    op = to_oracle_params(params)
    return float(lib.oracle_f32_reynolds(C.byref(op), _fp(np.ascontiguousarray(cells, dtype=np.float32)),
                                         _ip(np.ascontiguousarray(obstacles, dtype=np.int32))))

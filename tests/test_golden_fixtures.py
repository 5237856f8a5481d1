"""tests/golden/*.npz (made by tests/golden/make_fixtures.py from the fp32 oracle): the oracle must
reproduce them bit for bit on the CPU, and the CUDA path on the GPU — with every kernel family."""
import glob
import os

import numpy as np
import pytest

import helpers

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load(path):
    g = np.load(path)
    p, cells, obstacles = helpers.random_case(int(g["nx"]), int(g["ny"]), seed=int(g["seed"]), walls=bool(g["walls"]))
    return g, p, cells, obstacles


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_oracle_reproduces_fixture(oracle, path):
    g, p, cells, obstacles = load(path)
    out, av = oracle.run_f32(p, cells, obstacles, int(g["steps"]), reference_order=False)
    assert np.array_equal(out.view(np.uint32), g["cells_bits"])
    assert np.array_equal(av.view(np.uint32), g["av_bits"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
@pytest.mark.parametrize("opts", [{}, {"persistent": 0, "fuse2": 0}, {"persistent": 0, "fuse2": 1, "fuse2_rows": 8},
                                  {"persistent": 1, "cells_per_thread": 2}], ids=["auto", "one-step", "two-step", "persistent"])
def test_gpu_reproduces_fixture(lbm, path, opts):
    g, p, cells, obstacles = load(path)
    steps = int(g["steps"])
    with lbm.cabi.Simulation(p, options=opts) as sim:
        sim.upload(cells, obstacles)
        sim.run(steps)
        sim.sync()
        out, av = sim.download_cells(), sim.download_av_vels(steps)
    assert np.array_equal(out.view(np.uint32), g["cells_bits"])
    # the fixture's averages are the oracle's sequential fp32 sums; the GPU sums exactly and rounds once
    np.testing.assert_allclose(av, g["av_bits"].view(np.float32), rtol=2e-6, atol=0)

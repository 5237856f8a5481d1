"""BASELINE.json's full size (16384 x 16384, 268 M cells, 19 GB of lattice): the default two-step kernel
with its automatic 128/32-row tiling (only reachable at large slabs) against the CPU ORACLE itself —
five oracle time steps of the whole grid cost a few seconds on the box's cores (bench.py --impl
reference steps the same grid) — bit for bit over the whole lattice (order-sensitive checksum of the raw
bits) and on av_vels; then the size-independent properties: one-step kernel == two-step kernel ==
two-slab ring bitwise, mass conserved, values inside a solid block only move."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NX = NY = 16384
STEPS = 5          # two fused pairs + one single step


def checksum(a):
    """Order-sensitive 64-bit checksum of the raw bits."""
    v = a.reshape(-1).view(np.uint32).astype(np.uint64)
    w = (np.arange(v.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    return int((v * w).sum(dtype=np.uint64))


@pytest.fixture(scope="module")
def deck(lbm):
    free = lbm.cabi.load_library().lbm_device_count()
    if free < 1:
        pytest.skip("no GPU")
    p, cells, obstacles = lbm.decks.synthetic_channel(NX, NY)
    # a smooth, cheap, deterministic perturbation so that every row and column evolves differently
    x = np.arange(NX, dtype=np.float32)
    y = np.arange(NY, dtype=np.float32)
    for k in range(9):
        cells[k] *= (1.0 + 0.02 * np.sin(x * np.float32(0.001 * (k + 1)))[None, :]
                     * np.cos(y * np.float32(0.0013 * (9 - k)))[:, None]).astype(np.float32)
    return p, cells, obstacles


def run(lbm, deck, **kw):
    p, cells, obstacles = deck
    with lbm.cabi.Simulation(p, **kw) as sim:
        sim.upload(cells, obstacles)
        sim.run(STEPS)
        sim.sync()
        out = sim.download_cells()
        av = sim.download_av_vels(STEPS)
        info = sim.info()
    return out, av, info


def test_full_size_matches_oracle_and_kernels_agree(lbm, oracle, deck):
    p, cells, obstacles = deck
    two, av_two, info_two = run(lbm, deck)                                   # default: two-step kernel
    assert info_two["kernel_name"].startswith("fuse2q_kernel") and "rows=128/32" in info_two["kernel_name"]
    cs_two = checksum(two)
    # the oracle on the whole 16384 x 16384 grid, all host cores (oracle/lbm_oracle.c, OpenMP over rows)
    lib = oracle.load("fastest")
    lib.oracle_set_num_threads(oracle.host_threads())
    ref, ref_av = oracle.run_f32(p, cells, obstacles, STEPS, reference_order=False, variant="fastest")
    assert checksum(ref) == cs_two, "two-step kernel differs from the oracle at 16384 x 16384"
    assert np.array_equal(ref.view(np.uint32)[:, ::1021, ::509], two.view(np.uint32)[:, ::1021, ::509])
    np.testing.assert_allclose(av_two, ref_av, rtol=2e-6, atol=0)
    del ref
    mass0 = float(cells.sum(dtype=np.float64))
    mass1 = float(two.sum(dtype=np.float64))
    assert abs(mass1 - mass0) / mass0 < 1e-6
    assert np.all(np.isfinite(av_two)) and np.all(av_two > 0)
    # blocked cells never relax (kernels.cl:187-197 with lmask = 0): inside a solid block values are only
    # moved around, so the block's interior keeps its total mass exactly as a multiset sum in fp64
    yb, xb = 512, 512            # centre of a 64x64 solid block; STEPS cells away from its rim
    assert obstacles[yb - 8:yb + 8, xb - 8:xb + 8].all()
    core0 = np.sort(cells[:, yb - 8 + STEPS:yb + 8 - STEPS, xb - 8 + STEPS:xb + 8 - STEPS].reshape(-1))
    reach = np.sort(two[:, yb - 8:yb + 8, xb - 8:xb + 8].reshape(-1))
    assert np.isin(core0, reach).all()      # every value of the core is still somewhere in the block, unchanged
    del two

    one, av_one, info_one = run(lbm, deck, options={"fuse2": 0})               # one-step kernel
    assert info_one["kernel_name"].startswith("step_kernel")
    assert checksum(one) == cs_two
    assert np.array_equal(av_one.view(np.uint32), av_two.view(np.uint32))
    del one

    ring, av_ring, info_ring = run(lbm, deck, devices=[0, 0])                  # two row slabs, two-step kernel
    assert info_ring["nslabs"] == 2 and info_ring["kernel_name"].startswith("fuse2q_kernel")
    assert checksum(ring) == cs_two
    assert np.array_equal(av_ring.view(np.uint32), av_two.view(np.uint32))
    del ring

    # eight row slabs of 2048 rows: the per-GPU slab of the 8-GPU strong-scaling split and its own automatic
    # tiling (64-row segments, then 16-row ones)
    ring8, av_ring8, info_ring8 = run(lbm, deck, devices=[0] * 8)
    assert info_ring8["nslabs"] == 8 and info_ring8["kernel_name"].startswith("fuse2q_kernel")
    assert "rows=64/16" in info_ring8["kernel_name"], info_ring8
    assert checksum(ring8) == cs_two
    assert np.array_equal(av_ring8.view(np.uint32), av_two.view(np.uint32))

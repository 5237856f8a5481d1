"""GPU parity: the CUDA path through the C-ABI against the CPU oracle (bit-exact lattice) and the
reference's goldens (check.py's 1 % gate).  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import helpers
from helpers import bits, pct_diff, random_case

pytestmark = pytest.mark.gpu

# av_vels: the GPU sums speeds exactly (double-double) and scales once; the reference scales each
# work-item by FREE_CELLS_INV and adds in a fp32 tree (kernels.cl:202-229) — a few fp32 ulps apart.
AV_RTOL = 2e-6

SHAPES = [(128, 128), (256, 64), (100, 37), (34, 9), (33, 5), (8, 3), (1024, 16), (4096, 8), (4, 2), (1, 4)]


def run_gpu(lbm, p, cells, obstacles, nsteps, **kw):
    with lbm.cabi.Simulation(p, **kw) as sim:
        sim.upload(cells, obstacles)
        sim.run(nsteps)
        sim.sync()
        return sim.download_cells(), sim.download_av_vels(nsteps), sim.info()


@pytest.mark.parametrize("nx,ny", SHAPES)
def test_lattice_bit_exact_vs_oracle(lbm, oracle, nx, ny):
    p, cells, obstacles = random_case(nx, ny, seed=nx * 1000 + ny)
    nsteps = 7
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, nsteps)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, nsteps)
    assert np.array_equal(bits(got_cells), bits(ref_cells)), info
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


@pytest.mark.parametrize("V", [1, 2, 4])
@pytest.mark.parametrize("tpb", [128, 256, 512])
@pytest.mark.parametrize("streaming", [0, 1])
def test_kernel_variants_bit_exact(lbm, oracle, V, tpb, streaming):
    p, cells, obstacles = random_case(384, 24, seed=7, walls=False)  # open edges: y wrap carries fluid
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 5)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 5,
                                      options={"cells_per_thread": V, "threads_per_block": tpb, "streaming": streaming,
                                               "persistent": 0, "fuse2": 0})
    assert info["cells_per_thread"] == V and info["threads_per_block"] == tpb and info["streaming"] == streaming
    assert info["kernel_name"].startswith("step_kernel")
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


@pytest.mark.parametrize("V", [1, 2, 4])
@pytest.mark.parametrize("tpb", [128, 256, 512])
def test_persistent_kernel_bit_exact(lbm, oracle, V, tpb):
    """The multi-step cooperative kernel (grid barrier between steps) against the oracle and against
    the one-launch-per-step kernel; 37 steps with chunk_steps=8 also crosses launch boundaries."""
    p, cells, obstacles = random_case(384, 24, seed=7, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 37)
    opts = {"cells_per_thread": V, "threads_per_block": tpb, "persistent": 1, "chunk_steps": 8}
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 37, options=opts)
    assert info["kernel_name"].startswith("persistent_kernel")
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)
    # same segment size (cells_per_thread) => bitwise the same averages as one launch per step
    one_cells, one_av, _ = run_gpu(lbm, p, cells, obstacles, 37, options={"persistent": 0, "cells_per_thread": V})
    assert np.array_equal(bits(got_cells), bits(one_cells)) and np.array_equal(bits(got_av), bits(one_av))


TILE_SHAPES = [(128, 128), (128, 256), (256, 256), (100, 37), (34, 9), (33, 5), (8, 4), (1, 4), (300, 200), (4096, 8),
               (5, 700), (384, 384), (512, 512), (500, 301)]      # the last three: two cells per thread


@pytest.mark.parametrize("nx,ny", TILE_SHAPES)
def test_tile_kernel_bit_exact(lbm, oracle, nx, ny):
    """The small-deck default (lbm_tile.cuh: K time steps per hand-off on shared-memory tiles with a K-cell
    halo, flags between neighbouring tiles): lattice bit-exact vs the oracle for square, ragged, one-column
    and long thin grids; 11 steps = rounds of 4 + 4 + 3."""
    p, cells, obstacles = random_case(nx, ny, seed=nx * 31 + ny, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 11)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 11, options={"tile": 1})   # (forced: also one-step rounds)
    assert info["kernel_name"].startswith("tile_kernel<"), info
    assert ("2/thread" in info["kernel_name"]) == (nx * ny > 120000), info
    assert np.array_equal(bits(got_cells), bits(ref_cells)), info
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("tw,th", [(0, 0), (16, 8), (32, 4), (13, 7), (128, 2)])
def test_tile_kernel_variants(lbm, oracle, k, tw, th):
    """Steps per hand-off and tile shape change nothing: same bits as the oracle, and the averages (exact
    double-double sums of the cells' speeds) are bitwise those of the default tiling."""
    p, cells, obstacles = random_case(128, 64, seed=5 * k + tw, walls=False)
    ref_cells, _ = oracle.run_f32(p, cells, obstacles, 13)
    _, base_av, _ = run_gpu(lbm, p, cells, obstacles, 13)
    opts = {"tile": 1, "tile_steps": k, "tile_w": tw, "tile_h": th}
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 13, options=opts)
    assert info["kernel_name"].startswith("tile_kernel<"), info
    assert np.array_equal(bits(got_cells), bits(ref_cells)), info
    assert np.array_equal(bits(got_av), bits(base_av)), info


@pytest.mark.parametrize("chunk", [8, 6, 1])
def test_tile_kernel_across_launches(lbm, oracle, chunk):
    """37 steps in launches of `chunk` steps (not a multiple of the round length for 6): the buffer parity,
    the accelerate fold at launch boundaries and the ghost rows survive; run(5)+run(6) == run(11)."""
    p, cells, obstacles = random_case(128, 96, seed=3, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 37)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 37, options={"chunk_steps": chunk})
    assert info["kernel_name"].startswith("tile_kernel<")
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)
    with lbm.cabi.Simulation(p, options={"chunk_steps": chunk}) as sim:
        sim.upload(cells, obstacles)
        for n in (5, 6, 26):
            sim.run(n)
        sim.sync()
        b_cells, b_av = sim.download_cells(), sim.download_av_vels(37)
    assert np.array_equal(bits(b_cells), bits(ref_cells)) and np.array_equal(bits(b_av), bits(got_av))


def test_tile_kernel_then_other_kernels(lbm, oracle):
    """The tile kernel leaves the arena's ghost rows consistent: switching to the one-step kernel mid-run
    (set_option re-plans) continues bit-exactly."""
    p, cells, obstacles = random_case(256, 48, seed=12, walls=False)
    ref_cells, _ = oracle.run_f32(p, cells, obstacles, 12)
    with lbm.cabi.Simulation(p) as sim:
        sim.upload(cells, obstacles)
        sim.run(7)
        assert sim.info()["kernel_name"].startswith("tile_kernel<")
        sim.set_option("persistent", 0)
        assert sim.info()["kernel_name"].startswith("step_kernel<")
        sim.run(5)
        sim.sync()
        got = sim.download_cells()
    assert np.array_equal(bits(got), bits(ref_cells))


def test_persistent_grid_larger_than_device(lbm, oracle):
    """More warp segments than co-resident warps: blocks loop over several segments per step."""
    p, cells, obstacles = random_case(2048, 1024, seed=21)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 4)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 4, options={"persistent": 1, "cells_per_thread": 1})
    assert info["kernel_name"].startswith("persistent_kernel")
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


@pytest.mark.parametrize("tps", [768, 1024])
@pytest.mark.parametrize("packed", [0, 1])
@pytest.mark.parametrize("V", [2, 4])
@pytest.mark.parametrize("persistent", [0, 1])
def test_packed_and_register_bound_variants_bit_exact(lbm, oracle, tps, packed, V, persistent):
    """Scalar and packed-fp32x2 (sm_100 FADD2/FMUL2/FFMA2) instantiations give the same bits."""
    p, cells, obstacles = random_case(512, 20, seed=8, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 5)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 5,
                                      options={"persistent": persistent, "threads_per_sm": tps, "packed": packed,
                                               "cells_per_thread": V, "fuse2": 0})
    assert f"packed={packed}" in info["kernel_name"]
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


FUSE2_SHAPES = [(256, 64), (1024, 16), (100, 37), (4096, 8), (520, 40), (8, 4), (16384, 12)]


F2_NAMES = {2: "fuse2p_kernel", 3: "fuse2q_kernel"}


@pytest.mark.parametrize("nx,ny", FUSE2_SHAPES)
@pytest.mark.parametrize("nsteps", [8, 7])
@pytest.mark.parametrize("tma", [3, 2])
def test_two_step_kernel_bit_exact(lbm, oracle, nx, ny, nsteps, tma):
    """Temporal blocking (two time steps per HBM pass, step-1 rows in a shared-memory ring) gives the
    same bits as the oracle; 7 steps = three fused pairs + one single step."""
    p, cells, obstacles = random_case(nx, ny, seed=nx + ny, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, nsteps, reference_order=(nx % 128 == 0))
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, nsteps,
                                      options={"persistent": 0, "fuse2": 1, "fuse2_tma": tma})
    assert info["kernel_name"].startswith(F2_NAMES[tma] + "<") and info["steps_per_launch"] == 2
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


@pytest.mark.parametrize("packed", [0, 1])
@pytest.mark.parametrize("seg_rows", [4, 10, 256])
@pytest.mark.parametrize("tma", [3, 2])
def test_two_step_kernel_variants(lbm, oracle, packed, seg_rows, tma):
    """Row-segment length (redundant warm-up rows at every segment start), packed arithmetic and the stage depth
    (fuse2q_kernel: two-deep stage + a second barrier per row; fuse2p_kernel: its predecessor) do not change a
    bit; av_vels equal the one-step kernel's bitwise.  (Scalar arithmetic always runs fuse2p_kernel.)"""
    p, cells, obstacles = random_case(1280, 37, seed=77, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 6)
    opts = {"persistent": 0, "fuse2": 1, "fuse2_rows": seg_rows, "packed": packed, "fuse2_tma": tma}
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 6, options=opts)
    assert info["kernel_name"].startswith((F2_NAMES[tma] if packed else "fuse2p_kernel") + "<W=4"), info
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    one_cells, one_av, _ = run_gpu(lbm, p, cells, obstacles, 6, options={"persistent": 0, "fuse2": 0, "cells_per_thread": 4})
    assert np.array_equal(bits(got_av), bits(one_av))


@pytest.mark.parametrize("nx,ny", [(1024, 40), (1280, 37), (100, 37), (8, 4)])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("packed", [0, 1])
@pytest.mark.parametrize("tma", [3, 2])
def test_repipelined_two_step_kernel(lbm, oracle, nx, ny, mode, packed, tma):
    """fuse2p_kernel (body warps arrive on an `empty` mbarrier once the stage is in registers, the halo
    warp requests the next row's bulk copies; obstacle words and the periodic wrap columns travel with the
    copies; packed reciprocal / square root with one range check per pair or per thread): full-width and
    ragged strips, including the outermost strips' wrap; bits equal the oracle's, av_vels the one-step kernel's.
    fuse2q_kernel (two-deep stage) exists for packed arithmetic with the per-thread range check; the other
    combinations run fuse2p_kernel whatever fuse2_tma says."""
    p, cells, obstacles = random_case(nx, ny, seed=nx * 7 + ny, walls=False)
    ref_cells, _ = oracle.run_f32(p, cells, obstacles, 9)
    opts = {"persistent": 0, "fuse2": 1, "fuse2_tma": tma, "fuse2_rows": 8, "fuse2_mode": mode, "packed": packed}
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 9, options=opts)
    assert info["kernel_name"].startswith((F2_NAMES[tma] if packed and mode else "fuse2p_kernel") + "<")
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    _, one_av, _ = run_gpu(lbm, p, cells, obstacles, 9, options={"persistent": 0, "fuse2": 0, "cells_per_thread": 4})
    assert np.array_equal(bits(got_av), bits(one_av))


@pytest.mark.parametrize("tma", [3, 2])
@pytest.mark.parametrize("nx,ny,long_rows,short_rows,nslabs", [(1024, 40, 12, 4, 1), (1280, 37, 10, 8, 1), (512, 41, 6, 4, 2),
                                                                (1024, 64, 16, 4, 3)])
def test_two_step_kernel_two_segment_sizes(lbm, oracle, nx, ny, long_rows, short_rows, nslabs, tma):
    """Long row segments first, short ones for the last quarter of a slab (the automatic tiling of large slabs,
    forced here on small ones): same bits as the oracle and as the one-step kernel, also on a ring of slabs
    whose edge segments have different lengths."""
    p, cells, obstacles = random_case(nx, ny, seed=nx + 3 * ny, walls=False)
    ref_cells, _ = oracle.run_f32(p, cells, obstacles, 7)
    opts = {"persistent": 0, "fuse2": 1, "fuse2_tma": tma, "fuse2_rows": short_rows, "fuse2_long": long_rows}
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 7, devices=[0] * nslabs, options=opts)
    assert info["kernel_name"].startswith(F2_NAMES[tma] + "<") and f"rows={long_rows}/{short_rows}>" in info["kernel_name"]
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    _, one_av, _ = run_gpu(lbm, p, cells, obstacles, 7, options={"persistent": 0, "fuse2": 0, "cells_per_thread": 4})
    assert np.array_equal(bits(got_av), bits(one_av))


@pytest.mark.parametrize("nx,ny", [(1536, 1536), (1792, 1200), (2048, 1100)])
def test_two_step_kernel_mid_size_automatic_tiling(lbm, oracle, nx, ny):
    """Mid-size lattices (streamed from HBM, but only a few waves of blocks): the automatic choice is the two-step
    kernel with the segment length that fills the last wave (odd lengths such as 11 or 19 rows, ragged last
    segments and, for nx = 1792, a ragged last strip) — bits equal the oracle's."""
    p, cells, obstacles = lbm.decks.synthetic_channel(nx, ny, block=16, spacing=128)
    cells = lbm.decks.perturbed_rows(cells, 0)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 5, reference_order=False)
    got_cells, got_av, info = run_gpu(lbm, p, cells, obstacles, 5)
    assert info["kernel_name"].startswith("fuse2q_kernel<"), info
    assert np.array_equal(bits(got_cells), bits(ref_cells)), info
    np.testing.assert_allclose(got_av, ref_av, rtol=1e-5, atol=0)   # (the oracle adds its row sums in fp32)


def test_fast_reciprocal_and_square_root_exhaustive(lbm):
    """rcp_rn_fast / sqrt_rn_fast (MUFU + packed Newton step, one shared range check) return the bits of
    __frcp_rn / __fsqrt_rn for every one of the 2^32 float bit patterns inside their range."""
    rcp_bad, sqrt_bad = lbm.cabi.fastmath_mismatches()
    assert (rcp_bad, sqrt_bad) == (0, 0)


@pytest.mark.parametrize("tma", [3, 2])
def test_two_step_kernel_cells_at_rest(lbm, oracle, tma):
    """Exactly symmetric dyadic populations: u_sq is exactly 0 in the first step (and wherever the
    symmetry survives), which is outside the fast square root's range — the built-in fall-back
    must give the oracle's bits."""
    p, cells, obstacles = random_case(1024, 24, seed=5, walls=False, perturb=0.0)
    cells[0] = 0.5
    cells[1:5] = 0.125
    cells[5:9] = 0.03125
    ref_cells, _ = oracle.run_f32(p, cells, obstacles, 4)
    got_cells, _, info = run_gpu(lbm, p, cells, obstacles, 4,
                                 options={"persistent": 0, "fuse2": 1, "fuse2_tma": tma, "fuse2_rows": 8})
    assert info["kernel_name"].startswith(F2_NAMES[tma] + "<")
    assert np.array_equal(bits(got_cells), bits(ref_cells))


@pytest.mark.parametrize("nslabs", [2, 3])
def test_two_step_kernel_row_slabs(lbm, nslabs):
    """Two-step kernel on a ring of slabs (two-deep ghost rows written into the neighbours, epoch flags
    per launch) == single slab, one-step kernel: lattice and av_vels bitwise; odd step count."""
    p, cells, obstacles = random_case(512, 41, seed=13, walls=False)
    a_cells, a_av, _ = run_gpu(lbm, p, cells, obstacles, 9, options={"persistent": 0, "fuse2": 0, "cells_per_thread": 4})
    b_cells, b_av, info = run_gpu(lbm, p, cells, obstacles, 9, devices=[0] * nslabs,
                                  options={"fuse2": 1, "fuse2_rows": 8})
    assert info["nslabs"] == nslabs and info["kernel_name"].startswith("fuse2")
    assert np.array_equal(bits(a_cells), bits(b_cells))
    assert np.array_equal(bits(a_av), bits(b_av))


def test_two_step_split_runs(lbm):
    p, cells, obstacles = random_case(256, 32, seed=3)
    opts = {"persistent": 0, "fuse2": 1}
    a_cells, a_av, _ = run_gpu(lbm, p, cells, obstacles, 11, options=opts)
    with lbm.cabi.Simulation(p, options=opts) as sim:
        sim.upload(cells, obstacles)
        sim.run(5)
        sim.run(6)
        sim.sync()
        b_cells, b_av = sim.download_cells(), sim.download_av_vels(11)
        assert sim.pad_nonzero() == 0
    assert np.array_equal(bits(a_cells), bits(b_cells))
    assert np.array_equal(bits(a_av), bits(b_av))


def test_split_runs_equal_one_run(lbm):
    """accelerate is applied on the write side for the next step and skipped on the last step of a
    run: run(5)+run(6) must be run(11) bit for bit, and an odd count reads the right buffer."""
    p, cells, obstacles = random_case(256, 32, seed=3)
    a_cells, a_av, _ = run_gpu(lbm, p, cells, obstacles, 11)
    with lbm.cabi.Simulation(p) as sim:
        sim.upload(cells, obstacles)
        sim.run(5)
        sim.run(6)
        sim.sync()
        b_cells, b_av = sim.download_cells(), sim.download_av_vels(11)
    assert np.array_equal(bits(a_cells), bits(b_cells))
    assert np.array_equal(bits(a_av), bits(b_av))


@pytest.mark.parametrize("nslabs", [2, 3, 5])
def test_row_slabs_on_one_gpu_equal_single(lbm, nslabs):
    """The multi-slab path (ghost rows, edge stores into the neighbour, epoch flags) on ONE device:
    lattice and av_vels bitwise equal to the single-slab run."""
    p, cells, obstacles = random_case(256, 41, seed=11, walls=False)
    a_cells, a_av, _ = run_gpu(lbm, p, cells, obstacles, 9, options={"cells_per_thread": 4, "persistent": 0})
    b_cells, b_av, info = run_gpu(lbm, p, cells, obstacles, 9, devices=[0] * nslabs, options={"cells_per_thread": 4})
    assert info["nslabs"] == nslabs
    assert np.array_equal(bits(a_cells), bits(b_cells))
    assert np.array_equal(bits(a_av), bits(b_av))


def test_chunked_av_vels(lbm):
    p, cells, obstacles = random_case(128, 16, seed=5)
    _, a_av, _ = run_gpu(lbm, p, cells, obstacles, 23)
    _, b_av, _ = run_gpu(lbm, p, cells, obstacles, 23, options={"chunk_steps": 4})
    assert np.array_equal(bits(a_av), bits(b_av))


def test_mass_conserved(lbm):
    """total_density (d2q9-bgk.c:754-770) is invariant: collision conserves mass, rebound permutes,
    accelerate moves mass between links of a cell."""
    p, cells, obstacles = random_case(512, 64, seed=9)
    got, _, _ = run_gpu(lbm, p, cells, obstacles, 50)
    before = cells.astype(np.float64).sum()
    after = got.astype(np.float64).sum()
    assert abs(after - before) / before < 1e-6


@pytest.mark.parametrize("name", ["128x128", "128x256", "256x256", "1024x1024"])
def test_reference_decks_pass_checker(lbm, name):
    """The four reference decks at full length against the reference's goldens with check.py's measure
    (1 % on av_vels and on final-state pressure; check/check.py:84-135)."""
    p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(name))
    got_cells, got_av, _ = run_gpu(lbm, p, cells, obstacles, p.maxIters)
    worst_av, step = pct_diff(helpers.golden_av_vels(name), got_av)
    assert np.isfinite(worst_av) and abs(worst_av) < 1.0, (worst_av, step)
    ref_pressure = helpers.golden_pressure(name)
    if ref_pressure is not None:
        _, _, _, pressure = lbm.decks.final_state_fields(p, got_cells, obstacles)
        worst_p, where = pct_diff(ref_pressure, pressure.ravel())
        assert np.isfinite(worst_p) and abs(worst_p) < 1.0, (worst_p, where)


def test_reynolds_known_answer(lbm):
    """README.md:78 of the reference: Reynolds number 9.763598020526E+00 for 128x128 (fp64 serial);
    the fp32 path must land within the checker's 1 %."""
    p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths("128x128"))
    got_cells, got_av, _ = run_gpu(lbm, p, cells, obstacles, p.maxIters)
    re = lbm.decks.calc_reynolds(p, float(got_av[-1]))
    assert abs(re - 9.763598020526) / 9.763598020526 < 0.01


def test_full_row_length_slab_vs_oracle(lbm, oracle):
    """BASELINE's row length (nx = 16384) on a slab the oracle finishes in seconds."""
    p, cells, obstacles = lbm.decks.synthetic_channel(16384, 256, block=16, spacing=128)
    rng = np.random.default_rng(1)
    cells = (cells * (1.0 + 0.1 * (rng.random(cells.shape, dtype=np.float32) - 0.5))).astype(np.float32)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 3, reference_order=False)
    got_cells, got_av, _ = run_gpu(lbm, p, cells, obstacles, 3, options={"streaming": 1})
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    np.testing.assert_allclose(got_av, ref_av, rtol=AV_RTOL, atol=0)


@pytest.mark.parametrize("nx,ny,opts", [(100, 37, {}), (34, 9, {}), (33, 5, {}), (100, 37, {"persistent": 0}),
                                        (130, 40, {"persistent": 0, "cells_per_thread": 2}), (1, 4, {})])
def test_pad_columns_never_written(lbm, nx, ny, opts):
    """Out-of-bounds canary (no compute-sanitizer on the pool): pad columns [nx, pitch) stay zero, also
    for the multi-slab ring whose edge rows are stored into neighbours' ghost rows."""
    p, cells, obstacles = random_case(nx, ny, seed=17)
    for kw in ({}, {"devices": [0, 0, 0]} if ny >= 3 else {}):
        with lbm.cabi.Simulation(p, options=opts, **kw) as sim:
            sim.upload(cells, obstacles)
            sim.run(9)
            sim.sync()
            assert sim.info()["pitch"] > nx
            assert sim.pad_nonzero() == 0


@pytest.mark.parametrize("nx,ny,kw", [(128, 128, {}), (100, 37, {}), (256, 41, {"devices": [0, 0, 0]})])
def test_final_state_fields_on_device(lbm, oracle, nx, ny, kw):
    """Output stage on the GPU == the host maths of write_values (d2q9-bgk.c:789-831), bit for bit."""
    p, cells, obstacles = random_case(nx, ny, seed=31)
    with lbm.cabi.Simulation(p, **kw) as sim:
        sim.upload(cells, obstacles)
        sim.run(6)
        sim.sync()
        got_cells = sim.download_cells()
        fields = sim.download_final_state()
    ref = oracle.final_state_f32(p, got_cells, obstacles)
    for name, g, r in zip(("u_x", "u_y", "u", "pressure"), fields, ref):
        assert np.array_equal(bits(g), bits(r)), name


def test_tall_narrow_grid(lbm, oracle):
    """More than 65535 rows (the gridDim.y limit the upload's pack kernel and the output stage sit on):
    lattice bit-exact vs the oracle, output stage bit-exact vs the host maths."""
    p, cells, obstacles = random_case(64, 70001, seed=64)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, 3, reference_order=False)
    with lbm.cabi.Simulation(p) as sim:
        sim.upload(cells, obstacles)
        sim.run(3)
        sim.sync()
        got_cells, got_av = sim.download_cells(), sim.download_av_vels(3)
        fields = sim.download_final_state()
        fields_again = sim.download_final_state()      # staging buffers are kept and reused
    assert np.array_equal(bits(got_cells), bits(ref_cells))
    # the oracle adds 70 001 row sums one after the other in fp32 (nx is not the reference's multiple of 128, so
    # there is no reference tree to follow): its own rounding error is ~1e-5; the GPU sums exactly
    np.testing.assert_allclose(got_av, ref_av, rtol=1e-4, atol=0)
    exact = [float(v) for v in got_av]
    assert all(np.isfinite(exact)) and all(v > 0 for v in exact)
    ref = oracle.final_state_f32(p, got_cells, obstacles)
    for name, g, g2, r in zip(("u_x", "u_y", "u", "pressure"), fields, fields_again, ref):
        assert np.array_equal(bits(g), bits(r)) and np.array_equal(bits(g2), bits(r)), name


@pytest.mark.parametrize("nx,ny,kw", [(100, 37, {}), (1024, 40, {"options": {"persistent": 0, "fuse2": 1}}),
                                      (256, 41, {"devices": [0, 0, 0]})])
def test_upload_packed_equals_upload(lbm, nx, ny, kw):
    """lbm_upload_packed (obstacle bit mask packed by the caller, lbm_pack_obstacles) == lbm_upload (the
    reference-shaped int map, packed on the device): same lattice, same averages."""
    p, cells, obstacles = random_case(nx, ny, seed=nx + ny)
    a_cells, a_av, _ = run_gpu(lbm, p, cells, obstacles, 7, **kw)
    words = lbm.cabi.pack_obstacles(obstacles)
    assert words.shape == (ny, (nx + 31) // 32)
    with lbm.cabi.Simulation(p, **kw) as sim:
        sim.upload_packed(cells, words)
        sim.run(7)
        sim.sync()
        b_cells, b_av = sim.download_cells(), sim.download_av_vels(7)
    assert np.array_equal(bits(a_cells), bits(b_cells)) and np.array_equal(bits(a_av), bits(b_av))


def test_pinned_numa_local_buffers(lbm):
    """lbm_host_alloc_on: pinned memory usable for upload / download like any host array."""
    p, cells, obstacles = random_case(128, 16, seed=2)
    pc = lbm.cabi.PinnedArray(cells.shape, np.float32, 0)
    po = lbm.cabi.PinnedArray(obstacles.shape, np.int32, 0)
    out = lbm.cabi.PinnedArray(cells.shape, np.float32, 0)
    pc.array[...] = cells
    po.array[...] = obstacles
    a_cells, _, _ = run_gpu(lbm, p, cells, obstacles, 4)
    with lbm.cabi.Simulation(p) as sim:
        sim.upload(pc.array, po.array)
        sim.run(4)
        sim.sync()
        sim.download_cells(out.array)
    assert np.array_equal(bits(out.array), bits(a_cells))
    assert isinstance(pc.numa_node, int)
    for buf in (pc, po, out):
        buf.close()


def test_current_device_is_left_alone(lbm):
    """The library sets the device it needs and puts the caller's back."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs to tell devices apart")
    torch.cuda.set_device(1)
    p, cells, obstacles = random_case(64, 8)
    run_gpu(lbm, p, cells, obstacles, 2, devices=[0])
    assert torch.cuda.current_device() == 1
    torch.cuda.set_device(0)


def test_errors_are_loud(lbm):
    p, cells, obstacles = random_case(64, 8)
    with lbm.cabi.Simulation(p) as sim:
        with pytest.raises(lbm.cabi.LbmError, match="before lbm_upload"):
            sim.run(1)
        with pytest.raises(lbm.cabi.LbmError, match="unknown option"):
            sim.set_option("nonsense", 1)
        sim.upload(cells, obstacles)
        sim.run(2)
        with pytest.raises(lbm.cabi.LbmError, match="asked for"):
            sim.download_av_vels(3)

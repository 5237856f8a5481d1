"""Host logic on CPU: deck formats and messages (Python mirror and the C program), output formats,
row partition, the error-free av_vels combine — no device needed."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers
from opencl_lattice_boltzmann_b200 import decks, ring

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "d2q9-bgk")


# ---- .params / obstacle files ---------------------------------------------------------------

def test_shipped_decks_parse():
    expect = {"128x128": (128, 128, 40000, 15876), "128x256": (128, 256, 40000, 32130),
              "256x256": (256, 256, 80000, 64516), "1024x1024": (1024, 1024, 20000, 1043462)}
    for name, (nx, ny, iters, free) in expect.items():
        p, cells, obstacles = decks.load_deck(*decks.deck_paths(name))
        assert (p.nx, p.ny, p.maxIters, p.reynolds_dim) == (nx, ny, iters, 10)
        assert nx * ny - int(obstacles.sum()) == free           # duplicate lines are not double counted
        assert np.float32(p.free_cells_inv) == np.float32(1.0) / np.float32(free)
        assert cells.shape == (9, ny, nx) and cells.dtype == np.float32
        # d2q9-bgk.c:529-550
        d = np.float32(p.density)
        assert cells[0, 0, 0] == d * np.float32(4.0) / np.float32(9.0)
        assert cells[3, 5, 7] == d / np.float32(9.0) and cells[8, 1, 1] == d / np.float32(36.0)


def write(tmp_path, name, text):
    path = tmp_path / name
    path.write_text(text)
    return str(path)


GOOD_PARAMS = "8\n6\n10\n10\n0.1\n0.005\n1.85\n"

BAD_OBSTACLES = [
    ("1 1\n", "expected 3 values per line in obstacle file"),
    ("8 1 1\n", "obstacle x-coord out of range"),
    ("-1 1 1\n", "obstacle x-coord out of range"),
    ("1 6 1\n", "obstacle y-coord out of range"),
    ("1 1 2\n", "obstacle blocked value should be 1"),
]


@pytest.mark.parametrize("text,message", BAD_OBSTACLES)
def test_obstacle_errors_python(tmp_path, text, message):
    pf = write(tmp_path, "p.params", GOOD_PARAMS)
    of = write(tmp_path, "o.dat", text)
    with pytest.raises(decks.DeckError, match=re.escape(message)):
        decks.load_deck(pf, of)


def test_empty_obstacle_file_and_duplicates(tmp_path):
    """An empty obstacle list is legal (the reference's fscanf loop simply ends, d2q9-bgk.c:571); duplicate
    lines are not counted twice (d2q9-bgk.c:583-585)."""
    pf = write(tmp_path, "p.params", GOOD_PARAMS)
    p, cells, obstacles = decks.load_deck(pf, write(tmp_path, "empty.dat", ""))
    assert obstacles.sum() == 0 and np.float32(p.free_cells_inv) == np.float32(1.0) / np.float32(48)
    p, cells, obstacles = decks.load_deck(pf, write(tmp_path, "dup.dat", "1 1 1\n1 1 1\n2 3 1\n"))
    assert obstacles.sum() == 2 and np.float32(p.free_cells_inv) == np.float32(1.0) / np.float32(46)


def test_param_errors_python(tmp_path):
    with pytest.raises(decks.DeckError, match="could not open input parameter file"):
        decks.read_params(str(tmp_path / "missing.params"))
    pf = write(tmp_path, "p.params", "8\n6\nten\n")
    with pytest.raises(decks.DeckError, match="could not read param file: maxIters"):
        decks.read_params(pf)
    pf = write(tmp_path, "p2.params", "8\n6\n10\n10\n0.1\n")
    with pytest.raises(decks.DeckError, match="could not read param file: accel"):
        decks.read_params(pf)
    pf = write(tmp_path, "p3.params", GOOD_PARAMS)
    with pytest.raises(decks.DeckError, match="could not open input obstacles file"):
        decks.read_obstacles(str(tmp_path / "missing.dat"), decks.read_params(pf))


# ---- the C program's command line and messages (d2q9-bgk.c:183-186, :868-880) ----------------

needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="d2q9-bgk not built (run make)")


@needs_exe
def test_cli_usage():
    for argv in ([EXE], [EXE, "a"], [EXE, "a", "b", "c"]):
        r = subprocess.run(argv, capture_output=True, text=True)
        assert r.returncode != 0
        assert r.stderr.strip() == f"Usage: {EXE} <paramfile> <obstaclefile>"


@needs_exe
@pytest.mark.parametrize("text,message", BAD_OBSTACLES)
def test_cli_obstacle_errors(tmp_path, text, message):
    pf = write(tmp_path, "p.params", GOOD_PARAMS)
    of = write(tmp_path, "o.dat", text)
    r = subprocess.run([EXE, pf, of], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0
    lines = r.stderr.strip().splitlines()
    assert re.match(r"Error at line \d+ of file .*:", lines[0]) and lines[1] == message


@needs_exe
def test_cli_param_errors(tmp_path):
    r = subprocess.run([EXE, str(tmp_path / "nope.params"), "x"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and f"could not open input parameter file: {tmp_path / 'nope.params'}" in r.stderr
    pf = write(tmp_path, "p.params", "8\n6\n10\n10\n0.1\nxyz\n")
    r = subprocess.run([EXE, pf, "x"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "could not read param file: accel" in r.stderr
    pf = write(tmp_path, "p2.params", GOOD_PARAMS)
    r = subprocess.run([EXE, pf, str(tmp_path / "nope.dat")], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "could not open input obstacles file" in r.stderr


@needs_exe
def test_cli_without_gpu_fails_loudly(tmp_path, lbm):
    if lbm.cabi.load_library().lbm_device_count() > 0:
        pytest.skip("a GPU is visible")
    pf, of = decks.deck_paths("128x128")
    r = subprocess.run([EXE, pf, of], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "LBM error during 'creating device context'" in r.stderr
    assert not os.path.exists(tmp_path / "av_vels.dat")


# ---- output formats (d2q9-bgk.c:835, :850) -----------------------------------------------------

def test_output_formats_and_checker_roundtrip(tmp_path, oracle):
    """oracle state -> write_values -> check.py against files written from the same state: exercises
    the writers and the checker's parser on the exact reference formats."""
    p, cells, obstacles = decks.load_deck(*decks.deck_paths("128x128"))
    got, av = oracle.run_f32(p, cells, obstacles, 30)
    p.maxIters = 30
    decks.write_values(p, got, obstacles, av, outdir=str(tmp_path))
    fs = open(tmp_path / "final_state.dat").read().splitlines()
    assert len(fs) == 128 * 128
    assert re.fullmatch(r"0 0 0\.000000000000E\+00 0\.000000000000E\+00 0\.000000000000E\+00 3\.333333\d{6}E-02 1", fs[0])
    assert re.fullmatch(r"\d+ \d+ -?\d\.\d{12}E[+-]\d\d -?\d\.\d{12}E[+-]\d\d \d\.\d{12}E[+-]\d\d \d\.\d{12}E[+-]\d\d [01]",
                        fs[128 * 5 + 17])
    assert fs[128 * 5 + 17].startswith("17 5 ")      # x then y, x fastest
    avl = open(tmp_path / "av_vels.dat").read().splitlines()
    assert len(avl) == 30 and re.fullmatch(r"7:\t\d\.\d{12}E[+-]\d\d", avl[7])
    import importlib.util
    spec = importlib.util.spec_from_file_location("check_py", os.path.join(ROOT, "check", "check.py"))
    chk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(chk)
    sink = open(os.devnull, "w")
    same = chk.check_files(str(tmp_path / "av_vels.dat"), str(tmp_path / "final_state.dat"),
                           str(tmp_path / "av_vels.dat"), str(tmp_path / "final_state.dat"), out=sink)
    assert same == 0
    # python field maths == oracle field maths (d2q9-bgk.c:789-831)
    u_x, u_y, u, pressure = decks.final_state_fields(p, got, obstacles)
    o = oracle.final_state_f32(p, got, obstacles)
    assert np.array_equal(pressure, o[3]) and np.array_equal(u_x, o[0]) and np.array_equal(u, o[2])


def test_reynolds_formula(oracle):
    p, cells, obstacles = helpers.random_case(64, 32, seed=5)
    got, av = oracle.run_f32(p, cells, obstacles, 4)
    assert decks.calc_reynolds(p, oracle.av_velocity_f32(p, got, obstacles)) == pytest.approx(
        oracle.reynolds_f32(p, got, obstacles), rel=1e-6)


# ---- ring bookkeeping ---------------------------------------------------------------------------

def test_partition_rows_python_matches_c(lbm):
    lib = lbm.cabi.load_library()
    for ny in (2, 3, 7, 128, 1000, 16384, 16385):
        for parts in (1, 2, 3, 4, 8):
            if parts > ny:
                continue
            covered = 0
            for part in range(parts):
                y0, rows = C.c_int(), C.c_int()
                lib.lbm_partition_rows(ny, parts, part, C.byref(y0), C.byref(rows))
                assert (y0.value, rows.value) == lbm.cabi.partition_rows(ny, parts, part)
                assert y0.value == covered and rows.value >= 1
                covered += rows.value
            assert covered == ny
            r, local = ring.accel_owner(ny, parts)
            y0, rows = ring.slab_rows(ny, parts, r)
            assert y0 + local == ny - 2 and 0 <= local < rows


def test_ring_neighbours_and_planes():
    assert ring.neighbours(0, 4) == (3, 1) and ring.neighbours(3, 4) == (2, 0) and ring.neighbours(0, 1) == (0, 0)
    assert ring.UP_PLANES == (2, 5, 6) and ring.DOWN_PLANES == (4, 7, 8)   # kernels.cl:106-112
    assert ring.min_halo_bytes_per_step(16384) == 196608                   # 3 planes x one row
    # what the kernels really store into each neighbour per launch: 2 ghost rows x 9 planes (csrc GHOST = 2)
    assert ring.GHOST_ROWS == 2 and ring.halo_bytes_per_launch(16384) == 2 * 9 * 16384 * 4


def test_combine_av_sums_is_split_invariant(lbm):
    """Sums of fp32 terms held as double-doubles combine to the same fp32 average however the terms
    were split into parts — C and numpy implementations agree bit for bit."""
    lib = lbm.cabi.load_library()
    rng = np.random.default_rng(0)
    nsteps, nterms = 50, 4096
    terms = (rng.random((nsteps, nterms)) * 10.0 ** rng.integers(-9, 2, (nsteps, nterms))).astype(np.float32)
    fci = np.float32(1.0 / 12345.0)

    def dd_sum(block):                      # exact-ish sum of fp32 terms -> (hi, lo)
        hi = np.zeros(nsteps)
        lo = np.zeros(nsteps)
        for j in range(block.shape[1]):
            x = block[:, j].astype(np.float64)
            s = hi + x
            bb = s - hi
            lo += (hi - (s - bb)) + (x - bb)
            hi = s
        s = hi + lo
        return s, lo - (s - hi)

    results = []
    for cuts in ([nterms], [1000, 3096], [7, 1, 2000, 2088], [512] * 8):
        his, los, start = [], [], 0
        for c in cuts:
            h, l = dd_sum(terms[:, start:start + c])
            his.append(h)
            los.append(l)
            start += c
        hi, lo = np.stack(his), np.stack(los)
        av_np = lbm.cabi.combine_av_sums(hi, lo, float(fci))
        av_c = np.empty(nsteps, dtype=np.float32)
        dp = C.POINTER(C.c_double)
        lib.lbm_combine_av_sums(np.ascontiguousarray(hi).ctypes.data_as(dp), np.ascontiguousarray(lo).ctypes.data_as(dp),
                                len(cuts), nsteps, nsteps, C.c_float(fci), av_c.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(helpers.bits(av_np), helpers.bits(av_c))
        results.append(av_np)
    for r in results[1:]:
        assert np.array_equal(helpers.bits(r), helpers.bits(results[0]))
    exact = (terms.astype(np.float64).sum(axis=1) * np.float64(fci)).astype(np.float32)
    np.testing.assert_allclose(results[0], exact, rtol=1e-6)


def test_synthetic_channel_rows_tile_the_deck():
    p, cells, obstacles = decks.synthetic_channel(2048, 3072)
    parts = [decks.synthetic_channel_rows(2048, 3072, y0, rows) for y0, rows in ((0, 1000), (1000, 1048), (2048, 1024))]
    assert np.array_equal(np.concatenate(parts), obstacles)
    assert decks.synthetic_channel_free_cells(2048, 3072) == 2048 * 3072 - int(obstacles.sum())
    assert obstacles[0].all() and obstacles[-1].all() and not obstacles[3072 - 2].any()
    frac = obstacles[1:-1].mean()
    assert 0.003 < frac < 0.005      # 64x64 blocks every 1024 cells: ~0.4 %


def test_reference_arm_line_and_shared_config():
    """bench.py --impl reference runs on host cores alone and prints ONE JSON line whose `config` is the
    very dict the B200 arm prints for the same --gpus (the workload only; arm-specific details live under
    `run`), with the e2e / cpu_baseline keys the bench contract asks of the reference arm."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                           "--warmup", "1", "--ref-rows", "1024"], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["unit"] == "MLUPS"
    assert line["config"] == bench.workload_config(bench.NX, bench.ROWS_PER_GPU, 1)
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["gpu_launches"] == 0

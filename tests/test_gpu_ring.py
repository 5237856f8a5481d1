"""Multi-process ring on real GPUs (one process per GPU, CUDA-IPC peer memory, in-kernel halo
stores + epoch flags).  Needs >= 2 GPUs; skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import opencl_lattice_boltzmann_b200 as lbm
    return lbm.cabi.load_library().lbm_device_count()


@pytest.mark.parametrize("world,ny", [(2, 203), (2, 50), (4, 203), (8, 203)])
@pytest.mark.parametrize("fuse2", [1, 0])
def test_ring_matches_oracle(world, ny, fuse2):
    """ny = 50 on two ranks and 203 on eight give slabs of 25 rows: with 8-row segments the last segment
    holds one row, so the second edge row (rows-2) lives in another block of the two-step kernel."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29510 + world),
           os.path.join(ROOT, "tests", "ring_worker.py"), "--fuse2", str(fuse2), "--ny", str(ny)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "lattice_bit_exact=True" in proc.stdout and "av_bitwise_vs_1gpu=True" in proc.stdout
    assert ("fuse2q_kernel" in proc.stdout) == bool(fuse2)


def test_absent_neighbour_times_out_loudly():
    """One rank of a 2-rank ring never calls lbm_run: the other rank's in-kernel wait gives up after
    wait_timeout_ms and lbm_sync returns an error naming the epoch — no hang."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(ROOT, "tests", "ring_worker.py"), "--absent-rank", "1", "--ny", "64"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "TIMEOUT_REPORTED=True" in proc.stdout


def test_mismatched_plans_are_refused():
    """Ranks that would run different kernels (one forced to the one-step kernel) are refused by lbm_connect."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29532",
           os.path.join(ROOT, "tests", "ring_worker.py"), "--mismatch-rank", "1", "--ny", "64"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "MISMATCH_REFUSED=True" in proc.stdout


@pytest.mark.parametrize("ngpus", [2, 4, 8])
@pytest.mark.parametrize("fuse2", [1, 0])
def test_single_process_multi_device_ring(ngpus, fuse2):
    """The CLI's own multi-GPU path (LBM_NGPUS=N ./d2q9-bgk -> lbm_create(ngpus=N)): ONE process, N
    distinct devices, cudaDeviceEnablePeerAccess, kernels on different GPUs waiting on each other's
    epoch flags.  Lattice bit-exact vs the oracle; av_vels bitwise equal to the single-GPU run.
    (The reference analogue is device selection, d2q9-bgk.c:920-929.)"""
    if _gpu_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    import numpy as np

    import helpers
    import opencl_lattice_boltzmann_b200 as lbm
    import oracle_lib
    p, cells, obstacles = helpers.random_case(512, 203, seed=777, walls=False)
    ref_cells, ref_av = oracle_lib.run_f32(p, cells, obstacles, 13)
    opts = {"cells_per_thread": 4, "fuse2": fuse2, "fuse2_rows": 8}
    with lbm.cabi.Simulation(p, ngpus=ngpus, options=opts) as sim:
        sim.upload(cells, obstacles)
        sim.run(6)
        sim.run(7)
        sim.sync()
        got, av, info = sim.download_cells(), sim.download_av_vels(13), sim.info()
    assert info["nslabs"] == ngpus and ("fuse2q_kernel" in info["kernel_name"]) == bool(fuse2)
    assert np.array_equal(helpers.bits(got), helpers.bits(ref_cells))
    with lbm.cabi.Simulation(p, devices=[0], options={"cells_per_thread": 4, "persistent": 0}) as one:
        one.upload(cells, obstacles)
        one.run(13)
        one.sync()
        av1 = one.download_av_vels(13)
    assert np.array_equal(helpers.bits(av), helpers.bits(av1))

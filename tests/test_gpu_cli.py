"""The C program end to end on a GPU: `d2q9-bgk <params> <obstacles>` writes av_vels.dat and
final_state.dat that pass the reference checker; stdout keeps the reference's five lines
(d2q9-bgk.c:271-275)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "d2q9-bgk")
CHECK = os.path.join(ROOT, "check", "check.py")


def run_deck(tmp_path, name, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([EXE, os.path.join(ROOT, "decks", f"input_{name}.params"),
                        os.path.join(ROOT, "decks", f"obstacles_{name}.dat")], capture_output=True, text=True,
                       cwd=tmp_path, env=e, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


def check(tmp_path, name):
    return subprocess.run([sys.executable, CHECK, f"--ref-av-vels-file={ROOT}/check/{name}.av_vels.dat",
                           f"--ref-final-state-file={ROOT}/check/{name}.final_state.dat",
                           f"--av-vels-file={tmp_path}/av_vels.dat", f"--final-state-file={tmp_path}/final_state.dat"],
                          capture_output=True, text=True)


@pytest.mark.parametrize("name", ["128x128", "128x256", "256x256"])
def test_cli_deck_passes_make_check(tmp_path, name):
    out = run_deck(tmp_path, name)
    lines = out.splitlines()
    assert lines[0] == "==done=="
    assert re.fullmatch(r"Reynolds number:\t\t\d\.\d{12}E[+-]\d\d", lines[1])
    assert re.fullmatch(r"Elapsed time:\t\t\t\d+\.\d{6} \(s\)", lines[2])
    assert re.fullmatch(r"Elapsed user CPU time:\t\t\d+\.\d{6} \(s\)", lines[3])
    assert re.fullmatch(r"Elapsed system CPU time:\t\d+\.\d{6} \(s\)", lines[4])
    r = check(tmp_path, name)
    assert r.returncode == 0 and "Both tests passed!" in r.stdout, r.stdout + r.stderr


def test_cli_device_fields_equal_host_fields_and_slabs(tmp_path):
    """final_state.dat is byte-identical whether the fields come from the device output stage or from
    the reference's host maths, and whether the rows are one slab or three (LBM_NGPUS on one device)."""
    a, b, c = tmp_path / "a", tmp_path / "b", tmp_path / "c"
    for d in (a, b, c):
        d.mkdir()
    run_deck(a, "128x128", {"LBM_QUIET": "1"})
    out = run_deck(b, "128x128", {"LBM_HOST_FIELDS": "1", "LBM_QUIET": "1"})
    assert len(out.splitlines()) == 5           # only the reference's five lines
    run_deck(c, "128x128", {"LBM_NGPUS": "3", "LBM_DEVICES": "0,0,0"})
    ref = open(a / "final_state.dat", "rb").read()
    assert ref == open(b / "final_state.dat", "rb").read()
    assert ref == open(c / "final_state.dat", "rb").read()
    # 1 slab runs the persistent kernel with 1 cell/thread, 3 slabs the step kernel with 4: the lattice is
    # identical, the averages agree to fp32 rounding of the differently grouped segment sums
    av_a = [float(l.split()[1]) for l in open(a / "av_vels.dat")]
    av_c = [float(l.split()[1]) for l in open(c / "av_vels.dat")]
    assert len(av_a) == 40000 and max(abs(x - y) / x for x, y in zip(av_a, av_c)) < 1e-6


def test_cli_binary_and_none_modes(tmp_path):
    run_deck(tmp_path, "128x128", {"LBM_FINAL_STATE": "binary"})
    blob = open(tmp_path / "final_state.dat", "rb").read()
    assert blob.startswith(b"LBMFS1 128 128\n") and len(blob) == len(b"LBMFS1 128 128\n") + 5 * 4 * 128 * 128
    os.remove(tmp_path / "final_state.dat")
    run_deck(tmp_path, "128x128", {"LBM_FINAL_STATE": "none"})
    assert not os.path.exists(tmp_path / "final_state.dat") and os.path.exists(tmp_path / "av_vels.dat")


def _gpu_count():
    import opencl_lattice_boltzmann_b200 as lbm
    return lbm.cabi.load_library().lbm_device_count()


@pytest.mark.parametrize("ngpus", [2, 4, 8])
def test_cli_multi_gpu_on_distinct_devices(tmp_path, ngpus):
    """`LBM_NGPUS=N ./d2q9-bgk` on N DISTINCT GPUs (one process, peer access between devices, kernels on
    different GPUs waiting on each other's epoch flags — the reference's analogue is device selection,
    d2q9-bgk.c:920-929): final_state.dat byte-identical to the single-GPU run and `make check` green."""
    if _gpu_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    one, many = tmp_path / "one", tmp_path / "many"
    one.mkdir()
    many.mkdir()
    run_deck(one, "256x256", {"LBM_QUIET": "1"})
    out = run_deck(many, "256x256", {"LBM_NGPUS": str(ngpus)})
    assert f"x {ngpus} slab(s)" in out
    assert open(one / "final_state.dat", "rb").read() == open(many / "final_state.dat", "rb").read()
    r = check(many, "256x256")
    assert r.returncode == 0 and "Both tests passed!" in r.stdout, r.stdout + r.stderr

"""check/check.py keeps the reference checker's flags, arithmetic, messages and exit codes
(reference check/check.py:1-147)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "check", "check.py")


def write_case(tmp, av, pressure, nx=4):
    av_path, fs_path = os.path.join(tmp, "av.dat"), os.path.join(tmp, "fs.dat")
    with open(av_path, "w") as fp:
        for i, v in enumerate(av):
            fp.write("%d:\t%.12E\n" % (i, v))
    with open(fs_path, "w") as fp:
        for c, pr in enumerate(pressure):
            fp.write("%d %d %.12E %.12E %.12E %.12E %d\n" % (c % nx, c // nx, 0.0, 0.0, 0.0, pr, 0))
    return av_path, fs_path


def run_check(ref, sim, extra=()):
    cmd = [sys.executable, CHECK, f"--ref-av-vels-file={ref[0]}", f"--ref-final-state-file={ref[1]}",
           f"--av-vels-file={sim[0]}", f"--final-state-file={sim[1]}", *extra]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_pass_and_messages(tmp_path):
    ref = write_case(str(tmp_path / ""), [1.0, 2.0, 3.0], np.full(8, 0.0333))
    os.makedirs(tmp_path / "s", exist_ok=True)
    sim = write_case(str(tmp_path / "s"), [1.001, 2.0, 3.0], np.full(8, 0.0333) * 1.002)
    r = run_check(ref, sim)
    assert r.returncode == 0
    out = r.stdout
    assert "Total difference in av_vels : " in out and "Biggest difference (at step 0) : " in out
    assert "Total difference in final_state : " in out and "Biggest difference (at coord (" in out
    assert out.rstrip().endswith("Both tests passed!")
    assert " vs. " in out and "%" in out


def test_fail_on_tolerance_and_custom_tolerance(tmp_path):
    ref = write_case(str(tmp_path), [1.0, 2.0], np.full(4, 1.0))
    os.makedirs(tmp_path / "s", exist_ok=True)
    sim = write_case(str(tmp_path / "s"), [1.0, 2.05], np.full(4, 1.0))
    r = run_check(ref, sim)
    assert r.returncode == 1 and "av_vels failed check" in r.stdout and "Both tests passed!" not in r.stdout
    assert run_check(ref, sim, ["--tolerance", "5"]).returncode == 0
    sim2 = write_case(str(tmp_path / "s"), [1.0, 2.0], [1.0, 1.0, 1.5, 1.0])
    r = run_check(ref, sim2)
    assert r.returncode == 1 and "final state failed check" in r.stdout


def test_nan_fails_and_shape_mismatches(tmp_path):
    ref = write_case(str(tmp_path), [1.0, 2.0], np.full(4, 1.0))
    os.makedirs(tmp_path / "s", exist_ok=True)
    sim = write_case(str(tmp_path / "s"), [1.0, float("nan")], np.full(4, 1.0))
    assert run_check(ref, sim).returncode == 1
    sim = write_case(str(tmp_path / "s"), [1.0, 2.0, 3.0], np.full(4, 1.0))
    r = run_check(ref, sim)
    assert r.returncode == 1 and "Different number of steps in av_vels files" in r.stdout
    sim = write_case(str(tmp_path / "s"), [1.0, 2.0], np.full(4, 1.0), nx=2)
    r = run_check(ref, sim)
    assert r.returncode == 1 and "Final state files coordinates were not the same" in r.stdout


def test_argfile_and_required_flags(tmp_path):
    ref = write_case(str(tmp_path), [1.0], np.full(4, 1.0))
    args = tmp_path / "args.txt"
    args.write_text(f"--ref-av-vels-file={ref[0]}\n--ref-final-state-file={ref[1]}\n"
                    f"--av-vels-file={ref[0]}\n--final-state-file={ref[1]}\n")
    r = subprocess.run([sys.executable, CHECK, f"@{args}"], capture_output=True, text=True)
    assert r.returncode == 0 and "Both tests passed!" in r.stdout
    r = subprocess.run([sys.executable, CHECK], capture_output=True, text=True)
    assert r.returncode == 2  # argparse: required flags missing


def test_goldens_check_against_themselves():
    ref = (os.path.join(ROOT, "check", "128x128.av_vels.dat"), os.path.join(ROOT, "check", "128x128.final_state.dat"))
    r = run_check(ref, ref)
    assert r.returncode == 0 and "= 0%" in r.stdout


def write_velocity_case(tmp, ux, uy, nx=4):
    os.makedirs(tmp, exist_ok=True)
    av_path, fs_path = os.path.join(tmp, "av.dat"), os.path.join(tmp, "fs.dat")
    with open(av_path, "w") as fp:
        fp.write("0:\t1.000000000000E+00\n")
    with open(fs_path, "w") as fp:
        for c, (a, b) in enumerate(zip(ux, uy)):
            fp.write("%d %d %.12E %.12E %.12E %.12E %d\n" % (c % nx, c // nx, a, b, np.hypot(a, b), 0.0333, 0))
    return av_path, fs_path


def test_velocity_extension_is_off_by_default_and_fp32_aware(tmp_path):
    """--velocity-tolerance also checks u_x, u_y, |u|, which the reference ignores; values
    near zero are measured against a thousandth of the largest speed, not against themselves."""
    ux = np.array([0.05, -0.02, 1e-9, 0.0, 0.03, 0.01, -0.04, 0.02])
    uy = np.array([0.01, 0.00, -1e-9, 0.0, 0.02, -0.01, 0.01, 0.03])
    ref = write_velocity_case(str(tmp_path / "r"), ux, uy)
    # a sign flip of a 1e-9 velocity and 0.1 % noise elsewhere: fine for fp32, fatal for a naive relative measure
    sim = write_velocity_case(str(tmp_path / "s"), ux * 1.001 * np.where(np.abs(ux) < 1e-6, -1.0, 1.0), uy * 0.999)
    assert run_check(ref, sim).returncode == 0                       # reference behaviour: velocities not looked at
    r = run_check(ref, sim, ["--velocity-tolerance", "1"])
    assert r.returncode == 0 and "Total difference in final_state velocities" in r.stdout
    bad = write_velocity_case(str(tmp_path / "b"), ux * np.where(np.arange(8) == 4, 1.05, 1.0), uy)
    assert run_check(ref, bad).returncode == 0
    r = run_check(ref, bad, ["--velocity-tolerance", "1"])
    assert r.returncode == 1 and "final state velocities failed check" in r.stdout and "u_x at coord (0,1)" in r.stdout

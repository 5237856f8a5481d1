"""Summarise an ncu report into profiles/: the raw-page CSV of the step kernel (selected metrics)
and profiles/step_kernel_traffic.json (DRAM bytes per launch, read by bench.py's roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep r01 [cells_per_launch]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
    "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
]


def deck_summary(rep, tag, nx, ny, steps):
    """L2-resident deck captures: writes profiles/<tag>_ncu.csv with the L2 / DRAM figures per step."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[0]

    def val(name):
        i = hdr.index(name)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
                 "s": 1.0, "sector": 1.0, "%": 1.0, "inst": 1.0, "register/thread": 1.0}.get(units[i], 1.0)
        return float(r[i]) * scale

    dur = val("gpu__time_duration.sum")
    l2_bytes = 32.0 * (val("lts__t_sectors_srcunit_tex_op_read.sum") + val("lts__t_sectors_srcunit_tex_op_write.sum"))
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    cells = nx * ny
    rec = {
        "kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")],
        "deck": f"{nx}x{ny}", "steps_in_launch": steps, "duration_us": dur * 1e6, "us_per_step": dur * 1e6 / steps,
        "mlups": cells * steps / dur / 1e6,
        "algorithmic_gbs": 72.0 * cells * steps / dur / 1e9,
        "l2_bytes_from_sm": l2_bytes, "l2_gbs": l2_bytes / dur / 1e9,
        "l2_bytes_over_algorithmic": l2_bytes / (72.0 * cells * steps),
        "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct"),
        "lts_throughput_pct_of_peak": val("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_bytes": dram, "dram_gbs": dram / dur / 1e9,
        "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
        "registers_per_thread": val("launch__registers_per_thread"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": val("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}",
    }
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu.json")
    with open(out, "w") as fp:
        json.dump(rec, fp, indent=1)
    print(json.dumps(rec, indent=1))


def main():
    if sys.argv[1] == "--deck":
        return deck_summary(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))
    rep, tag = sys.argv[1], sys.argv[2]
    cells = int(sys.argv[3]) if len(sys.argv) > 3 else 16384 * 16384
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    out = os.path.join(ROOT, "profiles", f"{tag}_step_kernel_ncu_full.csv")
    with open(out, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for i in idx:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    print("wrote", out)

    def col(name):
        i = hdr.index(name)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        return [float(r[i]) * scale for r in data]

    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    per_launch = sum(a + b for a, b in zip(rd, wr)) / len(rd)
    kname = data[0][hdr.index("Kernel Name")]
    # ncu's demangled name -> the name lbm_get_info reports (bench.py's run.kernel)
    import re
    steps = 1
    m = re.search(r"step_kernel<(\d+), (\d+), (\d+), (\d+), (\d+)>", kname)
    name = kname
    if m:
        name = f"step_kernel<V={m.group(1)},hint={m.group(2)},tpb={m.group(3)},tps={m.group(4)},packed={m.group(5)}>"
    m = re.search(r"(fuse2q_kernel|fuse2p_kernel|fuse2_tma_kernel|fuse2_kernel)<(\d+), (\d+)", kname)
    if m:
        steps = 2
        rows = sys.argv[4] if len(sys.argv) > 4 else "128"
        name = f"{m.group(1)}<W={m.group(2)},packed={m.group(3)},rows={rows}>"
    rec = {"kernel": name, "ncu_kernel_name": kname, "cells_per_launch": cells, "steps_per_launch": steps,
           "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
           "dram_bytes_per_launch": per_launch, "algorithmic_bytes_per_launch": 72 * cells * steps,
           "traffic_over_algorithmic": per_launch / (72 * cells * steps), "launches_captured": len(rd),
           "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} ({tag})"}
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        with open(path) as fp:
            table = json.load(fp)
    except Exception:
        table = {}
    table[name] = rec
    with open(path, "w") as fp:
        json.dump(table, fp, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()

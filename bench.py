#!/usr/bin/env python3
"""bench.py — MLUPS of the D2Q9-BGK time step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N GPUs, weak scaling): BASELINE.json configs[4], the synthetic 16384-wide channel of
SURVEY.md §8(d): every rank holds a 16384 x 16384 row slab of a 16384 x (16384*N) grid (channel
walls on the global first/last row, 64x64 solid blocks every 1024 cells), one process per GPU, halo
rows exchanged inside the step kernel through CUDA-IPC peer memory.  A "step" is one LBM time step of
the whole grid.  `value` = cells * K / device time (CUDA events on the launching stream, max over
ranks) with the lattice resident in HBM; `e2e` = the reference's own timed region
(d2q9-bgk.c:196-263: upload + K steps + sync + download) through the C-ABI with pinned HOST buffers.

Beside it, in the same JSON line:
  N = 1   `decks`: the four reference decks (configs[0..3]) at full maxIters through the same C-ABI — us/step,
          MLUPS, the reference's timed region, av_vels against the golden — with the CPU oracle's time for the
          same deck; `cpu_baseline`; `roofline.no_arithmetic_ceiling`.
  N > 1   `parity_check`: BEFORE anything is timed, a 16384 x 512*N ring case run for 3 + 4 steps and compared
          bitwise with the CPU oracle on the whole grid (exit code 3 and nothing timed on a mismatch);
          `strong`: the N = 1 grid (16384 x 16384) split N ways, with the same grid timed on rank 0's GPU alone
          and every av_vels value of the two runs compared bitwise (exit code 3 on a mismatch).

--impl reference times the reference's algorithm on the box's host cores: the OpenMP fp32 CPU
restatement in oracle/ (the reference's OpenCL host cannot be built in this image), each step one
time step of a bounded 16384-wide sample of the same deck.

Prints exactly ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NX = 16384
ROWS_PER_GPU = 16384
BYTES_PER_UPDATE = 72  # 9 fp32 reads + 9 fp32 writes per cell update


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(kernel_name: str, cells: int):
    """DRAM bytes per launch of the step kernel from the committed ncu --set full capture
    (profiles/step_kernel_traffic.json, written by tools/ncu_summary.py); None if not captured for
    this kernel/grid."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as fp:
            rec = json.load(fp).get(kernel_name)
        if rec and rec.get("cells_per_launch") == cells:
            return float(rec["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reason_bits = [], 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # CUDA_VISIBLE_DEVICES may renumber; NVML index follows the physical order
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis:
                ids = [v for v in vis.split(",") if v != ""]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = str(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                try:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        reasons = [n for n, b in {**self.BAD, **self.NOTE}.items() if self.reason_bits & b]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": reasons, "samples": len(self.samples)}


def workload_config(nx, rows, world, scaling="weak"):
    """`config` of the JSON line: the WORKLOAD only, so that both arms (--impl b200 / reference) print the very same
    dict; what is specific to an arm (kernel, decomposition, the CPU arm's bounded sample) goes under `run`."""
    ny_global = rows * world
    return {
        "workload": f"synthetic {nx}x{rows} channel per GPU (BASELINE.json configs[4]): "
                    f"global {nx}x{ny_global}, walls + 64x64 blocks every 1024 cells",
        "nx": nx, "ny_global": ny_global, "rows_per_gpu": rows, "scaling": scaling,
        "l2": f"no flush needed: the two lattices are {2 * 36 * nx * rows / 2**30:.1f} GiB per GPU, "
              f"far larger than the 126 MB L2",
    }


def channel_free_cells(lbm, nx, ny):
    return lbm.decks.synthetic_channel_free_cells(nx, ny)


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle = the checker / CPU baseline; never on the product path)
# ---------------------------------------------------------------------------------------------

def cpu_sample_deck(lbm, ny):
    p, cells, obstacles = lbm.decks.synthetic_channel(NX, ny)
    return p, cells, obstacles


def time_oracle(lbm, ny, steps, warmup, variant="fastest"):
    """Times `steps` oracle time steps (accelerate + timestep, OpenMP over rows) on a NX x ny sample.
    Returns (seconds per step list, threads)."""
    import ctypes as C
    import oracle_lib
    lib = oracle_lib.load(variant)
    lib.oracle_set_num_threads(oracle_lib.host_threads())  # all host cores, whatever OMP_NUM_THREADS says
    p, cells, obstacles = cpu_sample_deck(lbm, ny)
    op = oracle_lib.to_oracle_params(p)
    a = np.ascontiguousarray(cells)
    b = np.empty_like(a)
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    ip = obstacles.ctypes.data_as(C.POINTER(C.c_int))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        lib.oracle_f32_accelerate(C.byref(op), fp(a), ip)
        lib.oracle_f32_timestep(C.byref(op), fp(a), fp(b), ip, 0)
        dt = time.perf_counter() - t0
        a, b = b, a
        if i >= warmup:
            times.append(dt)
    return times, int(lib.oracle_num_threads())


def _oracle_build():
    import oracle_lib
    return oracle_lib.describe("fastest")


def cpu_baseline(lbm, budget_s=15.0):
    """The oracle port on this box's host cores, bounded to ~budget_s of CPU work."""
    ny = 1024
    t1, threads = time_oracle(lbm, ny, 1, 1)
    steps = int(max(3, min(100, budget_s / max(t1[0], 1e-3))))
    times, threads = time_oracle(lbm, ny, steps, 0)
    mlups = NX * ny * len(times) / sum(times) / 1e6
    rec = {"value": round(mlups, 1), "unit": "MLUPS", "cores": threads, "kind": "port",
           "sample": f"{len(times)} time steps of a {NX}x{ny} slab of the same synthetic channel "
                     f"(oracle/lbm_oracle.c fp32, {_oracle_build()}, OMP threads = {threads})"}
    # beside it: the same source with the north star's plain flags and the cell-by-cell row loop (SURVEY 8d's recipe)
    import oracle_lib
    tp, _ = time_oracle(lbm, ny, 1, 1, variant="base")
    nplain = int(max(2, min(20, 5.0 / max(tp[0], 1e-3))))
    tplain, _ = time_oracle(lbm, ny, nplain, 0, variant="base")
    rec["plain_build"] = {"value": round(NX * ny * len(tplain) / sum(tplain) / 1e6, 1), "steps": len(tplain),
                          "build": oracle_lib.describe("base")}
    return rec


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on host cores (oracle port; see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import opencl_lattice_boltzmann_b200 as lbm
    import oracle_lib
    oracle_lib.build_oracle()
    # size the per-step sample so that warmup+steps finish in about two minutes
    t_probe, threads = time_oracle(lbm, 256, 1, 1)
    rate = NX * 256 / t_probe[0]  # cells per second
    budget = 120.0 / max(1, args.steps + args.warmup)
    ny = int(max(64, min(ROWS_PER_GPU, args.ref_rows, (rate * budget / NX) // 64 * 64)))
    times, threads = time_oracle(lbm, ny, args.steps, args.warmup)
    total = sum(times)
    mlups = NX * ny * len(times) / total / 1e6
    sample = (f"each step = one time step (accelerate + fused propagate/rebound/collision/av_vels) of a "
              f"{NX}x{ny} slab of the synthetic channel on {threads} OpenMP threads ({_oracle_build()})")
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * total / len(times), 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(NX, ROWS_PER_GPU, max(1, args.gpus)),
        "run": {"sample": f"the CPU arm steps a bounded {NX}x{ny} sample of the workload", "ny_sample": ny},
        "cpu_baseline": {"value": round(mlups, 1), "unit": "MLUPS", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(mlups, 1), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------

PREFLIGHT_ROWS = 512      # rows per rank of the N-rank parity pre-flight
PREFLIGHT_SPLIT = (3, 4)  # two lbm_run calls: an odd tail (one-step kernel) + a split run


def preflight_parity(lbm, torch, dist, rank, world, local_rank, barrier, emit):
    """N > 1 only, before anything is timed: a seeded 16384 x (512 N) ring case, 7 steps run as 3 + 4
    (odd tail + split run) with the two-step kernel and the long/short segment tiling of the bench
    workload, compared on rank 0 BITWISE with the CPU oracle on the whole grid (per-slab checksums of
    the raw bits) and, for av_vels, bitwise with a single-GPU run.  The oracle is the checker here,
    never the thing measured.  A mismatch ends the bench with a non-zero exit."""
    import oracle_lib
    t0 = time.time()
    nx, R = NX, PREFLIGHT_ROWS
    ny, y0 = R * world, rank * R
    nsteps = sum(PREFLIGHT_SPLIT)
    free_cells = lbm.decks.synthetic_channel_free_cells(nx, ny, walls=False)
    p = lbm.decks.Params(nx=nx, ny=ny, maxIters=nsteps, reynolds_dim=10, density=float(np.float32(0.1)),
                         accel=float(np.float32(0.005)), omega=float(np.float32(1.85)))
    p.free_cells_inv = float(np.float32(1.0) / np.float32(free_cells))

    def slab(y_first, rows):   # open channel (no walls: the periodic wrap carries fluid) + solid blocks
        one = lbm.decks.Params(nx=nx, ny=rows, maxIters=0, reynolds_dim=10, density=p.density, accel=p.accel,
                               omega=p.omega)
        cells = lbm.decks.perturbed_rows(lbm.decks.initial_cells(one), y_first)
        return cells, lbm.decks.synthetic_channel_rows(nx, ny, y_first, rows, walls=False)

    cells, obstacles = slab(y0, R)
    opts = {"fuse2": 1, "fuse2_rows": 32, "fuse2_long": 128}
    sim = lbm.cabi.Simulation(p, slab=(local_rank, rank, world, y0, R), options=opts)
    blobs = [None] * world
    dist.all_gather_object(blobs, sim.export_blob())
    sim.connect(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
    sim.upload(cells, obstacles)
    sim.halo_push()
    barrier()
    for n in PREFLIGHT_SPLIT:
        sim.run(n)
    sim.sync()
    barrier()
    got = sim.download_cells()
    hi, lo = sim.download_av_sums(nsteps)
    kernel = sim.info()["kernel_name"]
    sim.close()
    parts = [None] * world
    dist.all_gather_object(parts, (lbm.decks.bits_checksum(got), hi, lo))
    result = None
    if rank == 0:
        full_cells, full_obst = slab(0, ny)
        oracle_lib.build_oracle()
        lib = oracle_lib.load("fastest")
        lib.oracle_set_num_threads(oracle_lib.host_threads())
        ref, ref_av = oracle_lib.run_f32(p, full_cells, full_obst, nsteps, reference_order=False, variant="fastest")
        same = [parts[r][0] == lbm.decks.bits_checksum(ref[:, r * R:(r + 1) * R, :]) for r in range(world)]
        av = lbm.cabi.combine_av_sums(np.stack([x[1] for x in parts]), np.stack([x[2] for x in parts]), p.free_cells_inv)
        with lbm.cabi.Simulation(p, devices=[local_rank], options={"fuse2": 1, "persistent": 0}) as one:
            one.upload(full_cells, full_obst)
            for n in PREFLIGHT_SPLIT:
                one.run(n)
            one.sync()
            av1 = one.download_av_vels(nsteps)
            one_same = lbm.decks.bits_checksum(one.download_cells()) == lbm.decks.bits_checksum(ref)
        result = {
            "ranks": world, "grid": f"{nx}x{ny}", "rows_per_rank": R, "steps": "+".join(map(str, PREFLIGHT_SPLIT)),
            "kernel": kernel,
            "lattice_bit_exact": bool(all(same)),
            "av_bitwise": bool(np.array_equal(av.view(np.uint32), av1.view(np.uint32))),
            "av_close_to_oracle": bool(np.allclose(av, ref_av, rtol=2e-6, atol=0)),
            "single_gpu_bit_exact": bool(one_same),
            "checker": "oracle/lbm_oracle.c fp32 on the whole grid (per-slab checksums of the raw bits); "
                       "av_vels vs a 1-GPU run of the same grid",
            "seconds": None,
        }
        if not all(same):
            result["slabs_differing"] = [r for r in range(world) if not same[r]]
    ok = torch.tensor([1], device="cuda")
    if rank == 0:
        good = result["lattice_bit_exact"] and result["av_bitwise"] and result["av_close_to_oracle"] and \
            result["single_gpu_bit_exact"]
        ok[0] = 1 if good else 0
        result["seconds"] = round(time.time() - t0, 1)
        log(f"[parity pre-flight] {json.dumps(result)}")
    dist.broadcast(ok, 0)
    if int(ok.item()) != 1:
        if rank == 0:
            emit({"metric": "MLUPS", "value": None, "n_gpus": world, "parity_check": result,
                  "error": "multi-GPU parity pre-flight FAILED: nothing was timed"})
        dist.barrier()
        dist.destroy_process_group()
        raise SystemExit(3)
    return result


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import opencl_lattice_boltzmann_b200 as lbm
    from opencl_lattice_boltzmann_b200 import ring

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} processes (one per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # stdout must carry exactly ONE line (the JSON): libraries that print there (NCCL's version banner at communicator
    # creation, whatever NCCL_DEBUG says) are sent to stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = preflight_parity(lbm, torch, dist, rank, world, local_rank, barrier, emit)

    nx, rows = NX, args.rows_per_gpu
    if args.scaling == "strong":       # total work fixed: the 16384-row grid split into N slabs
        rows = args.rows_per_gpu // world
        if rows * world != args.rows_per_gpu:
            raise SystemExit(f"--scaling strong needs {args.rows_per_gpu} rows divisible by {world} GPUs")
    ny_global = rows * world
    y0 = rank * rows

    # ---- the deck: this rank's slab of the global synthetic channel, in pinned host memory on the GPU's NUMA node ----
    t0 = time.time()

    def channel_params(ny, steps):
        q = lbm.decks.Params(nx=nx, ny=ny, maxIters=steps, reynolds_dim=10, density=float(np.float32(0.1)),
                             accel=float(np.float32(0.005)), omega=float(np.float32(1.85)))
        q.free_cells_inv = float(np.float32(1.0) / np.float32(channel_free_cells(lbm, nx, ny)))
        return q

    p = channel_params(ny_global, args.steps)
    pin_cells = lbm.cabi.PinnedArray((9, rows, nx), np.float32, local_rank)
    pin_obst = lbm.cabi.PinnedArray((rows, nx), np.int32, local_rank)
    pin_out = lbm.cabi.PinnedArray((9, rows, nx), np.float32, local_rank)
    cells_h, obst_h, out_h = pin_cells.array, pin_obst.array, pin_out.array
    init = lbm.decks.initial_cells(lbm.decks.Params(nx=1, ny=1, maxIters=0, reynolds_dim=10, density=p.density,
                                                    accel=p.accel, omega=p.omega))
    for k in range(9):
        cells_h[k].fill(init[k, 0, 0])
    obst_h[...] = lbm.decks.synthetic_channel_rows(nx, ny_global, y0, rows)
    mask_h = lbm.cabi.pack_obstacles(obst_h)        # for the optional packed-mask upload (built once, outside any timing)
    log(f"[rank {rank}] deck {nx}x{rows} of {nx}x{ny_global} built in {time.time() - t0:.1f}s "
        f"(pinned on NUMA node {pin_cells.numa_node})")

    # ---- context, ring ----
    def make_sim(q, r0, nrows):
        s = lbm.cabi.Simulation(q, slab=(local_rank, rank, world, r0, nrows))
        if world > 1:
            bl = [None] * world
            dist.all_gather_object(bl, s.export_blob())
            s.connect(bl[(rank - 1) % world], bl[(rank + 1) % world])
        return s

    sim = make_sim(p, y0, rows)

    def upload(s=None, packed=False, c=None, o=None, m=None):
        s = s or sim
        c = cells_h if c is None else c
        if packed:
            s.upload_packed(c, mask_h if m is None else m)
        else:
            s.upload(c, obst_h if o is None else o)
        if world > 1:
            s.halo_push()
            barrier()

    # ---- value: lattice resident in HBM ----
    upload()
    barrier()
    if args.warmup > 0:
        sim.run(args.warmup)
        sim.sync()
    info0 = sim.info()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms = sim.run_timed(args.steps)
    barrier()
    clocks = sampler.stop()
    info1 = sim.info()
    ms = max_over_ranks(ms)
    cells_total = nx * ny_global
    mlups = cells_total * args.steps / (ms * 1e-3) / 1e6
    launches = int(info1["kernel_launches"] - info0["kernel_launches"])

    # per-rank av_vels -> deterministic cross-GPU combine (rank order, error-free sums)
    nav = args.warmup + args.steps
    hi, lo = sim.download_av_sums(nav)
    parts = gather((hi, lo))
    av = lbm.cabi.combine_av_sums(np.stack([x[0] for x in parts]), np.stack([x[1] for x in parts]), p.free_cells_inv)
    if not np.all(np.isfinite(av)) or not np.all(av[1:] > 0):
        raise SystemExit(f"av_vels look wrong: {av[:5]}")

    # ---- e2e: the reference's timed region through the C-ABI with host buffers ----
    def e2e_once(packed):
        barrier()
        t_0 = time.perf_counter()
        upload(packed=packed)
        t_up = time.perf_counter()
        sim.run(args.steps)
        sim.sync()
        t_loop = time.perf_counter()
        sim.download_cells(out_h)
        sim.download_av_sums(args.steps)
        t_down = time.perf_counter()
        barrier()
        total = max_over_ranks(time.perf_counter() - t_0)
        h2d_rank = cells_h.nbytes + (mask_h.nbytes if packed else obst_h.nbytes)
        d2h_rank = out_h.nbytes + 2 * 8 * args.steps
        per_rank = gather((t_up - t_0, t_loop - t_up, t_down - t_loop))
        return {
            "value": round(cells_total * args.steps / total / 1e6, 1), "unit": "MLUPS",
            "h2d_bytes_per_step": h2d_rank * world // args.steps, "d2h_bytes_per_step": d2h_rank * world // args.steps,
            "seconds": round(total, 4),
            "upload_s": round(max(x[0] for x in per_rank), 4), "loop_s": round(max(x[1] for x in per_rank), 4),
            "download_s": round(max(x[2] for x in per_rank), 4),
            "upload_gbs_per_rank": [round(h2d_rank / x[0] / 1e9, 1) for x in per_rank],
            "download_gbs_per_rank": [round(d2h_rank / x[2] / 1e9, 1) for x in per_rank],
        }

    # three times, the fastest one reported (all listed): on a shared host one PCIe transfer in a handful is several
    # times slower than the rest, and one such outlier would be the whole e2e number of a short run
    e2e_runs = [e2e_once(packed=False) for _ in range(3)]
    e2e = min(e2e_runs, key=lambda r: r["seconds"])
    e2e["runs_s"] = [r["seconds"] for r in e2e_runs]
    e2e["runs_phases_s"] = [[r["upload_s"], r["loop_s"], r["download_s"]] for r in e2e_runs]   # which phase the host noise hit
    checksum = float(out_h[:, ::257, ::263].astype(np.float64).sum())
    if not np.isfinite(checksum):
        raise SystemExit("final state is not finite")
    e2e["host_buffers"] = (f"cudaHostAlloc'ed on the NUMA node of each rank's GPU (lbm_host_alloc_on; nodes "
                           f"{gather(pin_cells.numa_node)})")
    e2e["note"] = ("one upload and one download per run as in the reference (d2q9-bgk.c:196-263), so the bytes per "
                   "step are the run's bytes / steps; PCIe-bound for short runs.  Upload = the reference-shaped "
                   "lbm_upload (int obstacle map, packed on the device)")
    e2e_packed = min((e2e_once(packed=True) for _ in range(2)), key=lambda r: r["seconds"])
    e2e["with_packed_mask_upload"] = {k: e2e_packed[k] for k in ("value", "seconds", "upload_s", "h2d_bytes_per_step")}
    e2e["with_packed_mask_upload"]["note"] = ("same region with lbm_upload_packed: the obstacle map crosses PCIe as the "
                                              "bit mask the kernels use (1/32 of the int map), packed once by the caller")

    # ---- the kernel's own memory ceiling: the same launches without the arithmetic (it wrecks the lattice) ----
    dry = None
    if world == 1 and info1["kernel_name"].startswith("fuse2") and not args.no_dry_run:
        try:
            sim.set_option("fuse2_mode", 3)     # bit 1: same bulk copies, ring traffic, barriers and stores, no collision
            sim.run(4)
            sim.sync()
            nd = max(2, min(args.steps, 100)) & ~1
            ms_dry = sim.run_timed(nd)
            dry = {"mlups": round(cells_total * nd / (ms_dry * 1e-3) / 1e6, 1), "steps": nd,
                   "frac_of_ceiling": round(mlups / (cells_total * nd / (ms_dry * 1e-3) / 1e6), 4),
                   "note": "the two-step kernel with option fuse2_mode bit 1: identical memory traffic and synchronisation, "
                           "arithmetic removed (results are garbage, measured after everything else); "
                           "value / this = how much of the arithmetic the kernel hides behind HBM"}
        except Exception as e:
            dry = {"mlups": None, "note": f"failed: {e}"}
    sim.close()

    # ---- strong scaling of the same 16384 x 16384 grid over the N GPUs (N > 1, weak runs only) ----
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong and args.rows_per_gpu % world == 0:
        rows_s = args.rows_per_gpu // world
        ny_s = args.rows_per_gpu
        q = channel_params(ny_s, args.steps)
        sim_s = make_sim(q, rank * rows_s, rows_s)
        obst_s = np.ascontiguousarray(lbm.decks.synthetic_channel_rows(nx, ny_s, rank * rows_s, rows_s))
        upload(sim_s, c=np.ascontiguousarray(cells_h[:, :rows_s, :]), o=obst_s)   # (the initial state is uniform)
        warm_s = max(2, args.warmup) & ~1
        sim_s.run(warm_s)
        sim_s.sync()
        barrier()
        ms_s = max_over_ranks(sim_s.run_timed(args.steps))
        barrier()
        info_s = sim_s.info()
        sums_s = gather(sim_s.download_av_sums(warm_s + args.steps))
        sim_s.close()
        strong_mlups = nx * ny_s * args.steps / (ms_s * 1e-3) / 1e6
        # the same grid on ONE GPU (rank 0), for the efficiency: the other ranks wait
        n1 = None
        if rank == 0:
            with lbm.cabi.Simulation(q, devices=[local_rank]) as one:
                o1 = np.ascontiguousarray(lbm.decks.synthetic_channel_rows(nx, ny_s, 0, ny_s))
                c1 = cells_h if rows == ny_s else np.ascontiguousarray(np.broadcast_to(cells_h[:, :1, :], (9, ny_s, nx)))
                one.upload(c1, o1)
                one.run(warm_s)
                one.sync()
                n1 = nx * ny_s * args.steps / (one.run_timed(args.steps) * 1e-3) / 1e6
                # every av_vels value of the split run (warm-up + timed steps: each one a sum over the whole lattice)
                # must equal the single-GPU run's bit for bit: the sums are exact, so any halo error would show
                av_split = lbm.cabi.combine_av_sums(np.stack([x[0] for x in sums_s]), np.stack([x[1] for x in sums_s]),
                                                    q.free_cells_inv)
                av_one = one.download_av_vels(warm_s + args.steps)
                av_same = bool(np.array_equal(av_split.view(np.uint32), av_one.view(np.uint32)))
        barrier()
        strong = {"value": round(strong_mlups, 1), "unit": "MLUPS", "ms_per_step": round(ms_s / args.steps, 5),
                  "grid": f"{nx}x{ny_s}", "rows_per_gpu": rows_s, "kernel": info_s["kernel_name"],
                  "n1_value": round(n1, 1) if n1 else None,
                  "efficiency_vs_n1": round(strong_mlups / (world * n1), 4) if n1 else None,
                  "av_vels_bitwise_vs_n1": av_same if rank == 0 else None,
                  "note": f"the {nx}x{ny_s} grid of the N = 1 bench split into {world} row slabs, {args.steps} steps, "
                          f"CUDA events, max over ranks; n1_value = the same grid and step count on rank 0's GPU alone, "
                          f"measured in this run; av_vels_bitwise_vs_n1: all {warm_s + args.steps} av_vels values of the "
                          f"split run equal that run's bit for bit"}
        bad = torch.tensor([1 if (rank == 0 and not av_same) else 0], device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            if rank == 0:
                print(f"bench.py: STRONG-SPLIT MISMATCH: av_vels of the {world}-slab run differ from the single-GPU run "
                      f"({info_s['kernel_name']})", file=sys.stderr, flush=True)
            dist.destroy_process_group()
            sys.exit(3)

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        per_gpu_cells = nx * rows
        achieved = BYTES_PER_UPDATE * per_gpu_cells / (ms / args.steps * 1e-3) / 1e9  # GB/s per GPU
        kname = info1["kernel_name"]
        spl = max(1, int(info1["steps_per_launch"]))          # 2: the two-step (temporal blocking) kernel
        traffic = ncu_traffic_per_launch(kname, per_gpu_cells)
        launch_ms = ms / args.steps * spl
        line = {
            "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 5),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(nx, rows, world, args.scaling),
            "run": {
                "decomposition": f"{world} row slab(s), one process per GPU, in-kernel halo stores over CUDA-IPC peer "
                                 f"memory ({ring.halo_bytes_per_launch(nx)} B per neighbour per launch), epoch flags",
                "kernel": kname,
                "e2e_region": f"upload + {args.steps} steps + sync + download (d2q9-bgk.c:196-263), pinned host buffers",
            },
            "roofline": {
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": traffic,
                "peak_source": peak_src,
                "steps_per_launch": spl,
                "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * per_gpu_cells * spl,
                "launch_ms": round(launch_ms, 5),
                "dram_gbs_actual": (round(traffic / (launch_ms * 1e-3) / 1e9, 1) if traffic else None),
                "dram_frac_of_peak": (round(traffic / (launch_ms * 1e-3) / 1e9 / peak, 4) if traffic else None),
                "traffic_source": "profiles/step_kernel_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                  "`ncu --set full` capture of this kernel on this grid (not re-measured in this run)",
                "no_arithmetic_ceiling": dry,
                "note": "per GPU; launch duration = CUDA-event time of the timed region / launches (includes 1 "
                        "accelerate pre-pass and the av_vels finalize launches). achieved = 72 B x cell updates / "
                        "time, the reference's own accounting." + (
                            " frac > 1 is real work, not skipped work: this kernel advances TWO time steps per pass "
                            "over HBM (step-1 rows live in a shared-memory ring), so its DRAM traffic (`traffic`, "
                            "ncu) is about half the algorithmic bytes and dram_gbs_actual / dram_frac_of_peak are what "
                            "HBM really carries; the lattice stays bit-identical to the one-step kernel and the CPU "
                            "oracle (tests/test_gpu_parity.py, tests/test_gpu_fullsize.py)." if spl > 1 else ""),
            },
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "av_vels_last": float(av[-1]),
        }
        if parity is not None:
            line["parity_check"] = parity
        if strong is not None:
            line["strong"] = strong
    if world == 1 and not args.no_decks:
        try:
            line["decks"] = decks_block(lbm, cpu=not args.no_cpu_baseline)
        except Exception as e:
            line["decks"] = {"error": str(e)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(lbm)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port",
                                        "sample": f"failed: {e}"}
        emit(line)
    for buf in (pin_cells, pin_obst, pin_out):
        buf.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


DECKS = ["128x128", "128x256", "256x256", "1024x1024"]


def decks_block(lbm, cpu=True):
    """The four reference decks (BASELINE.json configs[0..3]) at full maxIters on cuda:0 through the C-ABI:
    loop time by CUDA events, the reference's own timed region (d2q9-bgk.c:196-263: upload + loop + sync +
    download) by wall clock, av_vels against the reference's golden with check.py's measure; beside it the
    CPU oracle (all host cores) for the same deck and step count."""
    import helpers
    out = {}
    for name in DECKS:
        p, cells, obstacles = lbm.decks.load_deck(*lbm.decks.deck_paths(name))
        n = p.maxIters
        cells_out = np.empty_like(cells)
        with lbm.cabi.Simulation(p, devices=[0]) as sim:
            sim.upload(cells, obstacles)
            sim.run(min(2000, n))      # warm-up (module load, clocks)
            sim.sync()
            sim.upload(cells, obstacles)
            ms = sim.run_timed(n)
            av = sim.download_av_vels(n)
            info = sim.info()
            region = None
            for _ in range(2):         # the reference's timed region (best of two: the first pays page faults)
                t0 = time.perf_counter()
                sim.upload(cells, obstacles)
                sim.run(n)
                sim.sync()
                sim.download_cells(cells_out)
                sim.download_av_vels(n)
                dt = time.perf_counter() - t0
                region = dt if region is None else min(region, dt)
        worst, step = helpers.pct_diff(helpers.golden_av_vels(name), av)
        rec = {"steps": n, "kernel": info["kernel_name"], "loop_s": round(ms * 1e-3, 5),
               "us_per_step": round(ms * 1e3 / n, 4), "mlups": round(p.nx * p.ny * n / (ms * 1e-3) / 1e6, 1),
               "timed_region_s": round(region, 5),
               "l2_gbs_algorithmic": round(BYTES_PER_UPDATE * p.nx * p.ny * n / (ms * 1e-3) / 1e9, 1),
               "av_vels_worst_pct_vs_golden": round(worst, 4), "passes_check_py_gate": bool(abs(worst) < 1.0)}
        if cpu:
            import oracle_lib
            lib = oracle_lib.load("fastest")
            threads = oracle_lib.host_threads()
            lib.oracle_set_num_threads(threads)
            t0 = time.perf_counter()
            oracle_lib.run_f32(p, cells, obstacles, n, reference_order=False, variant="fastest")
            cpu_s = time.perf_counter() - t0
            rec["cpu_oracle"] = {"seconds": round(cpu_s, 3), "mlups": round(p.nx * p.ny * n / cpu_s / 1e6, 1),
                                 "cores": threads, "kind": "port", "build": _oracle_build()}
        out[name] = rec
        log(f"[decks] {name}: {json.dumps(rec)}")
    out["note"] = ("device loop time (CUDA events) of the full deck; l2_gbs_algorithmic = 72 B x updates / loop time (the "
                   "lattices sit in L2; ncu's lts__t_bytes for these kernels is in profiles/); timed_region_s = upload + "
                   "loop + sync + download by wall clock, the region the reference times (d2q9-bgk.c:196-263); "
                   "cpu_oracle = oracle/lbm_oracle.c fp32 with OpenMP on all host cores, same deck, same steps")
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000,
                    help="timed LBM time steps (the reference decks run 20 000 - 80 000 per upload)")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU,
                    help="rows of the 16384-wide channel per GPU (default: the BASELINE config)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak (default, the driver's contract): rows-per-gpu rows on every GPU; "
                         "strong: rows-per-gpu rows in total, split over the GPUs")
    ap.add_argument("--ref-rows", type=int, default=ROWS_PER_GPU,
                    help="--impl reference: at most this many rows in the CPU arm's per-step sample (tests use a small one)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="N > 1: skip the ring-vs-oracle bitwise pre-flight (development only)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling block")
    ap.add_argument("--no-decks", action="store_true", help="N = 1: skip the four reference decks")
    ap.add_argument("--no-dry-run", action="store_true",
                    help="skip the no-arithmetic run of the two-step kernel (roofline.no_arithmetic_ceiling)")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())

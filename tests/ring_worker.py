"""One rank of a multi-process ring (run under torchrun, one process per GPU): every rank holds a row
slab of a seeded global case, steps it through the C-ABI with in-kernel halo stores into its
neighbours' CUDA-IPC-mapped ghost rows, and rank 0 compares the gathered lattice with the CPU oracle
on the whole grid (bit-exact) and the combined av_vels."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import helpers
    import opencl_lattice_boltzmann_b200 as lbm

    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--ny", type=int, default=203)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--split", default="11,14")  # two run() calls
    ap.add_argument("--fuse2", type=int, default=-1)
    ap.add_argument("--fuse2-rows", type=int, default=8, help="row-segment length of the two-step kernel; 0: automatic tiling")
    ap.add_argument("--absent-rank", type=int, default=-1,
                    help="this rank never calls lbm_run: its neighbour must report a timeout, not hang")
    ap.add_argument("--mismatch-rank", type=int, default=-1,
                    help="this rank plans the one-step kernel, the others the two-step one: lbm_connect must refuse")
    args = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p, cells, obstacles = helpers.random_case(args.nx, args.ny, seed=4242, walls=False)
    y0, rows = lbm.cabi.partition_rows(args.ny, world, rank)
    fuse2 = args.fuse2
    if args.mismatch_rank >= 0:
        fuse2 = 0 if rank == args.mismatch_rank else 1
    sim = lbm.cabi.Simulation(p, slab=(local, rank, world, y0, rows),
                              options={"cells_per_thread": 4, "fuse2": fuse2, "fuse2_rows": args.fuse2_rows})
    blobs = [None] * world
    dist.all_gather_object(blobs, sim.export_blob())
    if args.mismatch_rank >= 0:
        try:
            sim.connect(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
            refused = False
        except lbm.cabi.LbmError as e:
            refused = "planned other kernels" in str(e)
        flag = torch.tensor([1 if refused else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"MISMATCH_REFUSED={bool(flag.item())}", flush=True)
        sim.close()
        dist.destroy_process_group()
        sys.exit(0 if flag.item() == 1 else 1)
    sim.connect(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
    if args.absent_rank >= 0:
        sim.set_option("wait_timeout_ms", 1500)
        sim.upload(np.ascontiguousarray(cells[:, y0:y0 + rows, :]), np.ascontiguousarray(obstacles[y0:y0 + rows, :]))
        sim.halo_push()
        torch.cuda.synchronize()
        dist.barrier()
        reported = True
        if rank != args.absent_rank:
            try:
                sim.run(8)
                sim.sync()
                reported = False
            except lbm.cabi.LbmError as e:
                reported = "did not reach epoch" in str(e)
            if reported:   # and the context stays failed
                try:
                    sim.run(1)
                    reported = False
                except lbm.cabi.LbmError:
                    pass
        flag = torch.tensor([1 if reported else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"TIMEOUT_REPORTED={bool(flag.item())}", flush=True)
        sim.close()
        dist.destroy_process_group()
        sys.exit(0 if flag.item() == 1 else 1)
    sim.upload(np.ascontiguousarray(cells[:, y0:y0 + rows, :]), np.ascontiguousarray(obstacles[y0:y0 + rows, :]))
    sim.halo_push()
    torch.cuda.synchronize()
    dist.barrier()
    done = 0
    for n in [int(x) for x in args.split.split(",")]:
        sim.run(n)
        done += n
    assert done == args.steps
    sim.sync()
    dist.barrier()
    got = sim.download_cells()
    hi, lo = sim.download_av_sums(args.steps)
    parts = [None] * world
    dist.all_gather_object(parts, (got, hi, lo))
    ok = True
    if rank == 0:
        import oracle_lib
        full = np.concatenate([x[0] for x in parts], axis=1)
        av = lbm.cabi.combine_av_sums(np.stack([x[1] for x in parts]), np.stack([x[2] for x in parts]),
                                      p.free_cells_inv)
        ref_cells, ref_av = oracle_lib.run_f32(p, cells, obstacles, args.steps)
        same = np.array_equal(helpers.bits(full), helpers.bits(ref_cells))
        av_ok = np.allclose(av, ref_av, rtol=2e-6, atol=0)
        # single-slab run on this GPU with the one-step kernel (same cells per thread = same segment sums;
        # the small-deck tile kernel sums per cell instead): av_vels must be bitwise identical to the ring's
        with lbm.cabi.Simulation(p, devices=[local], options={"cells_per_thread": 4, "persistent": 0}) as one:
            one.upload(cells, obstacles)
            one.run(args.steps)
            one.sync()
            av1 = one.download_av_vels(args.steps)
        av_same = np.array_equal(helpers.bits(av), helpers.bits(av1))
        ok = same and av_ok and av_same
        print(f"RING world={world} kernel={sim.info()['kernel_name']} lattice_bit_exact={same} av_close={av_ok} av_bitwise_vs_1gpu={av_same}", flush=True)
    sim.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

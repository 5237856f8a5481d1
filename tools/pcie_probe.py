"""Host<->device copy rates of all ranks at once, normal pinned vs write-combined pinned host memory (development
tool; run under torchrun with one process per GPU).  Each rank copies `--gb` GB up and down several times, all ranks
starting together; prints per-rank GB/s.

    python -m torch.distributed.run --nproc-per-node 4 tools/pcie_probe.py --gb 4"""
import argparse
import ctypes as C
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    n = int(args.gb * 1e9)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for name, flags in (("pinned", 1), ("pinned+write-combined", 1 | 4)):   # cudaHostAllocPortable, | WriteCombined
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
        C.memset(p, 1, n)
        up, down = [], []
        for _ in range(args.reps):
            barrier()
            t0 = time.perf_counter()
            assert rt.cudaMemcpy(C.c_void_p(dev.data_ptr()), p, n, 1) == 0      # H2D
            up.append(n / (time.perf_counter() - t0) / 1e9)
            barrier()
            t0 = time.perf_counter()
            assert rt.cudaMemcpy(p, C.c_void_p(dev.data_ptr()), n, 2) == 0      # D2H
            down.append(n / (time.perf_counter() - t0) / 1e9)
        rt.cudaFreeHost(p)
        res = [None] * world
        if world > 1:
            dist.all_gather_object(res, (max(up), max(down)))
        else:
            res = [(max(up), max(down))]
        if rank == 0:
            print(f"{name:24s} H2D GB/s per rank: {[round(r[0], 1) for r in res]}  D2H: {[round(r[1], 1) for r in res]}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

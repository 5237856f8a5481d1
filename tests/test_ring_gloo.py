"""world_size-2 (and 3) ring on CPU with the gloo backend: the host-side plan for N > 1 — row
partition, which planes travel which way, who owns the accelerate row, the rank-ordered error-free
av_vels combine — exercised with the oracle's slab kernels standing in for the GPU and gloo send/recv
standing in for the in-kernel peer stores.  Result must equal the single-domain oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, nx, ny, nsteps, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    import helpers
    import oracle_lib
    from opencl_lattice_boltzmann_b200 import cabi, ring

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = oracle_lib.load("base")
    lib.oracle_set_num_threads(1)

    p, cells, obstacles = helpers.random_case(nx, ny, seed=99, walls=False)
    op = oracle_lib.to_oracle_params(p)
    y0, rows = ring.slab_rows(ny, world, rank)
    down, up = ring.neighbours(rank, world)
    owner, accel_local = ring.accel_owner(ny, world)
    accel_local = accel_local if owner == rank else -1

    # slab with ghost rows: [9, rows + 2, nx]
    cur = np.zeros((9, rows + 2, nx), dtype=np.float32)
    cur[:, 1:rows + 1] = cells[:, y0:y0 + rows]
    nxt = np.zeros_like(cur)
    obst = np.ascontiguousarray(obstacles[y0:y0 + rows])
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    ip = obst.ctypes.data_as(C.POINTER(C.c_int))

    def exchange(buf):
        """top row of UP_PLANES -> up neighbour's ghost below; bottom row of DOWN_PLANES -> down
        neighbour's ghost above (ring.py); a one-rank ring talks to itself."""
        send_up = torch.from_numpy(np.ascontiguousarray(buf[list(ring.UP_PLANES), rows]))
        send_down = torch.from_numpy(np.ascontiguousarray(buf[list(ring.DOWN_PLANES), 1]))
        recv_below, recv_above = torch.empty_like(send_up), torch.empty_like(send_down)
        if world == 1:
            recv_below.copy_(send_up)
            recv_above.copy_(send_down)
        else:
            reqs = [dist.isend(send_up, up, tag=1), dist.isend(send_down, down, tag=2),
                    dist.irecv(recv_below, down, tag=1), dist.irecv(recv_above, up, tag=2)]
            for r in reqs:
                r.wait()
        buf[list(ring.UP_PLANES), 0] = recv_below.numpy()
        buf[list(ring.DOWN_PLANES), rows + 1] = recv_above.numpy()

    row_sums = np.zeros(rows, dtype=np.float32)
    his, los = [], []
    for _ in range(nsteps):
        lib.oracle_f32_slab_accelerate(C.byref(op), fp(cur), ip, rows, accel_local)
        exchange(cur)
        lib.oracle_f32_slab_timestep(C.byref(op), fp(cur), fp(nxt), ip, rows, fp(row_sums))
        cur, nxt = nxt, cur
        # this rank's sum of speeds as a double-double (what lbm_download_av_sums returns)
        hi = lo = 0.0
        for v in row_sums.astype(np.float64):
            s = hi + v
            bb = s - hi
            lo += (hi - (s - bb)) + (v - bb)
            hi = s
        his.append(hi)
        los.append(lo)

    gathered = [None] * world
    dist.all_gather_object(gathered, (cur[:, 1:rows + 1].copy(), np.array(his), np.array(los)))
    if rank == 0:
        full = np.concatenate([g[0] for g in gathered], axis=1)
        av = cabi.combine_av_sums(np.stack([g[1] for g in gathered]), np.stack([g[2] for g in gathered]),
                                  p.free_cells_inv)
        np.save(os.path.join(out_dir, "cells.npy"), full)
        np.save(os.path.join(out_dir, "av.npy"), av)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ring_equals_single_domain(tmp_path, oracle, world):
    import helpers
    nx, ny, nsteps = 96, 23, 6
    port = free_port()
    mp.spawn(worker, args=(world, port, nx, ny, nsteps, str(tmp_path)), nprocs=world, join=True)
    p, cells, obstacles = helpers.random_case(nx, ny, seed=99, walls=False)
    ref_cells, ref_av = oracle.run_f32(p, cells, obstacles, nsteps, reference_order=False)
    got = np.load(tmp_path / "cells.npy")
    av = np.load(tmp_path / "av.npy")
    assert np.array_equal(helpers.bits(got), helpers.bits(ref_cells))
    np.testing.assert_allclose(av, ref_av, rtol=2e-6, atol=0)
    # a different split gives bitwise the same averages (error-free combine)
    mp.spawn(worker, args=(1, free_port(), nx, ny, nsteps, str(tmp_path)), nprocs=1, join=True)
    assert np.array_equal(helpers.bits(np.load(tmp_path / "av.npy")), helpers.bits(av))

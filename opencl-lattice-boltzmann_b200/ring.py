"""Host-side description of the row-slab ring (pure Python, no device code): who neighbours whom,
which rows each slab holds, who owns the accelerate row, and what crosses NVLink per launch.
Used by bench.py (decomposition report, strong/weak splits, the N-rank parity pre-flight) and by the
CPU ring tests.

The y axis is periodic (kernels.cl:91-93), so N row slabs form a ring: slab r's up neighbour is
(r+1) % N, its down neighbour (r-1) % N.

What the CUDA path exchanges (csrc/lbm_kernels.cuh process_segment, csrc/lbm_fuse2p.cuh phase 2):
every launch — one time step of step_kernel, two of fuse2p_kernel — the warps that compute a slab's
top GHOST_ROWS rows store all nine planes a second time into the up neighbour's ghost rows below its
row 0, and the bottom GHOST_ROWS rows into the down neighbour's ghost rows above its last row
(peer-mapped memory, release/acquire epoch flags; no copy engine, no NCCL).  Two rows and nine planes
because the two-step kernel recomputes the neighbour's edge row of the intermediate step from them;
the one-step kernel writes the same rows so that one- and two-step launches can follow each other.

What ONE time step strictly needs is less — a pull at a slab's bottom row reads planes 2, 5, 6 of the
row below (kernels.cl:106,109,110), a pull at its top row planes 4, 7, 8 of the row above
(kernels.cl:108,111,112): UP_PLANES / DOWN_PLANES.  tests/test_ring_gloo.py exchanges exactly that
minimum between CPU ranks (the oracle's slab kernels standing in for the GPU) to pin the plan itself:
partition, neighbours, accelerate-row owner, rank-ordered av_vels combine.
"""
from __future__ import annotations

from .cabi import partition_rows

GHOST_ROWS = 2           # ghost rows on each side of a slab (csrc/lbm_cuda.cu GHOST)
NPLANES = 9
UP_PLANES = (2, 5, 6)    # the minimum a one-step pull at the up neighbour's bottom row reads
DOWN_PLANES = (4, 7, 8)  # the minimum a one-step pull at the down neighbour's top row reads


def neighbours(rank: int, world: int):
    """(down, up) ranks of `rank` in a ring of `world` slabs."""
    return (rank - 1) % world, (rank + 1) % world


def slab_rows(ny: int, world: int, rank: int):
    """(y0, rows) of slab `rank`: the same even split as lbm_partition_rows."""
    return partition_rows(ny, world, rank)


def accel_owner(ny: int, world: int):
    """(rank, local_row) holding global row ny-2, the row accelerate_flow modifies (kernels.cl:18)."""
    g = ny - 2
    for r in range(world):
        y0, rows = partition_rows(ny, world, r)
        if y0 <= g < y0 + rows:
            return r, g - y0
    raise ValueError("row ny-2 not found")


def halo_bytes_per_launch(nx: int) -> int:
    """Bytes one slab stores into ONE neighbour per kernel launch: GHOST_ROWS rows x nine planes of fp32
    (the same for the one-step and the two-step kernel; the latter advances two time steps with it)."""
    return GHOST_ROWS * NPLANES * nx * 4


def min_halo_bytes_per_step(nx: int) -> int:
    """The least one time step needs per direction: one row of three planes (SURVEY §8e's figure)."""
    return len(UP_PLANES) * nx * 4

"""Host-side description of the row-slab ring (pure Python, no device code).

The y axis is periodic (kernels.cl:91-93), so N row slabs form a ring: slab r's up neighbour is
(r+1) % N, its down neighbour (r-1) % N.  A pull at a slab's bottom row reads planes 2, 5, 6 of the
row below it (kernels.cl:106,109,110); a pull at its top row reads planes 4, 7, 8 of the row above
(kernels.cl:108,111,112).  So each step a slab sends its TOP row of planes 2,5,6 up and its BOTTOM
row of planes 4,7,8 down — in the CUDA path those are extra stores of the step kernel into the
neighbours' ghost rows; this module is the same plan for the callers and the CPU ring tests.
"""
from __future__ import annotations

from .cabi import partition_rows

UP_PLANES = (2, 5, 6)    # sent to the up neighbour, land in its ghost row below row 0
DOWN_PLANES = (4, 7, 8)  # sent to the down neighbour, land in its ghost row above its last row


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


def halo_bytes_per_step(nx: int) -> int:
    """Bytes one slab sends per direction per step (3 planes x one row of fp32)."""
    return 3 * nx * 4

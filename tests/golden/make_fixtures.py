#!/usr/bin/env python3
"""Generates tests/golden/*.npz: small seeded cases stepped by the fp32 CPU oracle (oracle/lbm_oracle.c).
They pin the oracle's arithmetic (operation order, explicit FMA contraction) bit for bit, so a change to it
cannot go unnoticed, and give the GPU tests a fixture that does not need the oracle at run time.

    python tests/golden/make_fixtures.py        # rewrites the fixtures (commit them together with the change)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [  # name, nx, ny, seed, steps, walls
    ("f32_64x32_seed7_10steps", 64, 32, 7, 10, True),
    ("f32_100x37_seed11_7steps_open", 100, 37, 11, 7, False),
    ("f32_1024x12_seed5_6steps_open", 1024, 12, 5, 6, False),
]


def main():
    import helpers
    import oracle_lib
    for name, nx, ny, seed, steps, walls in CASES:
        p, cells, obstacles = helpers.random_case(nx, ny, seed=seed, walls=walls)
        out, av = oracle_lib.run_f32(p, cells, obstacles, steps, reference_order=False)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), nx=nx, ny=ny, seed=seed, steps=steps, walls=walls,
                            cells_bits=out.view(np.uint32), av_bits=av.view(np.uint32))
        print(name, out.shape, float(av[-1]))


if __name__ == "__main__":
    main()

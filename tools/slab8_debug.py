import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import helpers, oracle_lib
import opencl_lattice_boltzmann_b200 as lbm
p, cells, obstacles = helpers.random_case(512, 203, seed=4242, walls=False)
ref_cells, ref_av = oracle_lib.run_f32(p, cells, obstacles, 25)
for ns in (8, 7, 5, 2):
    for split in ((11, 14), (25,)):
        with lbm.cabi.Simulation(p, devices=[0] * ns, options={"cells_per_thread": 4, "fuse2": 1, "fuse2_rows": 8}) as sim:
            sim.upload(cells, obstacles)
            for n in split: sim.run(n)
            sim.sync()
            got = sim.download_cells(); av = sim.download_av_vels(25); info = sim.info()
        diff = helpers.bits(got) != helpers.bits(ref_cells)
        k, y, x = np.nonzero(diff)
        print(ns, split, info["kernel_name"], "ndiff", diff.sum(), "rows", sorted(set(y.tolist()))[:20], "av_ok", np.allclose(av, ref_av, rtol=2e-6), flush=True)

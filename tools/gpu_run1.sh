set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_step or repipelined" > gpurun_out/r7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r7_tests.log
tail -3 gpurun_out/r7_tests.log
timeout 900 python tools/sweep.py --steps 40 --warmup 4 --variants "f2:4:1:64:1:2:0:1,f2:4:1:64:1:2:0:0,f2:4:1:64:1:2:0:2,f2:4:0:64:1:2:0:0,f2:4:1:64:1:2:0:1" > gpurun_out/r7_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r7_sweep.log
cat gpurun_out/r7_sweep.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fuse2p -s 2 -c 1 -o gpurun_out/r7_f2p -f python tools/sweep.py --steps 8 --warmup 4 --variants "f2:4:1:64:1:2:0:1" > gpurun_out/r7_ncu.log 2>&1; echo "ncu rc=$?"

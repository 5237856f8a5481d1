set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_step or repipelined" > gpurun_out/r4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4_tests.log
tail -5 gpurun_out/r4_tests.log
timeout 600 python tools/sweep.py --steps 40 --warmup 4 --variants "f2:4:1:64:1:1,f2:4:1:64:1:2:0:0,f2:4:1:64:1:2:0:1,f2:4:0:64:1:2:0:0,f2:4:1:64:1:2:0:2,f2:4:1:32:1:2:0:0,f2:4:1:128:1:2:0:0,f2:4:1:64:1:2:0:0" > gpurun_out/r4_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r4_sweep.log
cat gpurun_out/r4_sweep.log

set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/sweep.py --steps 60 --warmup 4 --variants "f2:4:1:64:1:2:0:1,f2:4:1:64:1:2:0:5,f2:4:1:64:1:2:0:2,f2:4:1:64:1:2:0:6,f2:4:1:64:1:2:0:5,f2:4:1:64:1:2:0:1" > gpurun_out/r12_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r12_sweep.log
cat gpurun_out/r12_sweep.log

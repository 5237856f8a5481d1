set -x
cd $GRAFT_REPO_ROOT
timeout 900 python tools/sweep.py --steps 60 --warmup 4 --variants "f2:4:1:64:1:2:0:1,f2:4:1:32:1:2:0:1,f2:4:1:48:1:2:0:1,f2:4:1:96:1:2:0:1,f2:4:1:128:1:2:0:1,f2:4:1:111:1:2:0:1,f2:4:1:56:1:2:0:1,f2:4:1:74:1:2:0:1,f2:4:1:64:1:2:0:1" > gpurun_out/r9_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r9_sweep.log
cat gpurun_out/r9_sweep.log

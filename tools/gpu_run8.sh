set -x
cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r10_n8_weak.json 2> gpurun_out/r10_n8_weak.err; echo "weak rc=$?"
cat gpurun_out/r10_n8_weak.json | cut -c1-300

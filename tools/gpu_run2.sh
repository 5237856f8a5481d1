set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ring.py -x -q -m gpu > gpurun_out/r8_ring.log 2>&1; echo "ring rc=$?" >> gpurun_out/r8_ring.log
tail -3 gpurun_out/r8_ring.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r8_n2_weak.json 2> gpurun_out/r8_n2_weak.err; echo "weak rc=$?"
cat gpurun_out/r8_n2_weak.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 200 --warmup 20 --scaling strong --no-cpu-baseline > gpurun_out/r8_n2_strong.json 2> gpurun_out/r8_n2_strong.err; echo "strong rc=$?"
cat gpurun_out/r8_n2_strong.json | cut -c1-400

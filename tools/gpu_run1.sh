set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r6_tests.log
tail -5 gpurun_out/r6_tests.log
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; echo "bench rc=$?"
cat gpurun_out/r6_bench.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fuse2p -s 3 -c 2 -o gpurun_out/r6_f2p -f python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r6_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r6_launches.csv python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r6_ncu2.log 2>&1; echo "ncu2 rc=$?"

# Round-end validation on one B200: the GPU test-suite, smoke(), the default bench line, the reference arm,
# a full ncu capture of the dominant kernel and the launch list of the bench command (outputs in gpurun_out/).
set -x
cd ${GRAFT_REPO_ROOT:-.}
TAG=${1:-final}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fuse2p -s 3 -c 1 -o gpurun_out/${TAG}_f2p -f python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-dry-run > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-dry-run > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu2 rc=$?"

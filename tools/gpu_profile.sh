# ncu evidence for profiles/ on one B200 (run under gpurun): the default bench line first (it must exit 0 without
# ncu), then one full capture of the dominant kernel, the launch list of the same command, and full captures of the
# small-deck kernels.  Outputs in gpurun_out/<TAG>_*.
set -x
cd ${GRAFT_REPO_ROOT:-.}
TAG=${1:-r02}
B="python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-dry-run --no-decks"
timeout 300 $B > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err; echo "bench rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fuse2 -s 3 -c 1 -o gpurun_out/${TAG}_f2p -f $B > gpurun_out/${TAG}_ncu_f2p.log 2>&1; echo "ncu f2p rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
[ "${2:-decks}" = "nodecks" ] && exit 0   # second argument "nodecks": only the two-step kernel
for d in 128x128 256x256 1024x1024; do
  timeout 120 python tools/run_deck.py $d --steps 12000 --chunk 4000 > gpurun_out/${TAG}_deck_$d.log 2>&1; echo "deck $d rc=$?"; cat gpurun_out/${TAG}_deck_$d.log
  timeout 600 ncu --set full --clock-control none -k regex:"tile_kernel|persistent_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_deck_$d -f python tools/run_deck.py $d --steps 12000 --chunk 4000 > gpurun_out/${TAG}_ncu_deck_$d.log 2>&1; echo "ncu deck $d rc=$?"
done

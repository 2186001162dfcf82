#!/bin/bash
# Round-2 baseline: default bench line (configs included), ncu launch list of the same command, ncu --set full of the
# three C2 passes, warm per-pass DRAM traffic.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -u
T=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_$T.json; echo
S="python bench.py --steps 1 --warmup 3 --cols 64 --no-e2e --no-cpu --no-pass-events --no-configs"
$S > gpurun_out/plain_$T.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $S > gpurun_out/ncu_launch_$T.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_ -s 288 -c 96 --csv --log-file gpurun_out/warm_traffic_$T.csv $S > gpurun_out/ncu_warm_$T.log 2>&1
echo "warm traffic rc=$?"
python profiles/tools/ncu_c2_small.py > gpurun_out/plain_small_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 6 -f -o gpurun_out/c2_$T python profiles/tools/ncu_c2_small.py > gpurun_out/ncu_full_$T.log 2>&1
echo "ncu full rc=$?"
echo done

#!/bin/bash
# Round-2 probe 1: validate the refactor (pytest -m gpu), then A/B the cheap pass-level changes on the headline config.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -3 gpurun_out/pytest_r2a.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "== $*" >> gpurun_out/probe1.log; $B "$@" >> gpurun_out/probe1.log 2>&1; }
run
run --opt prune_half=0
run --opt l2_persist_mb=64
run --opt l2_persist_mb=96
run --opt l2_persist_mb=48
run --opt l2_persist_mb=64 --opt overlap_streams=3
run --opt l2_persist_mb=32 --no-overlap
run --no-overlap
run --opt l2_persist_mb=64 --group-mb 80
grep -E "^==|value" gpurun_out/probe1.log | cut -c1-400
python profiles/tools/config_bench.py c6 c1 c3 c4 c5 > gpurun_out/config_r2a.jsonl 2> gpurun_out/config_r2a.err; cut -c1-300 gpurun_out/config_r2a.jsonl
# warm DRAM traffic per pass with and without the persisting window (short batch: 64 columns)
S="python bench.py --steps 1 --warmup 3 --cols 64 --no-e2e --no-cpu --no-pass-events"
$S > gpurun_out/plain_s.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_ -s 288 -c 96 --csv --log-file gpurun_out/warm_traffic_r2_base.csv $S > gpurun_out/ncu_s1.log 2>&1
$S --opt l2_persist_mb=64 > gpurun_out/plain_s2.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_ -s 288 -c 96 --csv --log-file gpurun_out/warm_traffic_r2_l2p.csv $S --opt l2_persist_mb=64 > gpurun_out/ncu_s2.log 2>&1
echo done

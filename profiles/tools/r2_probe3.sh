#!/bin/bash
# Round-2 probe 3: tests, cgs debug, straggler scan, Levinson sx A/B, default L2 window bench, memcheck.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2c.log
tail -12 gpurun_out/pytest_r2c.log
python profiles/tools/dbg_cgs.py > gpurun_out/dbg_cgs.log 2>&1; cat gpurun_out/dbg_cgs.log | cut -c1-600
for o in "lev_sx=0" "lev_sx=1" "lev_sx=2" "lev_sx=3"; do echo "== c4 $o" >> gpurun_out/config_r2c.jsonl; python profiles/tools/config_bench.py c4 $o >> gpurun_out/config_r2c.jsonl 2>> gpurun_out/config_r2c.err; done
python profiles/tools/config_bench.py c6scan >> gpurun_out/config_r2c.jsonl 2>> gpurun_out/config_r2c.err
cut -c1-420 gpurun_out/config_r2c.jsonl
python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_r2c.json; echo
python profiles/tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && \
compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/tools/sanitize_small.py > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitize_plain.log; tail -6 gpurun_out/sanitizer_memcheck_r02.log
echo done

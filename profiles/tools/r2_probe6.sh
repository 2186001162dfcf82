#!/bin/bash
# Round-2 probe 6: full GPU test-suite with recorded parity measurements, default bench line (configs block after the
# L2 carve-out release).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_r2g.log | tail -3
grep -E "^(FAILED|ERROR)|AssertionError" gpurun_out/pytest_r2g.log | head -20
python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r02b.json; echo
echo done

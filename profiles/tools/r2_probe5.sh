#!/bin/bash
# Round-2 probe 5: fused C+A launches (tests + headline A/B), Levinson with the two-step reciprocal, gated L2 window.
set -u
mkdir -p gpurun_out
O=gpurun_out/probe5.log; : > $O
python -m pytest tests/test_gpu_ca_fuse.py -m gpu -x -q > gpurun_out/pytest_r2f.log 2>&1; echo "pytest(ca_fuse) rc=$?" | tee -a $O
tail -15 gpurun_out/pytest_r2f.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs"
run() { echo "== $*" >> $O; $B "$@" >> $O 2>> gpurun_out/probe5.err; }
run
run --opt ca_fuse=1
run --opt ca_fuse=2
run --opt ca_fuse=2 --opt overlap_streams=3
run --opt ca_fuse=1 --no-overlap
run
run --opt ca_fuse=2
grep -E "^==|value" $O | cut -c1-330
echo "== c3 c4 c5 (defaults)" >> $O; python profiles/tools/config_bench.py c3 c4 c5 >> $O 2>> gpurun_out/probe5.err
echo "== c3 ca_fuse=1" >> $O; python profiles/tools/config_bench.py c3 ca_fuse=1 >> $O 2>> gpurun_out/probe5.err
tail -5 $O | cut -c1-330
echo done

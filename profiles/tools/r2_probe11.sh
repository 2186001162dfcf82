#!/bin/bash
# Round-2 probe 11: programmatic dependent launch of passes B / C.
set -u
mkdir -p gpurun_out
O=gpurun_out/probe11.log; : > $O
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?" | tee -a $O
grep -E "passed|failed|error" gpurun_out/pytest_r2k.log | tail -3
for o in "pdl=0" "pdl=1"; do echo "== c1 c5 $o" >> $O; python profiles/tools/config_bench.py c1 c5 $o >> $O 2>> gpurun_out/probe11.err; done
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs --no-pass-events"
for o in "pdl=0" "pdl=2" "pdl=0" "pdl=2"; do echo "== c2 $o" >> $O; $B --opt $o 2>> gpurun_out/probe11.err | cut -c1-160 >> $O; done
for o in "pdl=0" "pdl=2"; do echo "== c3 $o" >> $O; python profiles/tools/config_bench.py c3 $o >> $O 2>> gpurun_out/probe11.err; done
cut -c1-250 $O
echo done

#!/bin/bash
# Round-2 probe 9: multi-warp generator-form Levinson (512 < n <= 2048).
set -u
mkdir -p gpurun_out
O=gpurun_out/probe9.log; : > $O
python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "levinson or durbin or c4 or trench" > gpurun_out/pytest_r2j.log 2>&1; echo "pytest(levinson) rc=$?" | tee -a $O
tail -4 gpurun_out/pytest_r2j.log
for o in "lev_gen=1" "lev_gen=0"; do echo "== c4s $o" >> $O; python profiles/tools/config_bench.py c4s $o 2>> gpurun_out/probe9.err | grep -E "n=(768|1024|2048)" >> $O; done
cut -c1-260 $O
echo done

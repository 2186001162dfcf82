#!/bin/bash
# Round-2 probe 7: software-pipelined generator-form Levinson (tests + C4), C3 single-stream with the window.
set -u
mkdir -p gpurun_out
O=gpurun_out/probe7.log; : > $O
python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "levinson or durbin or c4 or trench" > gpurun_out/pytest_r2h.log 2>&1; echo "pytest(levinson) rc=$?" | tee -a $O
tail -4 gpurun_out/pytest_r2h.log
echo "== c4 c4s" >> $O; python profiles/tools/config_bench.py c4 c4s >> $O 2>> gpurun_out/probe7.err
for o in "overlap_streams=1" "overlap_streams=1 l2_persist_mb=72" "overlap_streams=2 l2_persist_mb=0"; do echo "== c3 $o" >> $O; python profiles/tools/config_bench.py c3 $o >> $O 2>> gpurun_out/probe7.err; done
cut -c1-300 $O
python profiles/tools/ncu_lev_small.py > gpurun_out/plain_lev.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_levinson -s 2 -c 2 -f -o gpurun_out/lev_gen2_r02 python profiles/tools/ncu_lev_small.py > gpurun_out/ncu_lev2.log 2>&1
echo "ncu lev rc=$?"
echo done

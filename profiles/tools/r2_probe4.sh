#!/bin/bash
# Round-2 probe 4: generator-form Levinson (tests, C4 A/B, small orders, ncu), L2 window A/B on the configs whose
# workspaces exceed L2.
set -u
mkdir -p gpurun_out
O=gpurun_out/probe4.log; : > $O
python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "levinson or c4" > gpurun_out/pytest_r2d.log 2>&1; echo "pytest(levinson) rc=$?" | tee -a $O
tail -5 gpurun_out/pytest_r2d.log
for o in "lev_gen=1" "lev_gen=0"; do echo "== c4 $o" >> $O; python profiles/tools/config_bench.py c4 c4s c4m $o >> $O 2>> gpurun_out/probe4.err; done
for o in "l2_persist_mb=0" "l2_persist_mb=64"; do echo "== c3 c5 $o" >> $O; python profiles/tools/config_bench.py c3 c5 $o >> $O 2>> gpurun_out/probe4.err; done
cut -c1-330 $O
python profiles/tools/ncu_lev_small.py > gpurun_out/plain_lev.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_levinson -s 2 -c 2 -f -o gpurun_out/lev_gen_r02 python profiles/tools/ncu_lev_small.py > gpurun_out/ncu_lev.log 2>&1
echo "ncu lev rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2e.log 2>&1; echo "pytest(all) rc=$?"; tail -4 gpurun_out/pytest_r2e.log
echo done

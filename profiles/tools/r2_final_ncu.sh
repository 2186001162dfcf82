#!/bin/bash
# ncu --set full of the config 3 and config 5a passes and the Levinson kernel on the final build (summaries -> profiles/)
set -u
mkdir -p gpurun_out
for c in c3 c5a; do
  S="python profiles/tools/ncu_${c}_small.py"
  $S > gpurun_out/plain_${c}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 3 -f -o gpurun_out/${c}_r02f $S > gpurun_out/ncu_${c}_f.log 2>&1
  echo "$c rc=$?"
done
python profiles/tools/ncu_lev_small.py > gpurun_out/plain_lev.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_levinson -s 2 -c 2 -f -o gpurun_out/lev_r02f python profiles/tools/ncu_lev_small.py > gpurun_out/ncu_lev_f.log 2>&1
echo "lev rc=$?"
ls -la gpurun_out/*_r02f.ncu-rep

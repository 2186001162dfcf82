#!/bin/bash
# Round-2 probe 10: persisting window on the spectrum for products whose workspaces do not fit (config 3).
set -u
mkdir -p gpurun_out
O=gpurun_out/probe10.log; : > $O
for o in "l2_spec_window=0" "l2_spec_window=1" "l2_spec_window=1 l2_persist_mb=72" "l2_spec_window=1 l2_persist_mb=80" "l2_spec_window=1 overlap_streams=3"; do echo "== c3 $o" >> $O; python profiles/tools/config_bench.py c3 $o >> $O 2>> gpurun_out/probe10.err; done
S="python profiles/tools/ncu_c3_small.py"
$S > gpurun_out/plain_c3.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_ -s 12 -c 24 --csv --log-file gpurun_out/warm_traffic_c3.csv $S > gpurun_out/ncu_c3.log 2>&1
$S l2_spec_window=1 > gpurun_out/plain_c3b.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_ -s 12 -c 24 --csv --log-file gpurun_out/warm_traffic_c3_specwin.csv $S l2_spec_window=1 > gpurun_out/ncu_c3b.log 2>&1
cut -c1-240 $O
echo done

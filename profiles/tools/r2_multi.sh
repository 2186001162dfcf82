#!/bin/bash
# Multi-GPU check of the driver's launch line: N ranks, weak + strong scaling legs, sharded configs, all-rank parity.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | head -25 >> gpurun_out/topo_n$N.txt; (numactl -H || true) >> gpurun_out/topo_n$N.txt 2>&1
for g in 0 1; do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node; done >> gpurun_out/topo_n$N.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; echo "bench N=$N rc=$?"
head -c 1500 gpurun_out/bench_r02_n$N.json; echo
tail -5 gpurun_out/bench_r02_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_r02_ref_n$N.json 2> gpurun_out/bench_r02_ref_n$N.err; echo "ref N=$N rc=$?"
head -c 600 gpurun_out/bench_r02_ref_n$N.json; echo
echo done

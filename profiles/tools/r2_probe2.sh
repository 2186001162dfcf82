#!/bin/bash
# Round-2 probe 2: tests (TMA variants, two-level batched Krylov, truth fixtures), the new bench line, TMA / L2 A-B.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -15 gpurun_out/pytest_r2b.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "== $*" >> gpurun_out/probe2.log; $B "$@" >> gpurun_out/probe2.log 2>&1; }
run
run --opt tma=1
run --opt tma=2
run --opt tma=3
run --opt l2_persist_mb=64
run --opt l2_persist_mb=64 --opt tma=3
run
run --opt l2_persist_mb=64
grep -E "^==|value" gpurun_out/probe2.log | cut -c1-330
for o in "" "tma=1" "tma=2" "tma=3"; do echo "== config_bench c3 c5 $o" >> gpurun_out/config_r2b.jsonl; python profiles/tools/config_bench.py c3 c5 $o >> gpurun_out/config_r2b.jsonl 2>> gpurun_out/config_r2b.err; done
python profiles/tools/config_bench.py c6 >> gpurun_out/config_r2b.jsonl 2>> gpurun_out/config_r2b.err
cut -c1-330 gpurun_out/config_r2b.jsonl
python - <<'PY'
import ctypes, torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
for name, attr in (("MaxPersistingL2CacheSize", 108), ("MaxAccessPolicyWindowSize", 109), ("L2CacheSize", 38)):
    v = ctypes.c_int()
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0)
    print(name, v.value)
PY
echo done

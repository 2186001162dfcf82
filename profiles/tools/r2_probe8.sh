#!/bin/bash
# Round-2 probe 8: full GPU suite (any-order Levinson block kernel, gated window), C3 stream / group sweeps.
set -u
mkdir -p gpurun_out
O=gpurun_out/probe8.log; : > $O
rm -f gpurun_out/parity_measured.jsonl
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?" | tee -a $O
grep -E "passed|failed|error" gpurun_out/pytest_r2i.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2i.log | head -20
for o in "overlap_streams=2" "overlap_streams=3" "overlap_streams=4" "overlap_streams=2 group_bytes=134217728" "overlap_streams=3 group_bytes=134217728"; do echo "== c3 $o" >> $O; python profiles/tools/config_bench.py c3 $o >> $O 2>> gpurun_out/probe8.err; done
echo "== c5 overlap_streams=3" >> $O; python profiles/tools/config_bench.py c5 overlap_streams=3 >> $O 2>> gpurun_out/probe8.err
echo "== c4s large" >> $O; python - >> $O 2>> gpurun_out/probe8.err <<'PY'
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import toeplitz_b200 as tb
from profiles.tools.config_bench import timed
for n, nsys in ((3000, 4096), (5000, 2048), (8000, 1024), (20000, 296)):
    R = (torch.rand((nsys, n), dtype=torch.float64, device="cuda") / n).t()
    B = torch.randn((nsys, n), dtype=torch.float64, device="cuda").t()
    X = tb.colmajor(n, nsys)
    t = timed(lambda: tb.levinson_(X, R[: n - 1, :], B), 2, warm=1)
    print(json.dumps({"config": "levinson block kernel n=%d x %d" % (n, nsys), "solves_per_s": nsys / t, "fma_tflops": 2 * 3.0 * n * n * nsys / t / 1e12}), flush=True)
PY
cut -c1-260 $O
echo done

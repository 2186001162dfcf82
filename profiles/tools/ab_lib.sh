#!/bin/bash
# A/B of an experiment build (TB200_LIB) against the default library: headline + passes, alternating, then the configs.
V=${1:?variant .so path}
mkdir -p gpurun_out
for r in 1 2 3; do
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', round(d['value']), {k: round(v/max(d['pass_cnt'][k],1)*1e3,2) for k,v in d['pass_ms'].items() if d['pass_cnt'][k]})"
  TB200_LIB=$V python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant', round(d['value']), {k: round(v/max(d['pass_cnt'][k],1)*1e3,2) for k,v in d['pass_ms'].items() if d['pass_cnt'][k]})"
done
for lib in "" $V; do
  TB200_LIB=$lib python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']
print('$lib' or 'default', 'C1', round(c['C1']['mul_us_c_abi'],2), 'C3', round(c['C3']['ldiv_solves_per_s']), 'C5a', round(c['C5a']['ms'],4), 'C5b', round(c['C5b']['ms_per_iteration'],4), 'parity', d['parity']['max_over_ranks_rel_err_vs_oracle'], c['C3']['parity_rel_err_vs_oracle'], c['C5a']['parity_rel_err_vs_dense_rows'])"
done

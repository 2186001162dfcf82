#!/bin/bash
# A/B of a runtime option on the headline + configs: ab_opt.sh name off on
N=$1; A=$2; B=$3
for r in 1 2 3; do
 for v in $A $B; do
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --opt $N=$v 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$N=$v', round(d['value']), {k: round(x/max(d['pass_cnt'][k],1)*1e3,2) for k,x in d['pass_ms'].items() if d['pass_cnt'][k]}, d['parity'])"
 done
done

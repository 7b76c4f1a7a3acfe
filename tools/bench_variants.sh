#!/bin/bash
# perf iteration helper (GPU box): runs bench.py under several env-var variants, one summary line each.
# usage: tools/bench_variants.sh "VAR=1 VAR2=x" "VAR=2" ...   ("-" = no extra env)
for v in "$@"; do
  [ "$v" = "-" ] && v=""
  echo "== variant: ${v:-default}"
  env $v python bench.py --steps 5 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.1f Mcells/s  ms/step %.2f  sweep frac %.3f  step frac %.3f' % (d['value'], d['ms_per_step'], r['frac'], r['step_frac']))
print({k:(round(v['ms_per_launch'],3), round(v['gbs'])) for k,v in r['per_direction'].items()})
"
done

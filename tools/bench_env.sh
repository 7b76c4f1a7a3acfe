#!/usr/bin/env bash
# A/B of environment switches (one B200): tools/bench_env.sh "<label>:<VAR=value ...>" ...   (512^3 fp64, ms per sweep launch)
for spec in "$@"; do
  label="${spec%%:*}"; envs="${spec#*:}"
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2> /tmp/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['roofline']['per_direction']
print('%-28s x %.3f  y %.3f  z %.3f  ms/step %.2f  value %.1f' % ('$label', p['sweep_x']['ms_per_launch'], p['sweep_y']['ms_per_launch'], p['sweep_z']['ms_per_launch'], d['ms_per_step'], d['value']))
" || tail -3 /tmp/bench_err.log
done

#!/bin/bash
# quick perf iteration helper (GPU box): fast-mode parity subset + 512^3 bench summary
python -m pytest tests -m gpu -x -q -k "fast or full_size or batched" 2>&1 | tail -4
python bench.py --steps 5 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.1f Mcells/s  ms/step %.2f  sweep frac %.3f  step frac %.3f' % (d['value'], d['ms_per_step'], r['frac'], r['step_frac']))
print({k:(round(v['ms_per_launch'],3), round(v['gbs'])) for k,v in r['per_direction'].items()})
print({k:round(v,2) for k,v in r['kernel_ms'].items()})
"

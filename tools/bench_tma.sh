#!/usr/bin/env bash
# A/B of the sweep-kernel variants at 512^3 (one B200): prints ms per sweep launch per direction and ms per step.
#   tools/bench_tma.sh [extra bench.py args]
set -u
run() { # label, env...
  local label="$1"; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline "${EXTRA[@]}" 2> /tmp/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
p=d['roofline']['per_direction']
print('%-34s x %.3f (%s)  y %.3f (%s)  z %.3f  ms/step %.2f  value %.1f  residual %.9e' % ('$label', p['sweep_x']['ms_per_launch'], p['sweep_x'].get('kernel','?'), p['sweep_y']['ms_per_launch'], p['sweep_y'].get('kernel','?'), p['sweep_z']['ms_per_launch'], d['ms_per_step'], d['value'], d['residual']))
" || tail -5 /tmp/bench_err.log
}
EXTRA=("$@")
run "direct (round 1)" CMC_TMA=
run "tma xy, no hints" CMC_TMA=xy CMC_TMA_HINTS=0
run "tma xy, L2 hints on copies" CMC_TMA=xy CMC_TMA_HINTS=1
run "tma xy, streaming stores" CMC_TMA=xy CMC_TMA_HINTS=2
run "tma xy, hints + streaming stores" CMC_TMA=xy CMC_TMA_HINTS=3

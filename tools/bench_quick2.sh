#!/usr/bin/env bash
# one line per variant: ms per sweep launch per direction, ms per step (512^3 fp64 unless extra args say otherwise)
run() { local label="$1"; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline "${EXTRA[@]}" 2>/tmp/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['roofline']['per_direction']
print('%-30s' % '$label', {k:(round(v['ms_per_launch'],3), v['kernel']) for k,v in p.items()}, round(d['ms_per_step'],2), d['residual'])" || tail -3 /tmp/bench_err.log; }
EXTRA=("$@")

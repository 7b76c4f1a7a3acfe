#!/usr/bin/env bash
# A/B of the tile shapes of the TMA sweep kernel (one B200): ms per sweep launch per direction and ms per step.
#   tools/bench_shapes.sh <size> <shape>...      shape = <lines>x<CTAs per tile>, e.g. 8x1 16x2
set -u
SIZE="$1"; shift
for shape in "$@"; do
  CMC_TMA_SHAPE="$shape" timeout 300 python bench.py --size "$SIZE" --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2> /tmp/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
p=d['roofline']['per_direction']
print('%-6s %-6s x %.3f (%.3f)  y %.3f (%.3f)  z %.3f  ms/step %.2f  value %.1f  residual %.9e' % ('$SIZE', '$shape', p['sweep_x']['ms_per_launch'], p['sweep_x']['frac'], p['sweep_y']['ms_per_launch'], p['sweep_y']['frac'], p['sweep_z']['ms_per_launch'], d['ms_per_step'], d['value'], d['residual']))
" || tail -5 /tmp/bench_err.log
done

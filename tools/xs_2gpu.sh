#!/usr/bin/env bash
# two-GPU check of the one-pass slab-coupled x-sweep: dist tests, then bench with and without it
mkdir -p gpurun_out
N="${1:-2}"
if [ "${SKIP_TESTS:-0}" != 1 ]; then
(timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8) > gpurun_out/r3b_dist_tests.log 2>&1; cat gpurun_out/r3b_dist_tests.log
fi
run() { # label env...
  local label="$1"; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r3b_bench_${N}gpu_$label.json 2> gpurun_out/r3b_bench_${N}gpu_$label.err
  tail -1 gpurun_out/r3b_bench_${N}gpu_$label.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['roofline']['per_direction']
print('$label', {k:(round(v['ms_per_launch'],3), v['kernel'][:28]) for k,v in p.items()}, 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1), 'residual', d['residual'], d['checksums']['u']['sum'])
" || tail -5 gpurun_out/r3b_bench_${N}gpu_$label.err
}
run xs16 CMC_XS=1
run xs8 CMC_XS=1 CMC_XS_NL=8
run twopass CMC_XS=0

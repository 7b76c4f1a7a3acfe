#!/usr/bin/env bash
# round-end evidence on one B200: GPU suite, bench lines (default, weak, 2d, reference arm), then the ncu launch list
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r3f_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3f_tests.log
python bench.py > gpurun_out/r3f_bench_1gpu.json 2> gpurun_out/r3f_bench_1gpu.err; echo "bench exit $?"
python bench.py --scaling weak --no-cpu-baseline > gpurun_out/r3f_bench_weak_1gpu.json 2> gpurun_out/r3f_weak.err; echo "weak exit $?"
python bench.py --config 2d > gpurun_out/r3f_bench_2d.json 2> gpurun_out/r3f_2d.err; echo "2d exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3f_bench_reference.json 2> gpurun_out/r3f_ref.err; echo "ref exit $?"
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r3f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3f_launches_512_fp64.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r3f_ncu.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/r3f_tests.log; cat gpurun_out/r3f_bench_1gpu.json gpurun_out/r3f_bench_weak_1gpu.json gpurun_out/r3f_bench_2d.json gpurun_out/r3f_bench_reference.json

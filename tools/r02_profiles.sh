#!/bin/bash
# one GPU box: the records of round 2 (launch list, ncu full of the sweep kernel in the 4-stage launch shape, bench lines of every config)
set -x
mkdir -p gpurun_out/r2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_C2_4stages.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --check 0 --fill-steps --rebalance 0 --no-micro > gpurun_out/r2/launches_bench.log 2>&1
QPG_SWEEP_CTAS=37 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -f -o gpurun_out/r2/prof_sweep37 \
    python tools/profile_sweep.py 1045 128 > gpurun_out/r2/prof_sweep37.log 2>&1
for c in C1 C3 C4 C5; do
    timeout 600 python bench.py --config $c > gpurun_out/r2/bench_$c.json 2> gpurun_out/r2/bench_$c.err
    tail -c 300 gpurun_out/r2/bench_$c.err
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_C2_1gpu.json 2> gpurun_out/r2/bench_C2_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/bench_C2_reference.json 2> gpurun_out/r2/bench_C2_reference.err
ls -la gpurun_out/r2

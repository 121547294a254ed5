#!/bin/bash
# bench (both arms) + ncu launch list + one full ncu capture of the dominant kernel
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline --no-generation > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline --no-generation > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -3 gpurun_out/ncu_launches.log
python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline --no-generation > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_bwd_kernel -s 8 -c 1 -o gpurun_out/prof_scan_bwd \
    python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline --no-generation > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/

#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, kernel table, ncu launch list of a bench step.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag] [pytest-args...]
tag=${1:-run}; shift
tests=${@:-tests}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/nvidia_smi.txt 2>&1
timeout 900 python -m pytest $tests -m gpu -x -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log
AG2V_BENCH_BREAKDOWN=$out/breakdown.txt timeout 600 python bench.py > $out/bench_1gpu.json 2> $out/bench_1gpu.err; echo "bench rc=$?" >> $out/bench_1gpu.err
timeout 300 python bench.py --generator-only --no-cpu-baseline > $out/bench_1gpu_generator_only.json 2> $out/bench_1gpu_generator_only.err
if [ -z "$SKIP_REF" ]; then timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; fi
timeout 300 python tools/profile_step.py > $out/step_profile.txt 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base -c 40000 --csv --log-file $out/launches_raw.csv \
    python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
python tools/launch_list.py $out/launches_raw.csv > $out/launches_step.csv 2> $out/launch_list.err
gzip -f $out/launches_raw.csv
tail -3 $out/pytest_gpu.log; tail -2 $out/smoke.log; cat $out/bench_1gpu.json; tail -2 $out/bench_1gpu.err; cat $out/bench_1gpu_generator_only.json; tail -2 $out/bench_1gpu_generator_only.err; cat $out/bench_reference.json; head -12 $out/launches_step.csv

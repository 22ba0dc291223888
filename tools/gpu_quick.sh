#!/bin/bash
# Quick GPU-box visit: selected tests + the default bench + kernel table.  bash tools/gpu_quick.sh tag [pytest args]
tag=${1:-run}; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest "$@" -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
AG2V_BENCH_BREAKDOWN=$out/breakdown.txt timeout 500 python bench.py ${BENCH_ARGS:-} > $out/bench_1gpu.json 2> $out/bench_1gpu.err; echo "bench rc=$?" >> $out/bench_1gpu.err
timeout 200 python tools/profile_step.py > $out/step_profile.txt 2>&1
grep -v "^\.*$" $out/pytest_gpu.log | tail -25 | cut -c1-600; cat $out/bench_1gpu.json; tail -3 $out/bench_1gpu.err; head -45 $out/step_profile.txt | cut -c1-150

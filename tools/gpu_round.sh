#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, kernel table.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag] [pytest-args...]
tag=${1:-run}; shift
tests=${@:-tests}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/nvidia_smi.txt 2>&1
timeout 900 python -m pytest $tests -m gpu -x -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log
AG2V_BENCH_BREAKDOWN=$out/breakdown.txt timeout 600 python bench.py > $out/bench_1gpu.json 2> $out/bench_1gpu.err; echo "bench rc=$?" >> $out/bench_1gpu.err
if [ -z "$SKIP_REF" ]; then timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; fi
timeout 300 python tools/profile_step.py > $out/step_profile.txt 2>&1
grep -v "^\.*$" $out/pytest_gpu.log | tail -8 | cut -c1-300; tail -2 $out/smoke.log; cat $out/bench_1gpu.json; tail -2 $out/bench_1gpu.err; cat $out/bench_reference.json

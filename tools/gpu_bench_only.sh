#!/bin/bash
# Short GPU-box visit: full-iteration bench, kernel table, ncu launch list of one step.  bash tools/gpu_bench_only.sh [tag]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
AG2V_BENCH_BREAKDOWN=$out/breakdown.txt timeout 600 python bench.py > $out/bench_1gpu.json 2> $out/bench_1gpu.err; echo "bench rc=$?" >> $out/bench_1gpu.err
timeout 300 python tools/profile_step.py > $out/step_profile.txt 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base -c 40000 --csv --log-file $out/launches_raw.csv \
    python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
python tools/launch_list.py $out/launches_raw.csv > $out/launches_step.csv 2> $out/launch_list.err
gzip -f $out/launches_raw.csv
cat $out/bench_1gpu.json; tail -3 $out/bench_1gpu.err; head -30 $out/step_profile.txt | cut -c1-160

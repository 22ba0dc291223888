#!/bin/bash
# ncu evidence for the bench: (1) launch list of one full train iteration of `python bench.py` (eager mode so
# every kernel is a launch; earlier steps skipped), (2) --set full captures of the K7 / K2 kernels.
tag=${1:-ncu}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base --launch-skip ${SKIP:-17000} -c ${COUNT:-8500} \
    --csv --log-file $out/launches_raw.csv python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline --skip-peak > $out/bench_under_ncu.log 2>&1
python tools/launch_list.py $out/launches_raw.csv > $out/launches_step.csv 2> $out/launch_list.err
gzip -f $out/launches_raw.csv
head -30 $out/launches_step.csv | cut -c1-150; tail -3 $out/launch_list.err; ls -la $out

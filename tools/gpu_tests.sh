#!/bin/bash
# Selected GPU tests under a hard timeout (a hung kernel must not hold the box): bash tools/gpu_tests.sh tag secs tests...
tag=$1; secs=$2; shift 2
out=gpurun_out/$tag
mkdir -p $out
timeout -s KILL $secs python -m pytest "$@" -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
grep -v "^\.*$" $out/pytest_gpu.log | tail -40 | cut -c1-700

N=${1:-2}
mkdir -p gpurun_out/r2pr
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2pr/$name.json 2> gpurun_out/r2pr/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2pr/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), round(d["ms_per_step"],3))
except Exception as e:
    print("$name failed", e)
PY
tail -2 gpurun_out/r2pr/$name.err | cut -c1-200
}
run prio A=1
run noprio AG2V_NCCL_HIGH_PRIORITY=0
run prio_static AG2V_TC_STATIC=1

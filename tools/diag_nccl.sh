N=${1:-2}
mkdir -p gpurun_out/r2nc
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2nc/$name.json 2> gpurun_out/r2nc/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2nc/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), round(d["ms_per_step"],3))
except Exception as e:
    print("$name failed", e)
PY
}
run simple NCCL_PROTO=Simple
run simple_c16 NCCL_PROTO=Simple NCCL_MAX_CTAS=16
run simple_c8 NCCL_PROTO=Simple NCCL_MAX_CTAS=8
run ll128 NCCL_PROTO=LL128
run default A=1

# A/B on one box at N GPUs: gradient all-reduce by NCCL (default) vs copy engines over IPC windows (K8b).
N=${1:-2}
mkdir -p gpurun_out/r2ce
run() { name=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2ce/$name.json 2> gpurun_out/r2ce/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2ce/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), round(d["ms_per_step"],3), d["impl_detail"].get("gradient_allreduce"))
except Exception as e:
    print("$name failed", e)
PY
  tail -3 gpurun_out/r2ce/$name.err | cut -c1-250
}
run ce AG2V_GRAD_ALLREDUCE=ce
run nccl AG2V_GRAD_ALLREDUCE=nccl

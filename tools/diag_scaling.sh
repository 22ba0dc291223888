# Diagnosis of the data-parallel overhead at N GPUs (default 2): full step, without SyncBN, without the gradient
# all-reduce, without both, with NCCL carrying the SyncBN sums, and one GPU alone.  Numbers are ms per step.
N=${1:-2}
mkdir -p gpurun_out/r2d
run() { # name, env...
  name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2d/$name.json 2> gpurun_out/r2d/$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2d/$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],1), round(d["ms_per_step"],3), d["impl_detail"].get("syncbn_collective",{}).get("collective"), d["impl_detail"].get("DIAGNOSIS_ONLY"))
PY
}
run full AG2V_DIAG=
run nosyncbn AG2V_DIAG=nosyncbn
run nograd AG2V_DIAG=nograd
run none AG2V_DIAG=nosyncbn,nograd
run full_nccl AG2V_PEER_SYNCBN=0
python bench.py --steps 20 --warmup 5 --skip-peak --no-cpu-baseline > gpurun_out/r2d/one.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r2d/one.json").read().strip().splitlines()[-1])
print("one gpu", round(d["value"],1), round(d["ms_per_step"],3), d["gpu_launches"])
PY

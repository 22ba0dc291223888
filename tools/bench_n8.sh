# N=8: the full step against the diagnosis variants (no SyncBN / no gradient all-reduce / neither).  ms per step.
mkdir -p gpurun_out/r2n8b
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --skip-peak > gpurun_out/r2n8b/$name.json 2> gpurun_out/r2n8b/$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2n8b/$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],1), round(d["ms_per_step"],3), d["impl_detail"].get("syncbn_collective",{}).get("collective"), d["clocks"])
PY
}
for v in "$@"; do
  case $v in
    full) run full AG2V_DIAG= ;;
    nccl) run full_nccl AG2V_PEER_SYNCBN=0 ;;
    none) run none AG2V_DIAG=nosyncbn,nograd ;;
    nosyncbn) run nosyncbn AG2V_DIAG=nosyncbn ;;
    nograd) run nograd AG2V_DIAG=nograd ;;
  esac
done

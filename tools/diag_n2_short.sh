N=2
mkdir -p gpurun_out/r2d
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2d/$name.json 2> gpurun_out/r2d/$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2d/$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],1), round(d["ms_per_step"],3), d["impl_detail"].get("syncbn_collective",{}).get("collective"), d["impl_detail"].get("DIAGNOSIS_ONLY"))
PY
}
run full AG2V_DIAG=
run nograd AG2V_DIAG=nograd
run none AG2V_DIAG=nosyncbn,nograd
run full2 AG2V_DIAG=

mkdir -p gpurun_out/r2d
run() { # name, env...
  name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --skip-peak > gpurun_out/r2d/$name.json 2> gpurun_out/r2d/$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2d/$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],1), round(d["ms_per_step"],3), d["impl_detail"].get("syncbn_collective",{}).get("collective"), d["impl_detail"].get("DIAGNOSIS_ONLY"))
PY
}
run base AG2V_DIAG=
run ctas2 NCCL_MAX_CTAS=2
run ctas4 NCCL_MAX_CTAS=4
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
grep -h "channels\|nChannels\|NVLS" gpurun_out/r2d/base.err | head -5

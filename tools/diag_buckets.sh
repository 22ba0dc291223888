N=${1:-2}
mkdir -p gpurun_out/r2bk
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2bk/$name.json 2> gpurun_out/r2bk/$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2bk/$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],1), round(d["ms_per_step"],3), d["impl_detail"].get("gradient_buckets"))
PY
}
run default A=1
run no_overlap AG2V_GRAD_OVERLAP=0
run mb8 AG2V_BUCKET_MB=8
run mb400 AG2V_BUCKET_MB=400

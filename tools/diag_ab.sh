# A/B on one box at N GPUs: dynamic vs static conv tile schedule, per-weight vs batched spectral-norm sigma gradient.
# (When profiles/r02_scaling/n2_r2ab_* were taken the per-weight nodes were the default and AG2V_SN_BATCHED_GRAD=1 selected
# the batched node; the defaults were then set from these runs.)
N=${1:-2}
mkdir -p gpurun_out/r2ab
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --skip-peak > gpurun_out/r2ab/$name.json 2> gpurun_out/r2ab/$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ab/$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],1), round(d["ms_per_step"],3))
PY
}
run dyn_perw AG2V_SN_PER_WEIGHT_GRAD=1
run static_perw AG2V_TC_STATIC=1 AG2V_SN_PER_WEIGHT_GRAD=1
run dyn_batched A=1
run static_batched AG2V_TC_STATIC=1
run dyn_perw_2 AG2V_SN_PER_WEIGHT_GRAD=1
run static_perw_2 AG2V_TC_STATIC=1 AG2V_SN_PER_WEIGHT_GRAD=1
run none AG2V_DIAG=nosyncbn,nograd

#!/bin/bash
# multi-GPU bench exactly as the driver launches it (torchrun, one rank per GPU): cfg2 (default) then cfg3
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt
for WL in cfg2 cfg3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $WL \
      > $OUT/scale_${TAG}_${WL}_n$N.json 2> $OUT/scale_${TAG}_${WL}_n$N.err
  echo "$WL exit $?"; python - <<PY
import json
d=json.loads(open("$OUT/scale_${TAG}_${WL}_n$N.json").read().strip().splitlines()[-1])
print("$WL n=%d value %.0f pairs/s, %.2f ms/step, e2e %.0f, conv3 %.0f TF/s, clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]))
print(d["recall"])
PY
  tail -2 $OUT/scale_${TAG}_${WL}_n$N.err
done

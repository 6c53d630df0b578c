#!/bin/bash
# multi-GPU bench exactly as the driver launches it (torchrun, one rank per GPU)
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/scale_${TAG}_n$N.json 2> $OUT/scale_${TAG}_n$N.err
echo "exit $?"; cat $OUT/scale_${TAG}_n$N.json; tail -5 $OUT/scale_${TAG}_n$N.err

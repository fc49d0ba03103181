#!/bin/bash
# every named particle shape (0 NGP, 1 tophat/CIC, 2 triangle/TSC, 3 spline3) on the 2-D headline grid at
# 1, 2, 4, 8 GPUs (weak scaling, 1e8 particles per GPU); under gpurun --gpus 8
out=gpurun_out/shapes_scaling.log
rm -f $out
for s in 0 1 2 3; do
  for n in 1 2 4 8; do
    args="--gpus $n --workload cfg2 --shape $s --steps 10 --warmup 3 --no-cpu --no-e2e --no-variants"
    if [ $n -eq 1 ]; then timeout 200 python bench.py $args >> $out 2>&1
    else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
           --master-port $((29700 + 10 * s + n)) bench.py $args >> $out 2>&1; fi
    echo "shape $s N=$n rc=$?"
  done
done

#!/bin/bash
# scaling series on one box: N = 1, 2, 4, 8 back to back (weak scaling, 1e8 particles per GPU)
# usage (under gpurun --gpus 8): bash profiles/run_scaling.sh [extra bench args]
out=gpurun_out/scaling.log
rm -f $out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu "$@" >> $out 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu "$@" >> $out 2>&1
  fi
  echo "N=$n rc=$?"
done
grep '^{' $out | cut -c1-200

#!/bin/bash
# all five BASELINE.json configurations at their full single-GPU sizes, device-resident timing
out=gpurun_out/configs_full.log
rm -f $out
run() { echo "### $*" >> $out; timeout 400 python bench.py --no-e2e --no-cpu --no-variants "$@" >> $out 2>&1; }
run --workload cfg1 --steps 50 --warmup 10
run --workload cfg1 --n 1e8 --steps 10
run --workload cfg2 --steps 20
run --workload cfg2 --order sorted --steps 20
run --workload cfg2 --order cellsorted --steps 20
run --workload cfg2 --n 1e9 --steps 5
run --workload cfg3 --steps 10
run --workload cfg3 --order cellsorted --steps 10
run --workload cfg4 --steps 5
run --workload cfg4 --order cellsorted --steps 5
run --workload cfg5 --steps 5
run --workload cfg5 --method rows --steps 5
run --workload cfg5 --method sorted --steps 3
run --workload cfg5 --order cellsorted --steps 5
grep -c '^{' $out

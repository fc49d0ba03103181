#!/bin/bash
# end of round 2: the kernels added last (bulk-copied wide runs for CIC, FIXED direct kernel inside a slab pass)
out=gpurun_out/r02/ncu3; mkdir -p $out
run() {
  python profiles/diag/perf_matrix.py --n $4 --steps 1 $3 > $out/plain_$1.log 2>&1 || { echo "plain run of $1 failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $5 -c 1 -o $out/ncu_$1 python profiles/diag/perf_matrix.py --n $4 --steps 1 $3 > $out/ncu_$1.log 2>&1
  tail -1 $out/ncu_$1.log | cut -c1-120
  python profiles/ncu_summary.py report $out/ncu_$1.ncu-rep > $out/ncu_full_$1.txt 2>&1
  rm -f $out/ncu_$1.ncu-rep
}
run runs_bulk_wide_cic_cellsorted "k_deposit_runs_bulk_wide" cfg2:cellsorted:runs_wide:1 1e8 3
run fixed_direct_slab_cfg3 "k_fixed_deposit" cfg3:random:fixed 1.25e8 5

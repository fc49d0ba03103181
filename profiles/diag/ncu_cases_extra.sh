#!/bin/bash
out=gpurun_out/r02/ncu2; mkdir -p $out
run() {
  python profiles/diag/perf_matrix.py --n $4 --steps 1 $3 > $out/plain_$1.log 2>&1 || { echo "plain run of $1 failed"; return; }
  ncu --set full --clock-control none -k regex:$2 -s 3 -c 1 -o $out/ncu_$1 python profiles/diag/perf_matrix.py --n $4 --steps 1 $3 > $out/ncu_$1.log 2>&1
  tail -1 $out/ncu_$1.log | cut -c1-120
  python profiles/ncu_summary.py report $out/ncu_$1.ncu-rep > $out/ncu_full_$1.txt 2>&1
  rm -f $out/ncu_$1.ncu-rep
}
run runs_wide_cic_cellsorted "k_deposit_runs$" cfg2:cellsorted:runs_wide:1 1e8
run runs_wide_spline3_cellsorted "k_deposit_runs$" cfg2:cellsorted:runs_wide:3 1e8

#!/bin/bash
# One `ncu --set full` capture per dominant kernel (run on the GPU box, after the same commands exited 0 without ncu).
#   profiles/diag/ncu_cases.sh <outdir>
# Each case: perf_matrix.py at 1e8 particles, 1 timed step after 3 warm-ups; the 4th launch of the kernel is captured.
set -u
out=$1; mkdir -p $out
run() {   # tag, kernel regex, case
  python profiles/diag/perf_matrix.py --n 1e8 --steps 1 $3 > $out/plain_$1.log 2>&1 || { echo "plain run of $1 failed"; return; }
  ncu --set full --clock-control none -k regex:$2 -s 3 -c 1 -o $out/ncu_$1 python profiles/diag/perf_matrix.py --n 1e8 --steps 1 $3 > $out/ncu_$1.log 2>&1
  tail -1 $out/ncu_$1.log | cut -c1-160
  # the report embeds the library's cubins (tens of MB): keep the summaries, drop the report
  python profiles/ncu_summary.py report $out/ncu_$1.ncu-rep > $out/ncu_full_$1.txt 2>&1
  ncu -i $out/ncu_$1.ncu-rep --page raw --csv > $out/ncu_raw_$1.csv 2>/dev/null
  rm -f $out/ncu_$1.ncu-rep
}
run quads_cfg2 k_deposit_quads cfg2:random:rows
run runs_bulk_cfg2 k_deposit_runs_bulk cfg2:cellsorted:runs
run tile_fx_cfg5 k_deposit_tile_fx cfg5:random:tiled
run tile_fx_cfg2_xsorted k_deposit_tile_fx cfg2:sorted:tiled
run small_cfg1 k_deposit_small cfg1:random:tiled
run quads_cfg3 k_deposit_quads cfg3:random:rows
run rows_cfg4 k_deposit_rows cfg4:random:rows

#!/usr/bin/env python
"""Summarise ncu output (run here, no GPU needed).

    python profiles/ncu_summary.py launches gpurun_out/launches.csv     # per-kernel time shares
    python profiles/ncu_summary.py report  gpurun_out/prof.ncu-rep      # key metrics of a --set full capture
"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
    'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed_op_shared_atom.sum', 'smsp__inst_executed_op_shared_atom_dot_cas.sum',
    'smsp__inst_executed_op_global_red.sum', 'smsp__inst_executed.sum', 'smsp__inst_issued.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.sum',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def read_csv_text(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = read_csv_text(open(path).read())
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, total = {}, 0.0
    for r in rows[1:]:
        name = r[kn]
        short = name.split("<")[0].split("(")[0].strip()
        if "ptg::" in name:
            short = name[name.index("ptg::"):].split("(")[0]
        ns = float(r[mv].replace(",", ""))
        a = agg.setdefault(short[-70:], [0.0, 0])
        a[0] += ns
        a[1] += 1
        total += ns
    print("%-72s %5s %10s %10s %7s" % ("kernel", "n", "total ms", "avg ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-72s %5d %10.3f %10.3f %6.1f%%" % (k, v[1], v[0] / 1e6, v[0] / 1e6 / v[1], 100 * v[0] / total))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = read_csv_text(out)
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("== %s  (ID %s)" % (r[kn][:110], r[0]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("   %-82s %16s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])

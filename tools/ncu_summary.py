#!/usr/bin/env python3
"""Summarise an .ncu-rep (needs ncu on PATH): key metrics per launch, and optionally write a CSV of the
metrics that the judge looks at into profiles/.   tools/ncu_summary.py in.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    out = []
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append([w, units[i]] + [row[i] for row in rows])
    names = [row[hdr.index('Kernel Name')][:60] for row in rows]
    print("kernels:", names)
    for o in out:
        print("  %-82s %-10s %s" % (o[0], o[1], "  ".join(o[2:])))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + names)
            w.writerows(out)


if __name__ == "__main__":
    main()

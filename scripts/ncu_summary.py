#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel) into a small CSV: `python scripts/ncu_summary.py rep.ncu-rep [out.csv]`."""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_xu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'sm__inst_executed_pipe_fp64_op_dfma.sum',
        'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, universal_newlines=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = list(KEEP) + [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    lines = ["metric,unit,value"]
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            lines.append('%s,%s,"%s"' % (k, units[i], vals[i]))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()

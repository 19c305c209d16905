"""Print the metrics we track from an .ncu-rep (needs only the ncu CLI, no GPU): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{rows[1][i]}]: " + ", ".join(r[i] for r in rows[2:]))
stall = [(h, [r[hdr.index(h)] for r in rows[2:]]) for h in hdr
         if 'issue_stalled' in h and (h.endswith('_per_warp_active.pct') or h.endswith('_per_issue_active.ratio'))]
print("## stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active.ratio), top 8")
for h, v in sorted(stall, key=lambda t: -float(t[1][0] or 0))[:8]:
    print(h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__warp_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)

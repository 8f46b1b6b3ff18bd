"""Key figures of every kernel in an ncu report:  python tools/summarize_ncu.py rep.ncu-rep [label]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print(f"kernel: {name}   {' '.join(sys.argv[2:])}")
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:72s} {vals[i]:>20s} {units[i]}")
    stalls = sorted(((float(vals[i] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                     for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h), reverse=True)
    print("  top stalls (cycles per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
    print()

"""Turn an `ncu --set full --import-source on` report of the tau-leap kernel into the committed text summary and
profiles/roofline_inputs.json.   python tools/summarize_profile.py <report.ncu-rep> <n_traj> <out.txt> [label]"""
import collections
import csv
import json
import os
import subprocess
import sys

rep, n, out_txt = sys.argv[1], int(sys.argv[2]), sys.argv[3]
label = sys.argv[4] if len(sys.argv) > 4 else ""
top = int(sys.argv[5]) if len(sys.argv) > 5 else 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_elapsed.avg']
d, lines = {}, []
for k in keep:
    if k in hdr:
        i = hdr.index(k)
        lines.append(f"{k:70s} {vals[i]:>22s} {units[i]}")
        d[k] = (vals[i], units[i])
stalls = sorted(((float(vals[i] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                 for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h),
                reverse=True)


def to_bytes(x):
    v, u = x
    return float(v) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]


inst = float(d['smsp__inst_executed.sum'][0])
dram = to_bytes(d['dram__bytes_read.sum']) + to_bytes(d['dram__bytes_write.sum'])
txt = f"ncu --set full --clock-control none --import-source on -k regex:tau_leap -s 2 -c 1 : python tools/profile_kernel.py {n} 3\n"
txt += f"kernel: {vals[hdr.index('Kernel Name')]}  {label}\n{n} influenza trajectories with per-trajectory beta / f_ud draws, full season, int32 output\n\n"
txt += "\n".join(lines) + "\n\nwarp stall reasons (cycles per issued instruction):\n"
txt += "\n".join(f"  {name:28s} {v:.3f}" for v, name in stalls[:8]) + "\n"
txt += f"\nderived: {inst / n:.0f} warp-instructions per trajectory ({inst / n * 32 / 7728:.0f} thread-slots per binomial variate incl. all model arithmetic and output);\n"
txt += f"         DRAM traffic {dram / n:.0f} B per trajectory vs 18 188 B algorithmic (output 18 144 B + 40 B drawn parameters + 4 B status)\n"

# per source line
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
sections, cur = [], None
for row in csv.reader(src.split('\n')):
    if not row:
        continue
    if row[0] == "File Path":
        cur = {'file': row[1], 'rows': [], 'hdr': None}
        sections.append(cur)
    elif row[0] == "Line No":
        cur['hdr'] = row
    elif row[0] != "Function Name" and cur is not None and cur['hdr']:
        cur['rows'].append(row)
tot, agg = 0.0, []
for s in sections:
    h = s['hdr']
    iL, iS, iI, iT = h.index('Line No'), h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
    per = collections.OrderedDict()
    for r in s['rows']:
        try:
            ln, ie, te = int(r[iL]), float(r[iI] or 0), float(r[iT] or 0)
        except ValueError:
            continue
        e = per.setdefault(ln, [r[iS], 0.0, 0.0])
        e[1] += ie
        e[2] += te
    for ln, (text, ie, te) in per.items():
        agg.append((ie, te, os.path.basename(s['file']), ln, text.strip()[:100]))
        tot += ie
agg.sort(reverse=True)
txt += "\nper source line (share of executed warp instructions, average active lanes):\n"
for ie, te, f, ln, text in agg[:top]:
    txt += f"{ie / tot:6.2%} lanes={te / max(ie, 1):5.1f} {f}:{ln}  {text}\n"
open(out_txt, "w").write(txt)
json.dump({"source": f"{out_txt} (ncu --set full, {n} trajectories, default kernel)", "warp_inst_per_traj": inst / n,
           "dram_bytes_per_traj": dram / n, "dram_bytes_per_launch": dram / n * 1_000_000,
           "issue_active_pct": float(d['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]),
           "lanes_per_inst": float(d['smsp__thread_inst_executed_per_inst_executed.ratio'][0])},
          open(os.path.join(os.path.dirname(out_txt), "roofline_inputs.json"), "w"), indent=1)
print(txt)

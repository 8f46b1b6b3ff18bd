"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> profiles/*_launches_summary.txt
(python tools/summarize_launches.py <launches.csv> <out.txt> "<command that was profiled>")."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= iV:
        continue
    v = float(r[iV].replace(",", ""))
    ms = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}[r[iU]]
    e = agg.setdefault(r[iK], [0, 0.0])
    e[0] += 1
    e[1] += ms
tot = sum(v[1] for v in agg.values())
lines = [f"{ms:10.2f} ms {100 * ms / tot:6.2f} %  x{c:4d}  {k[:140]}" for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
txt = (f"ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : {sys.argv[3]}\n"
       "(per-launch times under ncu are cold-cache and serialised: the SHARES are what agrees with the bench, not the absolutes)\n"
       "the first 400 launches = 3 warm-up + 2 timed device-resident steps (5 x tau_leap_kernel on 10^6 trajectories), then the e2e\n"
       "section (per chunk of 65 536 trajectories: 1 tau_leap_kernel + 7 draws_major_kernel)\n\n" + "\n".join(lines)
       + f"\n\ntotal {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")
open(sys.argv[2], "w").write(txt)
print(txt)

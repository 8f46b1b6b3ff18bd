"""Per-source-line table of an ncu report (needs -lineinfo and --import-source on): share of executed warp instructions, active
lanes, stall samples and the dominant stall reason.   python tools/top_lines.py <report.ncu-rep> [rows]"""
import collections, csv, os, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
sections, cur = [], None
for row in csv.reader(src.split("\n")):
    if not row:
        continue
    if row[0] == "File Path":
        cur = {"file": row[1], "rows": [], "hdr": None}
        sections.append(cur)
    elif row[0] == "Line No":
        cur["hdr"] = row
    elif row[0] != "Function Name" and cur is not None and cur["hdr"]:
        cur["rows"].append(row)
agg, tot_i, tot_s = [], 0.0, 0.0
for s in sections:
    h = s["hdr"]
    iL, iS, iI, iT, iW = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(k, c) for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    per = collections.OrderedDict()
    for r in s["rows"]:
        try:
            ln, ie, te, ws = int(r[iL]), float(r[iI] or 0), float(r[iT] or 0), float(r[iW] or 0)
        except ValueError:
            continue
        e = per.setdefault(ln, [r[iS], 0.0, 0.0, 0.0, collections.Counter()])
        e[1] += ie; e[2] += te; e[3] += ws
        for k, c in stall_cols:
            try:
                e[4][c] += float(r[k] or 0)
            except ValueError:
                pass
    for ln, (text, ie, te, ws, st) in per.items():
        agg.append((ws, ie, te, os.path.basename(s["file"]), ln, text.strip()[:90], st))
        tot_i += ie; tot_s += ws
agg.sort(key=lambda a: -a[0])
print(f"{'samples':>8s} {'inst':>7s} {'lanes':>5s}  line  (dominant stall)")
for ws, ie, te, f, ln, text, st in agg[:top]:
    dom = ", ".join(f"{k.replace('stall_', '')} {v / max(ws, 1):.0%}" for k, v in st.most_common(2))
    print(f"{ws / max(tot_s, 1):8.2%} {ie / max(tot_i, 1):7.2%} {te / max(ie, 1):5.1f}  {f}:{ln}  {text}   ({dom})")

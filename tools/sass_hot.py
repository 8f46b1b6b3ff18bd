"""Per-SASS-instruction execution counts from an ncu report (--import-source on): python tools/sass_hot.py rep.ncu-rep [lo hi]
Prints address offset, executions per warp-step (given n_warps*n_steps), avg threads, SASS."""
import csv, subprocess, sys
rep = sys.argv[1]
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, iex, ith = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
base = int(rows[2][ia], 16)
for r in rows[2:]:
    if len(r) <= iex: continue
    print(f"{int(r[ia],16)-base:6x} {float(r[iex])/norm:9.2f} {float(r[ith]):5.1f}  {r[isrc].strip()}")

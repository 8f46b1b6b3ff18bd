"""Registers / spills per kernel from an `nvcc -Xptxas -v` log (cu++filt-demangled, one line per entry)."""
import re, subprocess, sys

def main(path):
    name = None
    rows = []
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name.replace("sodm::", ""))
            spill = None
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            spill = m.groups()
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            rows.append((name, int(m.group(1)), spill))
            name = None
    for n, r, s in rows:
        print(f"{r:4d} regs  stack {s[0]:>4s}  spill st/ld {s[1]:>4s}/{s[2]:>4s}  {n}")

if __name__ == "__main__":
    main(sys.argv[1])

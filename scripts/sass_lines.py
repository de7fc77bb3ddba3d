#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info: executed warp-instructions and stall
samples per CUDA source line.  usage: sass_lines.py report.ncu-rep cubin mangled_kernel_name [top] [substring of the demangled kernel name]
(the first launch of the report whose name contains the substring is used; default: the first launch)"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
want = sys.argv[5] if len(sys.argv) > 5 else ""
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
line_of = {}
cur = None
active = False
for ln in dis.splitlines():
    if ln.startswith(".text.") and ln.rstrip(":").endswith(kernel):
        active = True
        continue
    if active and ln.startswith(".text.") or (active and ln.startswith("//-----")):
        if not ln.rstrip(":").endswith(kernel) and "text" in ln:
            active = False if ".text." in ln and kernel not in ln else active
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).split()[0])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
base = None
skipping = False
ex = collections.Counter(); samples = collections.Counter(); total = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        if hdr is not None and base is not None:
            break                      # one launch only
        skipping = want not in r[1]
        continue
    if skipping:
        continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or not r or not r[0].startswith("0x"):
        continue
    d = dict(zip(hdr, r))
    addr = int(d["Address"], 16)
    if base is None:
        base = addr
    off = addr - base
    n = int(d["Instructions Executed"] or 0)
    s = int(d["# Samples"] or 0)
    loc = line_of.get(off, (None, "?"))[0]
    ex[loc] += n; samples[loc] += s; total += n
ts = sum(samples.values())
print(f"total warp-instructions {total}, samples {ts}")
for loc, n in ex.most_common(top):
    print(f"{100*n/total:5.1f}% inst  {100*samples[loc]/max(ts,1):5.1f}% samples  {loc}")
# per file
byfile = collections.Counter()
for loc, n in ex.items():
    byfile[loc[0] if loc else None] += n
print({k: f"{100*v/total:.1f}%" for k, v in byfile.most_common()})
sbyfile = collections.Counter()
for loc, n in samples.items():
    sbyfile[loc[0] if loc else None] += n
print("samples by file:", {k: f"{100*v/max(ts,1):.1f}%" for k, v in sbyfile.most_common()})
print("top lines by samples:")
for loc, n in samples.most_common(25):
    print(f"{100*n/max(ts,1):5.1f}% samples  {100*ex[loc]/total:5.1f}% inst  {loc}")

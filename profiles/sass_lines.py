#!/usr/bin/env python
"""Join an ncu SASS source page (--page source --csv) with nvdisasm -g line info: instructions executed and stall
samples per CUDA source line.   usage: sass_lines.py <src.csv> <nvdisasm.sass> <mangled kernel substring> [top]"""
import csv, re, sys, collections
src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> line from nvdisasm
addr_line = {}
cur = None; infunc = False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        infunc = kern in ln
        cur = None
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inl" if "inlined" in m.group(3) else ""); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and cur: addr_line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in rows[h + 1:]:
    try: a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except Exception: continue
    if base is None: base = a
    off = a - base
    n = int(float(r[ii] or 0)); s = int(float(r[isamp] or 0))
    key = addr_line.get(off, ((("?", 0, ""), "")))[0]
    agg[key][0] += n; agg[key][1] += s
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
for key, (n, s) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{key[0]:>14s}:{key[1]:<5d} inst {100*n/max(tot_i,1):5.1f}%  stall-samples {100*s/max(tot_s,1):5.1f}%")

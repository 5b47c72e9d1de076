#!/usr/bin/env python
"""SASS opcode histogram per kernel of libgmc.so (cuobjdump -sass), with the Blackwell-specific opcodes called out.
usage: sass_opcodes.py [libgmc.so] > profiles/r2/sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcmc_gpu_b200", "libgmc.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
SPECIAL = ("UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTMACMDFLUSH", "SYNCS", "UTCBAR", "UTC", "LDTM", "STTM", "LDGSTS", "RED", "ATOMG", "ELECT",
           "FENCE", "VIADDMNMX", "R2UR", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "WARPSYNC", "CALL")
cur, hist = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.basename(lib)}: static SASS opcode counts per kernel (sm_100a)")
print("# TMA tensor copies = UTMALDG / UTMASTG; 1-D bulk copies = UBLKCP / UBLKPF; mbarrier = SYNCS.*; no tcgen05 (UTC*MMA / LDTM): the path is FP64")
for (name, c), dm in zip(hist.items(), demangle):
    tot = sum(c.values())
    print(f"\n{dm[:150]}\n  total {tot}")
    print("  special: " + ", ".join(f"{k} {c[k]}" for k in SPECIAL if c.get(k)))
    print("  top:     " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))

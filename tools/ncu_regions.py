#!/usr/bin/env python
"""Sum executed warp-instructions / samples of one kernel per source-line range.
usage: ncu_regions.py report.ncu-rep kernel_regex file.cu name:lo-hi [name:lo-hi ...]"""
import csv, io, subprocess, sys
rep, kern, fil = sys.argv[1:4]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", __import__("os").environ.get("SKIP","0"), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; acc = {}; tot = [0, 0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:])); g = lambda k: int(float(d.get(k) or 0))
        s, i, t = g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed")
        tot[0] += s; tot[1] += i; tot[2] += t
        key = "other:" + fname
        if fname == fil:
            ln = int(r[0])
            for n, lo, hi in regions:
                if lo <= ln <= hi: key = n; break
            else: key = "unassigned"
        a = acc.setdefault(key, [0, 0, 0]); a[0] += s; a[1] += i; a[2] += t
print(f"total: samples {tot[0]} warp-instr {tot[1]} lanes/instr {tot[2]/max(tot[1],1):.1f}")
for k, (s, i, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {100*i/tot[1]:5.1f}% instr  {100*s/tot[0]:5.1f}% samples  lanes/instr {t/max(i,1):4.1f}  warp-instr {i}")

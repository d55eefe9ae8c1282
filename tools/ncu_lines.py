#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [top] [ins]   (ins: sort by executed instructions instead of samples)"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
byins = len(sys.argv) > 5 and sys.argv[5] == "ins"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        g = lambda k: int(float(d.get(k) or 0))
        data.append((g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"), fname, int(r[0]), r[1],
                     {k[6:]: g(k) for k in d if k.startswith("stall_") and "Not Issued" not in k and g(k)}))
ts = sum(x[0] for x in data) or 1; ti = sum(x[1] for x in data) or 1
print(f"total samples {ts}  warp-instructions {ti}  thread-instr/instr {sum(x[2] for x in data)/ti:.1f}")
for s, i, t, f, ln, src, st in sorted(data, key=(lambda x: x[1]) if byins else (lambda x: x[0]), reverse=True)[:top]:
    tops = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/ts:5.1f}%smp {100*i/ti:5.1f}%ins thr/ins {t/max(i,1):4.1f} {f}:{ln:<4} {src.strip()[:90]}  [{tops}]")

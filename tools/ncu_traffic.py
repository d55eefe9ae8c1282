#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full` report: duration, warp-instructions, IPC, dram bytes, occupancy, registers.
usage: ncu_traffic.py report.ncu-rep [kernel_regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = {"Kernel Name": "kernel", "gpu__time_duration.sum": "ns", "smsp__inst_executed.sum": "winst", "sm__inst_executed.avg.per_cycle_active": "ipc_active",
        "dram__bytes_read.sum": "dram_r", "dram__bytes_write.sum": "dram_w", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
        "launch__registers_per_thread": "regs", "launch__occupancy_limit_registers": "occ_lim_regs", "sm__maximum_warps_per_active_cycle_pct": "theo_occ",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "thr_per_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
        "lts__t_bytes.sum": "l2_bytes"}
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
for r in rows[2:]:
    d = {}
    for k, name in want.items():
        if k in idx:
            v = r[idx[k]]
            u = units[idx[k]]
            if name in ("dram_r", "dram_w", "l2_bytes"):
                f = float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                v = f"{f/1e6:.2f}MB"
            if name == "kernel":
                v = v.split("(")[0].replace("<unnamed>::", "")
            d[name] = v
    print("  ".join(f"{k}={v}" for k, v in d.items()))

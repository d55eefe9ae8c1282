#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of `python bench.py --steps 2 --warmup 3` (the last two-pass frame captured is
the steady-state one): per stage the dram bytes, warp-instructions, IPC, duration.  usage: make_traffic.py report.ncu-rep cfgN capture-note"""
import csv, io, json, os, subprocess, sys
rep, cfg, note = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def val(r, k, scale_bytes=False):
    v = float(r[idx[k]].replace(",", ""))
    if scale_bytes:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[k]], 1)
    return v
launches = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
    launches.append((name, r))
# steady frame = the last occurrence of the sequence cull, raster, big, hiz, cull, raster, big, hiz
names = [n for n, _ in launches]
seq = ["cull_kernel", "raster_kernel", "raster_big_kernel", "hiz_tiled_kernel"] * 2
start = max(i for i in range(len(names) - 7) if names[i:i + 8] == seq)
stage_of = ["cull_a", "raster_a", "drain_a", "hiz_a", "cull_b", "raster_b", "drain_b", "hiz_b"]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
doc = json.load(open(path)) if os.path.exists(path) else {}
ent = {}
for st, (name, r) in zip(stage_of, launches[start:start + 8]):
    ent[st] = {"kernel": name, "dram_read_bytes": int(val(r, "dram__bytes_read.sum", True)), "dram_write_bytes": int(val(r, "dram__bytes_write.sum", True)),
               "ipc_active": round(val(r, "sm__inst_executed.avg.per_cycle_active"), 3), "warp_instructions": int(val(r, "smsp__inst_executed.sum")),
               "ncu_duration_us": round(val(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "usecond": 1, "msecond": 1e3, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1e-3), 2),
               "warps_active_pct": round(val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
               "registers": int(val(r, "launch__registers_per_thread")), "capture": note}
doc[cfg] = ent
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(ent, indent=1)[:1500])

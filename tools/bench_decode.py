#!/usr/bin/env python
"""Decode throughput of the device-side EXT_meshopt_compression path beside the reference's CPU decoder.

    python tools/bench_decode.py [--views 2000] [--reps 20]

Workload: a synthetic 'city' asset in the shape of BASELINE config 4 — `--views` compressed buffer views replicated from the
committed fixture (3000 x 24 B attribute streams, 2000 x 12 B attribute streams, 9126-index triangle streams), i.e. many
independent streams of a few thousand elements each, which is what the reference decodes with one enkiTS task per view
(assets.cpp:111-171).  Prints one JSON line: decoded GB/s with the compressed bytes resident in HBM (CUDA events around
vkv_meshopt_run), end to end from host memory (upload + decode + download), and the CPU decoder on this box's host cores
(the reference's meshoptimizer from oracle/_ref when present, else the oracle port; single thread — say so)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import meshopt_lib as M
from tests.test_gpu_meshopt import pack
from vk_gltf_viewer_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=2000)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
allc = {c["name"]: c for c in M.golden_cases()}
base = [allc["vertex_smooth_3000x24"], allc["vertex_mixed_2000x12"], allc["index_grid40_v1_32"], allc["vertex_smooth_1100x32"]]
cases = [base[i % len(base)] for i in range(a.views)]
src, views, dst_bytes = pack(cases)
r = api.Renderer(64, 64)
s_dev, d_dev = r.upload(src), r.alloc(dst_bytes)
plan = r.meshopt_plan(views)
for _ in range(3):
    r.meshopt_run(plan, s_dev, src.size, d_dev, dst_bytes)
assert (r.meshopt_results(plan, len(views)) == 0).all()
ms = []
for _ in range(a.reps):
    r.flush_l2(256 << 20)
    r.event_record(0); r.meshopt_run(plan, s_dev, src.size, d_dev, dst_bytes); r.event_record(1)
    ms.append(r.event_elapsed(0, 1))
out = r.download(d_dev, dst_bytes)
t0 = time.perf_counter()
for _ in range(3):
    s2 = r.upload(src); r.meshopt_run(plan, s2, src.size, d_dev, dst_bytes); out = r.download(d_dev, dst_bytes); r.free(s2)
e2e = (time.perf_counter() - t0) / 3
# CPU: the reference's decoder, one thread, same views
kind = "reference" if M.ref_lib() is not None else "port"
dec = (lambda c: M.ref_decode(c["kind"], c["count"], c["stride"], c["enc"])) if kind == "reference" else (lambda c: M.oracle_decode(c["kind"], c["count"], c["stride"], c["enc"]))
n_cpu = min(len(cases), 400)
t0 = time.perf_counter()
for c in cases[:n_cpu]:
    dec(c)
cpu_s = (time.perf_counter() - t0) * len(cases) / n_cpu
decoded = sum(c["dec"].size for c in cases)
best = float(np.median(ms))
print(json.dumps({"metric": "decoded GB/s (EXT_meshopt_compression, attribute + index views)", "views": len(cases), "compressed_bytes": int(sum(c["enc"].size for c in cases)),
                  "decoded_bytes": int(decoded), "gpu_ms_resident": round(best, 4), "gpu_GBps_resident": round(decoded / best / 1e6, 2),
                  "e2e_ms": round(e2e * 1e3, 3), "e2e_GBps": round(decoded / e2e / 1e9, 3),
                  "cpu": {"kind": kind, "cores": 1, "ms": round(cpu_s * 1e3, 2), "GBps": round(decoded / cpu_s / 1e9, 3), "sample": f"{n_cpu} of {len(cases)} views, scaled"}}))
r.close()

#!/usr/bin/env python
"""Throughput of the device-side meshlet builder beside the CPU builders.

    python tools/bench_meshlets.py [--prims 200] [--grid 160]

Workload: `--prims` primitives, each a (grid x grid) quad patch (2*(grid-1)^2 triangles) — the shape of BASELINE config 4's
buildings / config 3's patch.  Prints one JSON line: triangles/s of vkv_build_meshlets (wall clock around the call: two
device phases and one 12-byte-per-primitive read-back in between; inputs resident in HBM), of the reference's
meshopt_buildMeshletsScan and meshopt_buildMeshlets (the function the reference calls) on one host core when oracle/_ref is
present, else of the oracle port."""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import meshlet_lib as ML
from tests import meshopt_lib as M
from tests.test_gpu_meshlets import vertices24
from vk_gltf_viewer_b200 import abi, api

ap = argparse.ArgumentParser()
ap.add_argument("--prims", type=int, default=200)
ap.add_argument("--grid", type=int, default=160)
a = ap.parse_args()
n = a.grid
u, v = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
pos = np.stack([u, v, 0.05 * np.sin(40 * u) * np.cos(33 * v)], -1).reshape(-1, 3).astype(np.float32)
q = np.arange(n * n).reshape(n, n)
idx = np.stack([q[:-1, :-1].ravel(), q[:-1, 1:].ravel(), q[1:, :-1].ravel(), q[1:, :-1].ravel(), q[:-1, 1:].ravel(), q[1:, 1:].ravel()], 1).astype(np.uint32).reshape(-1)
r = api.Renderer(64, 64)
vi, vv = r.upload(idx), r.upload(vertices24(pos))
inp = np.zeros(a.prims, abi.MESHLET_BUILD_INPUT_DTYPE)
inp[:] = (vi, vv, idx.size, pos.shape[0])
tris = a.prims * idx.size // 3
best = 1e9
for rep in range(4):
    r.sync()
    t0 = time.perf_counter()
    out = r.build_meshlets(inp)
    dt = time.perf_counter() - t0
    for x in (out[0]["meshlets"], out[0]["vertex_indices"], out[0]["triangles"]):
        r.free(int(x))
    if rep:
        best = min(best, dt)
res = {"metric": "triangles/s (meshlet partition + bounds)", "primitives": a.prims, "triangles": tris, "meshlets": int(out["meshlet_count"].sum()),
       "gpu_ms": round(best * 1e3, 3), "gpu_Mtris_per_s": round(tris / best / 1e6, 1)}
t0 = time.perf_counter(); ML.oracle_scan(idx, pos.shape[0]); t_port = time.perf_counter() - t0
res["cpu_port_scan_Mtris_per_s_1core"] = round(idx.size / 3 / t_port / 1e6, 2)
if M.ref_lib() is not None:
    t0 = time.perf_counter(); ML.ref_scan(idx, pos.shape[0]); t_ref = time.perf_counter() - t0
    res["cpu_reference_scan_Mtris_per_s_1core"] = round(idx.size / 3 / t_ref / 1e6, 2)
    L = M.ref_lib()
    L.meshopt_buildMeshlets.restype = C.c_size_t
    nb = ML.bound(idx.size)
    m = np.zeros((nb, 4), np.uint32); mv = np.zeros(nb * 64, np.uint32); mt = np.zeros(nb * 124 * 3, np.uint8)
    t0 = time.perf_counter()
    L.meshopt_buildMeshlets(m.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p), mt.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), C.c_size_t(idx.size),
                            pos.ctypes.data_as(C.c_void_p), C.c_size_t(pos.shape[0]), C.c_size_t(12), C.c_size_t(64), C.c_size_t(124), C.c_float(0.0))
    t_full = time.perf_counter() - t0
    res["cpu_reference_buildMeshlets_Mtris_per_s_1core"] = round(idx.size / 3 / t_full / 1e6, 2)
print(json.dumps(res))
r.close()

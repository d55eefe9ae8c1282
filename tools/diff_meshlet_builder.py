#!/usr/bin/env python
"""Randomised differential test of the host meshlet builder (host/clusterizer.cpp) against the reference's meshoptimizer
(oracle/_ref/libmeshopt_ref.so, built from the reference tree by `make ref`): meshopt_buildMeshletsBound / buildMeshlets / optimizeMeshlet /
computeMeshletBounds must agree byte for byte on random vertex soups, shuffled grids, degenerate and duplicated triangles, coincident positions,
coordinates from 1e-6 to 1e9, with the reference's limits (64 / 124 / 0) and random ones.

    python tools/diff_meshlet_builder.py [seed] [iterations]      (1 650 cases at seeds 1 and 2: 0 mismatches)
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests import test_host as TH, scenes as S
from vk_gltf_viewer_b200._native import host_lib
ref = C.CDLL(os.path.join(ROOT, 'oracle', '_ref', 'libmeshopt_ref.so')); H = host_lib()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
bad = 0; n = 0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 150):
    kind = it % 5
    if kind == 0:      # random soup with shared vertices
        nv = int(rng.integers(3, 400)); nt = int(rng.integers(1, 900))
        pos = rng.uniform(-1,1,(nv,3)).astype(np.float32); idx = rng.integers(0,nv,nt*3).astype(np.uint32)
    elif kind == 1:    # grid, shuffled triangles, random heights
        a,b = int(rng.integers(2,60)), int(rng.integers(2,60))
        f = float(rng.uniform(1,20))
        pos,idx = S.grid_mesh(a,b,lambda u,v: (u*3, np.sin(u*f)*np.cos(v*f*0.7), v*2)); idx = rng.permutation(idx.reshape(-1,3)).reshape(-1).astype(np.uint32)
    elif kind == 2:    # grid in order, with duplicated/degenerate triangles sprinkled in
        a,b = int(rng.integers(2,40)), int(rng.integers(2,40))
        pos,idx = S.grid_mesh(a,b,lambda u,v: (u, v, 0.2*u*v)); idx = idx.astype(np.uint32).copy()
        for _ in range(int(rng.integers(0,10))):
            k = int(rng.integers(0, idx.size//3))*3; idx[k+1] = idx[k]
    elif kind == 3:    # coincident positions (zero-size boxes) and collinear runs
        nv = int(rng.integers(3, 100)); pos = np.repeat(rng.uniform(-1,1,(max(1,nv//4),3)).astype(np.float32), 4, 0)[:nv]
        if pos.shape[0] < 3: continue
        idx = rng.integers(0,pos.shape[0],int(rng.integers(1,200))*3).astype(np.uint32)
    else:              # large coordinates / tiny triangles
        a,b = int(rng.integers(2,30)), int(rng.integers(2,30)); s = float(10**rng.uniform(-6,6))
        pos,idx = S.grid_mesh(a,b,lambda u,v: (u*s+1000*s, np.sin(u*5)*s, v*s)); idx = idx.astype(np.uint32)
    pos = np.ascontiguousarray(pos, np.float32); idx = np.ascontiguousarray(idx, np.uint32)
    for (mv_,mt_,cw) in ((64,124,0.0),(64,124,0.7),(int(rng.integers(3,256)), int(rng.integers(1,129))*4, float(rng.uniform(0,1)))):
        try:
            r = TH._run_builder(ref.meshopt_buildMeshletsBound, ref.meshopt_buildMeshlets, ref.meshopt_optimizeMeshlet, ref.meshopt_computeMeshletBounds, pos, idx, mv_, mt_, cw, bounds_by_value=True)
            g = TH._run_builder(H.vkvh_meshlets_bound, H.vkvh_meshlets_build, H.vkvh_meshlet_optimize, H.vkvh_meshlet_bounds, pos, idx, mv_, mt_, cw)
        except Exception as e:
            print("exception", it, kind, (mv_,mt_,cw), repr(e)[:100]); bad += 1; continue
        n += 1
        same = g[0]==r[0] and np.array_equal(g[1],r[1]) and np.array_equal(g[2],r[2]) and np.array_equal(g[3],r[3])
        nb = sum(a!=b for a,b in zip(g[4],r[4]))
        if not same or nb:
            bad += 1; print("MISMATCH", it, kind, (mv_,mt_,round(cw,3)), "verts", pos.shape[0], "tris", idx.size//3, "partition same", same, "bounds differing", nb)
print("cases", n, "mismatches", bad)
sys.exit(1 if bad else 0)

"""ctypes faces for the meshlet-build tests: the oracle restatement (oracle/meshlet_build.cpp), the reference's
meshopt_buildMeshletsScan when oracle/_ref is present, test meshes and the golden fixture.  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

from tests.conftest import ROOT
from tests.oracle_lib import lib as oracle_lib
from tests.meshopt_lib import ref_lib

GOLDEN = os.path.join(ROOT, "tests", "golden", "meshlet_scan.npz")
MAXV, MAXT = 64, 124  # mesh_common.h.glsl:36-37 maxVertices, alignDown(maxPrimitives, 4) (assets.cpp:324)


def bound(index_count, maxv=MAXV, maxt=MAXT):
    """meshopt_buildMeshletsBound (clusterizer.cpp:513-533)"""
    return max((index_count + maxv - 3) // (maxv - 2), (index_count // 3 + maxt - 1) // maxt)


def _run(fn, indices, nverts, maxv, maxt):
    idx = np.ascontiguousarray(indices, np.uint32)
    nb = max(1, bound(idx.size, maxv, maxt))
    m = np.zeros((nb, 4), np.uint32)
    mv = np.zeros(nb * maxv, np.uint32)
    mt = np.zeros(nb * maxt * 3, np.uint8)
    fn.restype = C.c_size_t
    n = fn(m.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p), mt.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p),
           C.c_size_t(idx.size), C.c_size_t(nverts), C.c_size_t(maxv), C.c_size_t(maxt))
    m = m[:n]
    if n == 0:
        return m, mv[:0], mt[:0]
    nv = int(m[-1, 0]) + int(m[-1, 2])
    nt = int(m[-1, 1]) + ((int(m[-1, 3]) * 3 + 3) & ~3)
    return m.copy(), mv[:nv].copy(), mt[:nt].copy()


def oracle_scan(indices, nverts, maxv=MAXV, maxt=MAXT):
    return _run(oracle_lib().orc_meshlets_scan, indices, nverts, maxv, maxt)


def ref_scan(indices, nverts, maxv=MAXV, maxt=MAXT):
    return _run(ref_lib().meshopt_buildMeshletsScan, indices, nverts, maxv, maxt)


def oracle_bounds(meshlets, mverts, positions):
    pos = np.ascontiguousarray(positions, np.float32)
    out = np.zeros((meshlets.shape[0], 6), np.float32)
    m = np.ascontiguousarray(meshlets, np.uint32)
    oracle_lib().orc_meshlet_bounds(m.ctypes.data_as(C.c_void_p), C.c_size_t(m.shape[0]), mverts.ctypes.data_as(C.c_void_p),
                                    pos.ctypes.data_as(C.c_void_p), C.c_size_t(pos.strides[0]), out.ctypes.data_as(C.c_void_p))
    return out


def meshes(seed=0):
    """name -> (positions float32 (n,3), indices uint32): the shapes that stress the partition rules"""
    rng = np.random.default_rng(0x5EED0F4 + seed)

    def grid(n, shuffle=False):
        u, v = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
        pos = np.stack([u, v, 0.1 * np.sin(7 * u) * np.cos(5 * v)], -1).reshape(-1, 3).astype(np.float32)
        q = np.arange(n * n).reshape(n, n)
        a, b, c, d = q[:-1, :-1].ravel(), q[:-1, 1:].ravel(), q[1:, :-1].ravel(), q[1:, 1:].ravel()
        tris = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)]).astype(np.uint32)
        if shuffle:
            tris = tris[rng.permutation(tris.shape[0])]
        return pos, tris.reshape(-1)

    out = {}
    out["grid24"] = grid(24)
    out["grid40_shuffled"] = grid(40, True)           # every triangle brings new vertices: the 64-vertex limit closes meshlets
    p, i = grid(30)
    order = np.lexsort((np.arange(i.size // 3) % 2, np.arange(i.size // 3) // 2 % 29))
    out["grid30_strips"] = (p, i.reshape(-1, 3)[order].reshape(-1))
    n = 900                                           # unindexed soup: 3 new vertices per triangle, 21 triangles per meshlet
    out["soup"] = (rng.uniform(-1, 1, (3 * n, 3)).astype(np.float32), np.arange(3 * n, dtype=np.uint32))
    fan = np.stack([np.zeros(700), np.arange(1, 701), np.arange(2, 702)], 1).astype(np.uint32).reshape(-1)
    out["fan"] = (rng.uniform(-1, 1, (702, 3)).astype(np.float32), fan)   # one shared vertex: the 124-triangle limit closes meshlets
    deg = rng.integers(0, 50, (400, 3)).astype(np.uint32)                  # tiny vertex pool: repeated corners (a == b), re-used ids
    deg[::7, 1] = deg[::7, 0]
    deg[::11, 2] = deg[::11, 0]
    out["degenerate"] = (rng.uniform(-1, 1, (50, 3)).astype(np.float32), deg.reshape(-1))
    out["single"] = (np.eye(3, dtype=np.float32), np.arange(3, dtype=np.uint32))
    big = rng.integers(0, 3000, (5000, 3)).astype(np.uint32)
    out["random_ids"] = (rng.uniform(-4, 4, (3000, 3)).astype(np.float32), big.reshape(-1))
    return out

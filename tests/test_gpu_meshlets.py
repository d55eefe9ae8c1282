"""GPU parity of the device-side meshlet builder through the C ABI (vkv_build_meshlets): PINNED to
meshopt_buildMeshletsScan of the reference's meshoptimizer (vectors frozen in tests/golden/meshlet_scan.npz, generator
tests/golden/make_meshlet_golden.py) — meshlet records, vertex-index lists and triangle bytes byte for byte — and bounds equal to
assets.cpp:349-372 as restated in oracle/meshlet_build.cpp, bit for bit."""
import numpy as np
import pytest

from tests import meshlet_lib as ML
from vk_gltf_viewer_b200 import abi, api

pytestmark = pytest.mark.gpu
G = np.load(ML.GOLDEN)
NAMES = [str(n) for n in G["names"]]


def vertices24(pos):
    v = np.zeros(pos.shape[0], abi.VERTEX_DTYPE)
    v["position"] = pos
    return v


def build(r, meshes, maxv=ML.MAXV, maxt=ML.MAXT):
    """meshes: list of (pos, idx) -> list of (meshlets (n,) MESHLET_DTYPE, mverts u32, mtris u8)"""
    inp = np.zeros(len(meshes), abi.MESHLET_BUILD_INPUT_DTYPE)
    keep = []
    for i, (pos, idx) in enumerate(meshes):
        idx = np.ascontiguousarray(idx, np.uint32)
        vi = r.upload(idx) if idx.size else 0
        vv = r.upload(vertices24(pos)) if pos.shape[0] else 0
        keep += [vi, vv]
        inp[i] = (vi, vv, idx.size, pos.shape[0])
    out = r.build_meshlets(inp, 24, maxv, maxt)
    res = []
    for o in out:
        m = r.download(int(o["meshlets"]), int(o["meshlet_count"]) * 36).view(abi.MESHLET_DTYPE) if o["meshlet_count"] else np.zeros(0, abi.MESHLET_DTYPE)
        mv = r.download(int(o["vertex_indices"]), int(o["vertex_index_count"]) * 4).view(np.uint32) if o["vertex_index_count"] else np.zeros(0, np.uint32)
        mt = r.download(int(o["triangles"]), int(o["triangle_bytes"])) if o["triangle_bytes"] else np.zeros(0, np.uint8)
        res.append((m, mv, mt))
    if len(out) and out[0]["meshlets"]:
        for a in (out[0]["meshlets"], out[0]["vertex_indices"], out[0]["triangles"]):
            r.free(int(a))
    for a in keep:
        if a:
            r.free(a)
    return res


def check(got, pos, want_m, want_mv, want_mt, label):
    m, mv, mt = got
    have = np.stack([m["vertexOffset"], m["triangleOffset"], m["vertexCount"].astype(np.uint32), m["triangleCount"].astype(np.uint32)], 1) if m.size else np.zeros((0, 4), np.uint32)
    assert np.array_equal(have, want_m), f"{label}: meshlet records differ ({have.shape[0]} vs {want_m.shape[0]})"
    assert np.array_equal(mv, want_mv), f"{label}: vertex-index lists differ"
    assert np.array_equal(mt, want_mt), f"{label}: triangle bytes differ"
    if m.size:
        b = ML.oracle_bounds(want_m, want_mv, pos)
        assert np.array_equal(m["aabbExtents"].view(np.uint32), b[:, :3].copy().view(np.uint32)), f"{label}: extents differ"
        assert np.array_equal(m["aabbCenter"].view(np.uint32), b[:, 3:].copy().view(np.uint32)), f"{label}: centres differ"
        raw = m.view(np.uint8).reshape(-1, 36)
        assert not raw[:, 10:12].any()  # padding bytes written as zero


@pytest.mark.parametrize("name", NAMES)
def test_golden_mesh(name):
    r = api.Renderer(64, 64)
    pos, idx = G[name + "_pos"], G[name + "_idx"]
    check(build(r, [(pos, idx)])[0], pos, G[name + "_m"], G[name + "_mv"], G[name + "_mt"], name)
    r.close()


def test_all_golden_meshes_in_one_call_with_empty_primitives_between():
    r = api.Renderer(64, 64)
    empty = (np.zeros((0, 3), np.float32), np.zeros(0, np.uint32))
    meshes, names = [empty], [None]
    for n in NAMES:
        meshes += [(G[n + "_pos"], G[n + "_idx"]), empty]
        names += [n, None]
    for got, n, (pos, _) in zip(build(r, meshes), names, meshes):
        if n is None:
            assert got[0].size == 0 and got[1].size == 0 and got[2].size == 0
        else:
            check(got, pos, G[n + "_m"], G[n + "_mv"], G[n + "_mt"], n)
    r.close()


@pytest.mark.parametrize("limits", [(64, 124), (32, 64), (3, 4), (64, 8), (17, 252)])
def test_other_limits_match_the_oracle(limits):
    maxv, maxt = limits
    r = api.Renderer(64, 64)
    ms = ML.meshes(5)
    got = build(r, list(ms.values()), maxv, maxt)
    for g, (name, (pos, idx)) in zip(got, ms.items()):
        m, mv, mt = ML.oracle_scan(idx, pos.shape[0], maxv, maxt)
        check(g, pos, m, mv, mt, f"{name} {maxv}/{maxt}")
    r.close()


def test_large_mesh_many_segments():
    """700 k triangles = 340 chain segments, two primitives: the same bytes as the sequential builder; validity at full size"""
    n = 420
    u, v = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
    pos = np.stack([u, v, 0.05 * np.sin(40 * u) * np.cos(33 * v)], -1).reshape(-1, 3).astype(np.float32)
    q = np.arange(n * n).reshape(n, n)
    a, b, c, d = q[:-1, :-1].ravel(), q[:-1, 1:].ravel(), q[1:, :-1].ravel(), q[1:, 1:].ravel()
    idx = np.stack([a, b, c, c, b, d], 1).astype(np.uint32).reshape(-1)
    rng = np.random.default_rng(3)
    idx2 = idx.reshape(-1, 3)[rng.permutation(idx.size // 3)[:200000]].reshape(-1)
    r = api.Renderer(64, 64)
    got = build(r, [(pos, idx), (pos, idx2)])
    r.close()
    for g, i in zip(got, (idx, idx2)):
        m, mv, mt = ML.oracle_scan(i, pos.shape[0])
        check(g, pos, m, mv, mt, "large")
        tris = np.concatenate([mv[vo:vo + vc][mt[to:to + tc * 3].reshape(-1, 3)] for vo, to, vc, tc in m])
        assert np.array_equal(tris.reshape(-1), i)


def test_argument_validation():
    r = api.Renderer(64, 64)
    inp = np.zeros(1, abi.MESHLET_BUILD_INPUT_DTYPE)
    inp[0] = (0, 0, 4, 3)
    with pytest.raises(api.VkvError):
        r.build_meshlets(inp)                     # index count not a multiple of 3
    inp[0] = (0, 0, 3, 3)
    with pytest.raises(api.VkvError):
        r.build_meshlets(inp)                     # NULL buffers
    inp[0] = (0, 0, 0, 0)
    with pytest.raises(api.VkvError):
        r.build_meshlets(inp, max_vertices=65)
    out = r.build_meshlets(inp)
    assert out[0]["meshlet_count"] == 0
    r.close()


def test_out_of_range_index_is_rejected_before_any_vertex_is_read():
    """vkv_build_meshlets range-checks the indices on the device when vertex_count is given (untrusted glTF) and fails
    with VKV_ERR_INVALID; vertex_count == 0 skips the check"""
    r = api.Renderer(64, 64)
    pos = np.zeros((10, 6), np.float32)
    idx = np.array([0, 1, 2, 2, 3, 11, 4, 5, 6], np.uint32)   # 11 >= 10 vertices
    inp = np.zeros(1, abi.MESHLET_BUILD_INPUT_DTYPE)
    inp["indices"] = r.upload(idx); inp["vertices"] = r.upload(pos); inp["index_count"] = idx.size; inp["vertex_count"] = 10
    with pytest.raises(api.VkvError) as e:
        r.build_meshlets(inp)
    assert e.value.code == -2 and "index" in str(e.value)
    idx[5] = 9
    inp["indices"] = r.upload(idx)
    out = r.build_meshlets(inp)
    assert out["meshlet_count"][0] == 1
    r.close()

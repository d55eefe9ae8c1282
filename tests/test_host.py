"""Host-side input generators (libvkv_host.so) against the reference's own host libraries where they compile
(glm, fastgltf::math, meshoptimizer from oracle/_ref) and against structural invariants."""
import ctypes as C

import numpy as np
import pytest

from tests import scenes as S
from vk_gltf_viewer_b200 import abi
from vk_gltf_viewer_b200.scene import Camera, Scene, set_meshlet_builder


def check_meshlets(prim, n_tris_expected):
    """every source triangle exactly once; <=64 vertices, <=124 triangles; local indices in range; 4-byte aligned
    triangle offsets; AABB = exact min/max of the meshlet's vertices (assets.cpp:349-372)"""
    ml, vi, tb, vx = prim["meshlets"], prim["vertex_indices"], prim["triangles"], prim["vertices"]["position"]
    seen = []
    for m in ml:
        vc, tc = int(m["vertexCount"]), int(m["triangleCount"])
        assert 0 < vc <= 64 and 0 < tc <= 124
        assert m["triangleOffset"] % 4 == 0
        loc = tb[m["triangleOffset"]: m["triangleOffset"] + 3 * tc].reshape(tc, 3)
        assert loc.max() < vc
        glob = vi[m["vertexOffset"]: m["vertexOffset"] + vc]
        assert len(set(glob.tolist())) == vc
        seen.append(glob[loc])
        p = vx[glob]
        mn, mx = p.min(0), p.max(0)
        c = (mn + mx) * np.float32(0.5)
        assert np.array_equal(m["aabbCenter"], c) and np.array_equal(m["aabbExtents"], mx - c)
    seen = np.concatenate(seen)
    assert seen.shape[0] == n_tris_expected
    return seen


def canon(tris):
    """rotation-invariant canonical form of oriented triangles"""
    t = np.asarray(tris, np.int64)
    k = np.argmin(t, axis=1)
    r = np.stack([np.roll(row, -s) for row, s in zip(t, k)]) if len(t) < 5000 else np.take_along_axis(
        np.concatenate([t, t], 1), (k[:, None] + np.arange(3)[None, :]), 1)
    return r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]


def test_icosphere_counts_and_orientation():
    f = 9
    s = Scene.icosphere(f)
    c = s.counts()
    assert c.triangles_unique == 20 * f * f and c.vertices_unique == 10 * f * f + 2 and c.draws == c.meshlets_unique
    prim = s.primitive(0)
    tris = check_meshlets(prim, 20 * f * f)
    pos = prim["vertices"]["position"]
    assert np.allclose(np.linalg.norm(pos, axis=1), 1.0, atol=1e-6)
    a, b, cc = pos[tris[:, 0]], pos[tris[:, 1]], pos[tris[:, 2]]
    n = np.cross(b - a, cc - a)
    assert (np.einsum("ij,ij->i", n, a + b + cc) > 0).all(), "triangles must be CCW seen from outside (glTF convention)"
    assert np.unique(canon(tris), axis=0).shape[0] == 20 * f * f  # no duplicates


def test_cfg1_numbers():
    c = Scene.icosphere(57).counts()
    assert c.triangles_unique == 64980 and c.vertices_unique == 32492 and 500 < c.meshlets_unique < 1100


def test_builtin_builder_partitions_every_triangle_once():
    rng = np.random.default_rng(5)
    pos, idx = S.grid_mesh(37, 23, lambda u, v: (u * 3, np.sin(u * 7) * np.cos(v * 3), v * 2))
    perm = rng.permutation(idx.reshape(-1, 3))  # shuffled index order must not matter for validity
    s = Scene.new()
    s.add_primitive(pos, perm.reshape(-1))
    got = check_meshlets(s.primitive(0), perm.shape[0])
    assert np.array_equal(canon(got), canon(perm))


class _Meshlet(C.Structure):
    _fields_ = [("vertex_offset", C.c_uint), ("triangle_offset", C.c_uint), ("vertex_count", C.c_uint), ("triangle_count", C.c_uint)]


class _Bounds(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("cone_apex", C.c_float * 3), ("cone_axis", C.c_float * 3),
                ("cone_cutoff", C.c_float), ("cone_axis_s8", C.c_byte * 3), ("cone_cutoff_s8", C.c_byte)]


def _run_builder(bound, build, optimize, bounds, pos, idx, max_v=64, max_t=124, cone_weight=0.0, bounds_by_value=False):
    """assets.cpp:322-346 call sequence through one set of meshoptimizer-shaped entry points -> raw output bytes"""
    vtx = np.zeros((pos.shape[0], 6), np.float32)   # glsl::Vertex stride: 24 bytes, position first
    vtx[:, :3] = pos
    for f in (bound, build):
        f.restype = C.c_size_t
    bound.argtypes = [C.c_size_t] * 3
    build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float]
    optimize.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    optimize.restype = None
    nmax = bound(idx.size, max_v, max_t)
    ms = (_Meshlet * nmax)()
    mv = np.zeros(nmax * max_v, np.uint32)
    mt = np.zeros(nmax * max_t * 3, np.uint8)
    n = build(ms, mv.ctypes.data, mt.ctypes.data, idx.ctypes.data, idx.size, vtx.ctypes.data, vtx.shape[0], 24, max_v, max_t, cone_weight)
    rec = np.array([[m.vertex_offset, m.triangle_offset, m.vertex_count, m.triangle_count] for m in ms[:n]], np.int64)
    last = rec[-1]
    mv = mv[: last[0] + last[2]]
    mt = mt[: last[1] + ((last[3] * 3 + 3) & ~3)]
    bnd = []
    for vo, to, vc, tc in rec:
        optimize(mv.ctypes.data + 4 * int(vo), mt.ctypes.data + int(to), int(tc), int(vc))
        a = (mv.ctypes.data + 4 * int(vo), mt.ctypes.data + int(to), int(tc), vtx.ctypes.data, vtx.shape[0], 24)
        if bounds_by_value:   # meshoptimizer returns the 48-byte struct by value
            bounds.restype = _Bounds
            bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]
            b = bounds(*a)
        else:
            bounds.restype = None
            bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(_Bounds)]
            b = _Bounds()
            bounds(*a, C.byref(b))
        bnd.append(bytes(b))
    return nmax, rec, mv, mt, bnd


def _builder_meshes():
    rng = np.random.default_rng(11)
    out = {}
    out["patch224"] = S.grid_mesh(224, 224, lambda u, v: (u, 0.05 * np.sin(u * 9) * np.cos(v * 7), v))           # the cfg-3 / cfg-5 patch shape
    pos, idx = S.grid_mesh(37, 23, lambda u, v: (u * 3, np.sin(u * 7) * np.cos(v * 3), v * 2))
    out["shuffled"] = (pos, rng.permutation(idx.reshape(-1, 3)).reshape(-1).astype(np.uint32))              # index order must be followed exactly
    sphere = Scene.icosphere(12)            # a closed mesh: take its triangle soup back out of the meshlets
    ico = sphere.primitive(0)
    pos = ico["vertices"]["position"].copy()
    soup = np.concatenate([ico["vertex_indices"][int(m["vertexOffset"]) + ico["triangles"][int(m["triangleOffset"]): int(m["triangleOffset"]) + 3 * int(m["triangleCount"])].astype(np.int64)]
                           for m in ico["meshlets"]]).astype(np.uint32)
    out["icosphere"] = (pos, soup)
    # disconnected islands + degenerate triangles + a repeated corner: exercises the kd-tree fallback, zero-area cones, dangling triangles
    p1, i1 = S.grid_mesh(9, 9, lambda u, v: (u, v, 0 * u))
    p2, i2 = S.grid_mesh(5, 7, lambda u, v: (u + 3, v, 0.3 * u))
    p3 = rng.uniform(-1, 1, (60, 3)).astype(np.float32) + np.float32(8)
    i3 = rng.integers(0, 60, 180).astype(np.uint32)
    i3[:3] = [5, 5, 9]
    pos = np.concatenate([p1, p2, p3])
    idx = np.concatenate([i1, i2 + p1.shape[0], i3 + p1.shape[0] + p2.shape[0]]).astype(np.uint32)
    out["islands"] = (pos, idx)
    return out


@pytest.mark.parametrize("limits", [(64, 124, 0.0), (64, 124, 0.5), (32, 20, 0.0), (255, 512, 0.25)])
def test_builtin_clusterizer_is_byte_identical_to_the_reference_meshoptimizer(meshopt_ref, limits):
    """a-7: the default meshlet builder (host/clusterizer.cpp) against meshoptimizer built from the reference tree: bound, every
    meshlet record, vertex list, triangle byte (meshopt_buildMeshlets + meshopt_optimizeMeshlet, assets.cpp:331-346) and every field
    of meshopt_computeMeshletBounds, for the reference's limits (64 / 124 / cone weight 0) and three others"""
    from vk_gltf_viewer_b200._native import host_lib
    H = host_lib()
    mv_, mt_, cw = limits
    for name, (pos, idx) in _builder_meshes().items():
        ref = _run_builder(meshopt_ref.meshopt_buildMeshletsBound, meshopt_ref.meshopt_buildMeshlets, meshopt_ref.meshopt_optimizeMeshlet,
                           meshopt_ref.meshopt_computeMeshletBounds, pos, idx, mv_, mt_, cw, bounds_by_value=True)
        got = _run_builder(H.vkvh_meshlets_bound, H.vkvh_meshlets_build, H.vkvh_meshlet_optimize, H.vkvh_meshlet_bounds, pos, idx, mv_, mt_, cw)
        assert got[0] == ref[0], name
        assert np.array_equal(got[1], ref[1]), f"{name}: meshlet records differ"
        assert np.array_equal(got[2], ref[2]), f"{name}: vertex lists differ"
        assert np.array_equal(got[3], ref[3]), f"{name}: triangle bytes differ"
        assert got[4] == ref[4], f"{name}: bounds differ in {sum(a != b for a, b in zip(got[4], ref[4]))} meshlets"


def test_default_scene_builder_uploads_the_reference_partition(meshopt_ref):
    """Scene.add_primitive (no injection) == the same call sequence through the reference's library, buffer for buffer"""
    pos, idx = S.grid_mesh(96, 64, lambda u, v: (u * 2, 0.1 * np.sin(u * 11) * np.cos(v * 5), v))
    s = Scene.new()
    s.add_primitive(pos, idx)
    mine = s.primitive(0)
    set_meshlet_builder(meshopt_ref.meshopt_buildMeshletsBound, meshopt_ref.meshopt_buildMeshlets, meshopt_ref.meshopt_optimizeMeshlet)
    try:
        s2 = Scene.new()
        s2.add_primitive(pos, idx)
        ref = s2.primitive(0)
    finally:
        set_meshlet_builder(None, None, None)
    for k in ("vertex_indices", "triangles", "meshlets"):
        assert mine[k].tobytes() == ref[k].tobytes(), k
    tc = mine["meshlets"]["triangleCount"].astype(np.int64)
    assert 85 < tc[:-1].mean() <= 124   # vertex-limited ~95-triangle meshlets on regular grids, not the Morton packer's ~67


def test_reference_meshoptimizer_can_be_injected(meshopt_ref):
    """the reference's pinned meshoptimizer builds the meshlets (assets.cpp:322-346 call sequence) -> same invariants"""
    set_meshlet_builder(meshopt_ref.meshopt_buildMeshletsBound, meshopt_ref.meshopt_buildMeshlets, meshopt_ref.meshopt_optimizeMeshlet)
    try:
        pos, idx = S.grid_mesh(224, 224, lambda u, v: (u, 0.05 * np.sin(u * 9) * np.cos(v * 7), v))
        s = Scene.new()
        s.add_primitive(pos, idx)
        prim = s.primitive(0)
        got = check_meshlets(prim, idx.size // 3)
        assert np.array_equal(canon(got), canon(idx.reshape(-1, 3)))
        # SURVEY §6 probe: 1,059 meshlets for the 100,352-triangle grid, vertex-limited (64 v, ~94.8 t)
        assert 1000 <= prim["meshlets"].shape[0] <= 1120
    finally:
        set_meshlet_builder(None, None, None)
    s2 = Scene.new()
    s2.add_primitive(pos, idx)
    assert s2.primitive(0)["meshlets"].shape[0] != 0


def test_draw_list_and_transforms_follow_world_cpp():
    s = Scene.new()
    p0 = s.add_primitive(*S.grid_mesh(20, 20, lambda u, v: (u, v, 0 * u)))
    p1 = s.add_primitive(*S.grid_mesh(3, 3, lambda u, v: (u, v, 0 * u)))
    root = s.add_node(-1, translation=(1, 2, 3))                      # transform-only node: takes no transform slot
    a = s.add_node(p0, parent=root, scale=(2, 2, 2))
    s.add_node(p1, parent=a, translation=(0.5, 0, 0))
    s.add_node(p0, translation=(-4, 0, 0))
    s.finalize()
    d, t = s.draws(), s.transforms()
    n0, n1 = s.primitive(0)["meshlets"].shape[0], s.primitive(1)["meshlets"].shape[0]
    assert t.shape[0] == 3 and d.shape[0] == 2 * n0 + n1
    # depth-first: a (p0), its child (p1), then the second root (p0)  — world.cpp:242-264
    assert (d["transformIndex"][:n0] == 0).all() and (d["primitiveIndex"][:n0] == 0).all()
    assert np.array_equal(d["meshletIndex"][:n0], np.arange(n0))
    assert (d["transformIndex"][n0:n0 + n1] == 1).all() and (d["primitiveIndex"][n0:n0 + n1] == 1).all()
    assert (d["transformIndex"][n0 + n1:] == 2).all()
    assert np.allclose(t[0][3], [1, 2, 3, 1]) and np.allclose(np.diag(t[0])[:3], 2)
    assert np.allclose(t[1][3], [2, 2, 3, 1])   # child translation is scaled by the parent
    assert np.allclose(t[2][3], [-4, 0, 0, 1])


def test_int16_positions_expand_exactly():
    q = np.array([[-32768, 0, 32767], [100, -200, 300], [1, 2, 3]], np.int16)
    s = Scene.new()
    s.add_primitive_i16(q, [0, 1, 2])
    assert np.array_equal(s.primitive(0)["vertices"]["position"], q.astype(np.float32))
    s.add_primitive_i16(q, [0, 1, 2], normalized=True)
    want = np.maximum(q.astype(np.float32) / np.float32(32767.0), np.float32(-1.0))  # fastgltf tools.hpp:282-283
    assert np.array_equal(s.primitive(1)["vertices"]["position"], want)


def test_atrium_cfg2_counts():
    s = Scene.atrium(128)
    c = s.counts()
    assert c.triangles_instanced == 262144 and c.primitives == 5 and c.transforms == 28
    # quantised: all positions are integers in int16 range, node matrices carry the 1/32767 scale
    for i in range(c.primitives):
        p = s.primitive(i)["vertices"]["position"]
        assert np.array_equal(p, np.round(p)) and np.abs(p).max() <= 32767


def test_camera_matches_reference_glm(ref_shim):
    """Camera::updateCamera matrix assembly + generateCameraFrustum through the reference's own glm build"""
    rng = np.random.default_rng(0)
    worst = 0
    for _ in range(50):
        eye = rng.uniform(-20, 20, 3).astype(np.float32)
        ctr = rng.uniform(-5, 5, 3).astype(np.float32)
        W, H = int(rng.integers(64, 4000)), int(rng.integers(64, 2200))
        cam = Camera(W, H).look_at(eye, ctr)
        vp = (C.c_float * 16)()
        fr = (C.c_float * 24)()
        ref_shim.ref_camera((C.c_float * 3)(*eye), (C.c_float * 3)(*ctr), (C.c_float * 3)(0, 1, 0), W, H, vp, fr)
        got_vp = np.ctypeslib.as_array(cam.c.viewProjection)
        got_fr = np.ctypeslib.as_array(cam.c.frustum).reshape(-1)
        assert np.array_equal(got_vp.view(np.uint32), np.ctypeslib.as_array(vp).view(np.uint32)), "viewProjection must be bit-identical to glm's"
        assert np.array_equal(got_fr.view(np.uint32), np.ctypeslib.as_array(fr).view(np.uint32)), "frustum planes must be bit-identical"
        worst += 1
    assert worst == 50


def test_node_matrix_matches_fastgltf_math(ref_shim):
    rng = np.random.default_rng(1)
    for _ in range(100):
        parent = rng.uniform(-2, 2, 16).astype(np.float32)
        t = rng.uniform(-5, 5, 3).astype(np.float32)
        q = rng.normal(size=4).astype(np.float32)
        q /= np.linalg.norm(q)
        sc = rng.uniform(0.1, 3, 3).astype(np.float32)
        s = Scene.new()
        p = s.add_primitive([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [0, 1, 2])
        # build parent as a root whose own TRS is identity cannot express an arbitrary matrix; compare the TRS of a root instead
        s.add_node(p, translation=t, rotation=q, scale=sc)
        s.finalize()
        out = (C.c_float * 16)()
        ident = np.eye(4, dtype=np.float32).reshape(-1)
        ref_shim.ref_node_matrix(ident.ctypes.data_as(C.POINTER(C.c_float)), (C.c_float * 3)(*t), (C.c_float * 4)(*q), (C.c_float * 3)(*sc), out)
        assert np.array_equal(s.transforms()[0].reshape(-1).view(np.uint32), np.ctypeslib.as_array(out).view(np.uint32))


def test_camera_prev_matrices_shift():
    cam = Camera(640, 480).look_at((0, 0, 3), (0, 0, 0))
    first = cam.matrix("viewProjection")
    assert np.array_equal(cam.matrix("prevOcclusionViewProjection"), first)  # headless start
    cam.look_at((1, 0, 3), (0, 0, 0))
    assert np.array_equal(cam.matrix("prevOcclusionViewProjection"), first) and np.array_equal(cam.matrix("prevViewProjection"), first)
    assert not np.array_equal(cam.matrix("viewProjection"), first)
    # reverse-Z: a point on the near plane maps to depth 1, far away to ~0 (camera.cpp:38-48)
    vp = cam.matrix("viewProjection").T  # numpy row-major of column-major
    eye = np.array([1, 0, 3], np.float32)
    fwd = (np.zeros(3) - eye) / np.linalg.norm(eye)
    for dist, want in ((0.1, 1.0), (1000.0, 0.0)):
        p = vp @ np.append(eye + fwd * dist, 1.0)
        assert abs(p[2] / p[3] - want) < 1e-3


def test_draw_limit_is_enforced():
    lib_counts = Scene.lattice(2, 1, 2, 8).counts()
    assert lib_counts.draws == lib_counts.meshlets_unique * 4 and lib_counts.transforms == 4
    assert abi.VISBUFFER_CLEAR == 0xFFFFFFFF


def test_builtin_clusterizer_against_the_reference_on_random_meshes(meshopt_ref):
    """tools/diff_meshlet_builder.py (random soups, shuffled grids, degenerate / duplicated triangles, coincident positions, tiny and huge
    coordinates, random limits and cone weights): byte-identical partitions and bounds; a short run here, 1 650 cases in the tool's docstring"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "diff_meshlet_builder.py"), "7", "25"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "mismatches 0" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]

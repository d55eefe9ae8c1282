"""The reference's task- and mesh-shader arithmetic executed AS GLSL — its own text from /root/reference/shaders, compiled by Mesa's GLSL
compiler and run by llvmpipe (tests/llvmpipe_glsl.py) — against the CPU oracle, on BASELINE configs 1-4 at full size.

tests/test_oracle.py pins the same lines by compiling them as C++ against glm; glm associates mat4 * vec4 differently from a shader compiler,
so that test can only demand "differences stay inside the oracle's own ambiguity flags".  Here the text runs through a real GLSL front end and
back end, and the result is stronger:
  * mesh shader (visbuffer.mesh.glsl:44,61,71,90-98): clip positions, determinant(mat3(v0.xyw, v1.xyw, v2.xyw)), determinant(transform) and
    gl_CullPrimitiveEXT are BIT-IDENTICAL to orc_mesh_shader — the oracle's arithmetic policy (mat4 * vec4 left to right, mat4 * mat4
    column by column, cofactor determinants) is exactly Mesa's lowering of the reference's GLSL;
  * task shader (culling.h.glsl whole, visbuffer.task.glsl:50-52,56-61 + the texture fetch through the oracle's sampler): the same class
    (frustum-culled / occluded / visible) for every MeshletDraw, flagged or not.
Skipped where the reference tree or the Mesa library is absent (both are present where the driver runs the CPU suite).
"""
import ctypes as C

import numpy as np
import pytest

from vk_gltf_viewer_b200.scene import Camera

from . import llvmpipe_glsl as G
from . import llvmpipe_lib as LP
from . import oracle_lib as O
from . import scenes as S
from .test_oracle import PIN_CONFIGS

# cfg 5 is cfg 3's patch in a larger lattice at 8K: the same arithmetic on more instances, and a minute of scene building per test — left to
# tests/test_oracle.py's glm pin
GLSL_CONFIGS = {k: v for k, v in PIN_CONFIGS.items() if k != "cfg5"}

pytestmark = pytest.mark.skipif(not (LP.available() and G.available()), reason="needs the Mesa xlib libGL (llvmpipe) and /root/reference")


def per_draw_boxes(scene, draws):
    c, e = np.zeros((draws.shape[0], 3), np.float32), np.zeros((draws.shape[0], 3), np.float32)
    for pi in np.unique(draws["primitiveIndex"]):
        ml = scene.primitive(int(pi))["meshlets"]
        sel = draws["primitiveIndex"] == pi
        c[sel], e[sel] = ml["aabbCenter"][draws["meshletIndex"][sel]], ml["aabbExtents"][draws["meshletIndex"][sel]]
    return c, e


@pytest.mark.parametrize("name", sorted(GLSL_CONFIGS))
def test_task_shader_text_run_as_glsl_gives_the_oracles_classes(name):
    make, (W, H) = PIN_CONFIGS[name]
    scene = make()
    cam = Camera(W, H).look_at(*scene.default_view(0, 64))
    pc = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    O.frame(pc, tg, two_pass=False)                       # a real previous-frame pyramid
    cam.look_at(*scene.default_view(1, 64))
    all_draws = scene.draws()
    step = max(1, -(-all_draws.shape[0] // 300000))       # at most 300 k draws: cfg 3 every 4th, cfg 5 every 37th
    pick = np.arange(0, all_draws.shape[0], step)
    draws = all_draws[pick]
    N = draws.shape[0]
    xf = scene.transforms()[draws["transformIndex"]]
    boxC, boxE = per_draw_boxes(scene, draws)
    Wc = 2048
    tex = {"xf0": G.to_tex(xf[:, 0, :], Wc), "xf1": G.to_tex(xf[:, 1, :], Wc), "xf2": G.to_tex(xf[:, 2, :], Wc), "xf3": G.to_tex(xf[:, 3, :], Wc),
           "boxC": G.to_tex(boxC, Wc), "boxE": G.to_tex(boxE, Wc)}
    Hc = tex["xf0"].shape[0]
    lp, fs, L = LP.instance(), G.task_shader(), O.lib()
    frustum = np.ctypeslib.as_array(cam.c.frustum).reshape(6, 4).copy()
    flagged = O.AMBIG_FRUSTUM | O.AMBIG_HIZ | O.AMBIG_LEVEL | O.AMBIG_FOOTPRINT | O.CROSSES_CAMERA
    report = {}
    for vp_select, matrix in ((0, "prevOcclusionViewProjection"), (1, "viewProjection")):   # pass A, and pass B's rule (current matrix)
        st = O.cull(pc, W, H, tg.pyramid, vp_select)[0][pick]
        uni = {"camera.frustum": frustum, "camera.prevOcclusionViewProjection": cam.matrix(matrix), "pyramidSize": (tg.layout[0][1], tg.layout[0][2])}
        o0 = lp.compute(fs, Wc, Hc, tex, dict(uni, mode=0)).reshape(-1, 4)[:N]
        o1 = lp.compute(fs, Wc, Hc, tex, dict(uni, mode=1)).reshape(-1, 4)[:N]
        # mat4 * vec4 as Mesa lowers it == the oracle's policy ((c0 x + c1 y) + c2 z) + c3: the world-space box centre, bit for bit
        want = ((xf[:, 0, :3] * boxC[:, 0:1] + xf[:, 1, :3] * boxC[:, 1:2]) + xf[:, 2, :3] * boxC[:, 2:3]) + xf[:, 3, :3] * np.float32(1)
        assert np.array_equal(o1[:, 1:4].view(np.uint32), want.astype(np.float32).view(np.uint32))
        status = np.zeros(N, np.uint8)
        for i in np.nonzero(o0[:, 0] > 0.5)[0]:           # the texture fetch + comparison (task.glsl:62-64) with the oracle's min sampler
            lvl = o0[i, 1]
            level = 0 if not (lvl > 0) else int(min(lvl, 16, tg.levels - 1))   # sampler lod clamp (application.cpp:451-452), NaN -> 0
            off, w, h = tg.layout[level]
            depth = L.orc_sample_min(tg.pyramid[off:].ctypes.data, w, h, o0[i, 2], o0[i, 3], None)
            status[i] = O.VISIBLE if depth < o1[i, 0] else O.OCCLUDED
        differ = (st & O.STATUS_MASK) != status
        report[vp_select] = dict(draws=N, differ=int(differ.sum()), unflagged=int((differ & ((st & flagged) == 0)).sum()),
                                 classes=np.bincount(status, minlength=3).tolist())
    print(f"\n{name}: reference task-shader GLSL on llvmpipe vs oracle: {report}")
    for r in report.values():
        assert r["unflagged"] == 0 and r["differ"] <= 2, r


def triangles_of(scene, draws, ids):
    """per triangle of the MeshletDraws `ids`: node matrix [n,4,4], three object-space positions [n,3] each, and (k, local index triple) to find
    the oracle's per-meshlet outputs again"""
    T, P, ref, single = [], [[], [], []], [], []
    cache = {}
    mats = scene.materials()
    for k, d in enumerate(ids):
        pi = int(draws[d]["primitiveIndex"])
        if pi not in cache:
            p = scene.primitive(pi)
            cache[pi] = (p["meshlets"], p["triangles"], p["vertex_indices"], p["vertices"]["position"], int(mats[p["header"].materialIndex]["doubleSided"]) == 0)
        meshlets, tris, vidx, pos, one_sided = cache[pi]
        ml = meshlets[int(draws[d]["meshletIndex"])]
        tc, to, vo = int(ml["triangleCount"]), int(ml["triangleOffset"]), int(ml["vertexOffset"])
        tri = tris[to:to + 3 * tc].reshape(tc, 3).astype(np.int64)
        vp = pos[vidx[vo:vo + int(ml["vertexCount"])]]
        for j in range(3):
            P[j].append(vp[tri[:, j]])
        T.append(np.repeat(scene.transforms()[draws[d]["transformIndex"]][None], tc, 0))
        ref.append((k, tri))
        single.append(np.full(tc, one_sided))
    return np.concatenate(T), [np.concatenate(p) for p in P], ref, np.concatenate(single)


MESH_CASES = dict(GLSL_CONFIGS)
MESH_CASES["mirrored"] = (lambda: S.mirrored_instances(), (320, 240))
MESH_CASES["blobs_trs"] = (lambda: S.cull_stress(seed=11, instances=61), (640, 360))


@pytest.mark.parametrize("name", sorted(MESH_CASES))
def test_mesh_shader_text_run_as_glsl_is_bit_identical_to_the_oracle(name):
    make, (W, H) = MESH_CASES[name]
    scene = make()
    cam = Camera(W, H).look_at(*scene.default_view(1, 64)) if name in PIN_CONFIGS else S.camera(W, H)
    pc = scene.host_push_constants(cam)
    draws = scene.draws()
    ids = np.unique(np.linspace(0, draws.shape[0] - 1, min(draws.shape[0], 2500)).astype(np.uint32))
    clip, cull, det, tdet = O.mesh_shader(pc, ids)[:4]
    T, P, ref, single_sided = triangles_of(scene, draws, ids)
    n = T.shape[0]
    Wc = 1024
    tex = {"xf0": G.to_tex(T[:, 0, :], Wc), "xf1": G.to_tex(T[:, 1, :], Wc), "xf2": G.to_tex(T[:, 2, :], Wc), "xf3": G.to_tex(T[:, 3, :], Wc),
           "p0": G.to_tex(P[0], Wc), "p1": G.to_tex(P[1], Wc), "p2": G.to_tex(P[2], Wc)}
    Hc = tex["xf0"].shape[0]
    lp, fs = LP.instance(), G.mesh_shader()
    out = [lp.compute(fs, Wc, Hc, tex, {"viewProjection": cam.matrix("viewProjection"), "mode": m}).reshape(-1, 4)[:n] for m in range(4)]
    oc = [np.concatenate([clip[k][tri[:, j]] for k, tri in ref]) for j in range(3)]
    od = np.concatenate([det[k, :tri.shape[0]] for k, tri in ref])
    ocull = np.concatenate([cull[k, :tri.shape[0]] for k, tri in ref])
    otd = np.concatenate([np.full(tri.shape[0], tdet[k], np.float32) for k, tri in ref])
    for j in range(3):   # mesh.glsl:44 + :61 — gl_Position of the triangle's three vertices
        assert np.array_equal(out[j].view(np.uint32), oc[j].view(np.uint32)), (name, "clip", j)
    assert np.array_equal(out[3][:, 1].view(np.uint32), otd.view(np.uint32)), "determinant(transformMatrix), mesh.glsl:71"
    # mesh.glsl:93: determinant(mat3(v0, v1, v2)) — every bit, not only the sign — and :94-98 the decision; the reference evaluates them for
    # single-sided materials only (:86), and so does the oracle
    m = single_sided & ~np.isnan(od)
    assert np.array_equal(out[3][m, 0].view(np.uint32), od[m].view(np.uint32))
    assert np.array_equal(out[3][single_sided, 2] > 0.5, ocull[single_sided] == 1)
    assert (ocull[~single_sided] != 1).all()
    print(f"\n{name}: {n} triangles of {ids.size} MeshletDraws: clip positions, both determinants and gl_CullPrimitiveEXT bit-identical "
          f"({int(single_sided.sum())} single-sided, {int((ocull == 1).sum())} culled)")

"""Shared by tests/test_llvmpipe.py and tests/golden/make_llvmpipe_golden.py: the inputs handed to llvmpipe and to the CPU oracle, and
the comparison of what comes back.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

from vk_gltf_viewer_b200 import abi
from vk_gltf_viewer_b200.scene import Camera, Scene

from . import oracle_lib as O
from . import scenes as S


def identity_camera(W, H):
    """viewProjection = I: positions are clip-space coordinates with w = 1 (mat4 * vec4(p, 1) is then exact)"""
    cam = Camera(W, H).look_at((0, 0, 3), (0, 0, 0))
    eye = (C.c_float * 16)(*np.eye(4, dtype=np.float32).reshape(-1))
    for f in ("prevViewProjection", "prevOcclusionViewProjection", "viewProjection", "occlusionViewProjection"):
        setattr(cam.c, f, eye)
    return cam


def soup_scene(P):
    """one double-sided primitive of independent triangles, positions P [3n, 3]"""
    s = Scene.new()
    m = s.add_material(double_sided=True)
    s.add_node(s.add_primitive(np.ascontiguousarray(P, np.float32), np.arange(P.shape[0]), m))
    s.finalize()
    return s


def lattice_positions(rng, W, H, ntri):
    """clip-space triangles (w = 1) whose vertices lie exactly on the 1/256-pixel lattice, half of them exactly on pixel centres: every
    edge-through-a-pixel-centre case of the fill rule, with viewport transform and snapping exact in any implementation.
    W, H powers of two (so that the NDC coordinates are exact)."""
    sub = np.stack([rng.integers(0, W * 256 + 1, (ntri, 3)), rng.integers(0, H * 256 + 1, (ntri, 3))], -1)
    cent = np.stack([rng.integers(0, W, (ntri, 3)), rng.integers(0, H, (ntri, 3))], -1) * 256 + 128
    sub = np.where(rng.random((ntri, 3, 2)) < 0.5, cent, sub)
    P = np.zeros((ntri * 3, 3), np.float32)
    P[:, 0] = (sub[..., 0].reshape(-1) / 256.0) / (W / 2) - 1.0
    P[:, 1] = (sub[..., 1].reshape(-1) / 256.0) / (H / 2) - 1.0
    P[:, 2] = np.repeat(np.linspace(0.1, 0.9, ntri), 3)       # constant depth per triangle: the depth image is exact too
    return P


def float_positions(rng, ntri, scale, inside=False):
    """arbitrary float vertices, w = 1, depth varying per vertex; inside: no triangle reaches the viewport's border"""
    P = np.zeros((ntri * 3, 3), np.float32)
    lim = 1 - scale - 1e-3 if inside else 1
    c = rng.uniform(-lim, lim, (ntri, 1, 2))
    P[:, :2] = (c + rng.uniform(-scale, scale, (ntri, 3, 2))).reshape(-1, 2)
    P[:, 2] = rng.uniform(0.05, 0.95, ntri * 3)
    return P


# (name, builder of (scene, camera), (W, H)): geometry with perspective, near-plane clipping, mirrored instances, coplanar instances
SCENE_CASES = {
    "icosphere": (lambda W, H: (lambda s: (s, s.default_camera(W, H)))(Scene.icosphere(57)), (640, 480)),
    "atrium": (lambda W, H: (lambda s: (s, s.default_camera(W, H)))(Scene.atrium(32)), (640, 360)),
    "city": (lambda W, H: (lambda s: (s, s.default_camera(W, H)))(Scene.city(6, 5, 2000, 0x5EED0004)), (640, 360)),
    "lattice": (lambda W, H: (lambda s: (s, s.default_camera(W, H)))(Scene.lattice(3, 3, 3, 60, 0x5EED0003)), (960, 540)),
    "atrium_cfg2_full": (lambda W, H: (lambda s: (s, s.default_camera(W, H)))(Scene.atrium(128)), (1920, 1080)),   # BASELINE cfg 2 as benchmarked
    "lattice_4k": (lambda W, H: (lambda s: (s, s.default_camera(W, H)))(Scene.lattice(4, 4, 4, 100, 0x5EED0003)), (3840, 2160)),
    "ground_clipped": (lambda W, H: (S.ground_plane(8, 30.0, -1.0), Camera(W, H).look_at((0, 0, 0), (0, 0, -1))), (400, 300)),
    "mirrored": (lambda W, H: (S.mirrored_instances(), Camera(W, H).look_at((0, 0, 4), (0, 0, 0))), (320, 240)),
    "coplanar": (lambda W, H: (S.coplanar_overlap(), S.camera(W, H)), (160, 120)),
}


def oracle_triangles(scene, pc, facing_from=None):
    """what the mesh shader hands to the fixed-function rasteriser, per the oracle (orc_mesh_shader = visbuffer.mesh.glsl:43-102, itself pinned
    against the reference's text): clip-space vertices [3n, 4] of every triangle the facing test keeps, in draw / triangle order, and the
    packVisBuffer id of each ([3n], as float32: exact below 2^24)."""
    n = pc.meshletDrawCount
    assert n < (1 << 17), "ids must stay exact in float32"
    all_ids = np.arange(n, dtype=np.uint32)
    clip, cull = O.mesh_shader(pc, all_ids)[:2]
    if facing_from is not None:       # positions under `pc`'s matrices, the facing decision of another camera state (motion vectors: same triangles)
        cull = O.mesh_shader(facing_from, all_ids)[1]
    draws = scene.draws()
    prims, V, I = {}, [], []
    for d in range(n):
        pi = int(draws[d]["primitiveIndex"])
        if pi not in prims:
            prims[pi] = scene.primitive(pi)
        ml = prims[pi]["meshlets"][int(draws[d]["meshletIndex"])]
        tc, to = int(ml["triangleCount"]), int(ml["triangleOffset"])
        tri = prims[pi]["triangles"][to:to + 3 * tc].reshape(tc, 3).astype(np.int64)
        keep = cull[d, :tc] == 0
        V.append(clip[d][tri[keep]].reshape(-1, 4))
        I.append(np.repeat(((d << 7) | np.nonzero(keep)[0]).astype(np.float32), 3))
    return np.concatenate(V), np.concatenate(I)


def oracle_images(pc, W, H):
    tg = O.Targets(W, H)
    O.raster(pc, tg, np.arange(pc.meshletDrawCount, dtype=np.uint32))
    return tg


def compare(tg, lp_ids, lp_depth):
    """oracle targets vs llvmpipe's id (float, -1 = nothing drawn) and depth images"""
    oid = tg.ids_ref
    ocov, lcov = oid != abi.VISBUFFER_CLEAR, lp_ids >= 0
    both = ocov & lcov
    same = both & (oid == np.where(lcov, lp_ids, 0).astype(np.uint32))
    dd = np.abs(tg.depth - lp_depth)
    q = np.quantile(dd[same], [0.5, 0.99, 1.0]) if same.any() else np.zeros(3)
    return dict(covered=int(ocov.sum()), coverage_differs=int((ocov ^ lcov).sum()), id_differs=int((both & ~same).sum()),
                depth_bits_equal=float((tg.depth[same].view(np.uint32) == lp_depth[same].view(np.uint32)).mean()) if same.any() else 1.0,
                depth_q50=float(q[0]), depth_q99=float(q[1]), depth_max=float(q[2]),
                depth_at_id_differs=float(dd[both & ~same].max()) if (both & ~same).any() else 0.0)


# ------------------------------------------------------------------------------------------ the sampler
def indicator(w, h, axis, k, value):
    """image that is `value` where (x or y) % 3 == k and 0 elsewhere: a 2-texel footprint {i, i+1} touches exactly two of the three classes"""
    im = np.zeros((h, w), np.float32)
    if axis == 0:
        im[:, np.arange(w) % 3 == k] = value
    else:
        im[np.arange(h) % 3 == k, :] = value
    return im


def oracle_footprint_classes(w, h, uv):
    """[n, 2, 3] bool: does the oracle's min sampler read a texel of class k along x / y at coordinate uv[i]"""
    L = O.lib()
    out = np.zeros((uv.shape[0], 2, 3), bool)
    for axis in (0, 1):
        for k in range(3):
            im = indicator(w, h, axis, k, -1.0)
            for i in range(uv.shape[0]):
                out[i, axis, k] = L.orc_sample_min(im.ctypes.data, w, h, uv[i, 0], uv[i, 1], None) == -1.0
    return out


def llvmpipe_footprint_classes(lp, w, h, uv):
    """the same from llvmpipe's LINEAR filter: a class is in the footprint iff it gets a non-zero weight"""
    out = np.zeros((uv.shape[0], 2, 3), bool)
    for axis in (0, 1):
        for k in range(3):
            out[:, axis, k] = lp.sample_linear(indicator(w, h, axis, k, 1.0), uv) != 0
    return out


def sampler_coords(rng, w, h, n):
    """random coordinates (some outside [0,1]: CLAMP_TO_EDGE), texel centres (fraction 0 after the -0.5 shift) and texel edges (fraction 0.5)"""
    uv = rng.uniform(-0.1, 1.1, (n, 2)).astype(np.float32)
    t = n // 3
    xs, ys = rng.integers(0, w + 1, t), rng.integers(0, h + 1, t)
    uv[:t, 0], uv[:t, 1] = ((xs + 0.5) / w).astype(np.float32), ((ys + 0.5) / h).astype(np.float32)
    uv[t:2 * t, 0], uv[t:2 * t, 1] = (xs / w).astype(np.float32), (ys / h).astype(np.float32)
    return uv


def hiz_level_sizes(W, H):
    """(source size, destination size) of every hiz_reduce dispatch (application.cpp:964-979) that writes something"""
    out, sw, sh = [], W, H
    i = 1
    while (W >> i) > 0 and (H >> i) > 0:
        out.append(((sw, sh), (W >> i, H >> i)))
        sw, sh = max(1, (W >> 1) >> (i - 1)), max(1, (H >> 1) >> (i - 1))
        i += 1
    return out

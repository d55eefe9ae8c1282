"""The CPU oracle's fixed-function parts against llvmpipe — Mesa's software rasteriser, the engine underneath lavapipe.

SURVEY §8c names lavapipe as the oracle one would want for the stages that are Vulkan fixed function (no reference code exists for them):
triangle coverage and the depth test (pipeline state application.cpp:326-340,772-841, pipeline_builder.cpp:225-277) and the sampler's texel
footprint (application.cpp:438-453).  lavapipe itself cannot run here, but this image carries llvmpipe inside Nsight Compute (Mesa 18.1.9's
xlib libGL); tests/llvmpipe_lib.py drives it without an X server.  Same gallium rasteriser and texture unit as lavapipe, reached through
OpenGL: a user FBO with glClipControl(GL_LOWER_LEFT, GL_ZERO_TO_ONE) gives Vulkan's conventions in memory order (test_conventions).

What is held to llvmpipe here:
  * coverage — which pixels a triangle produces fragments for (pixel centres, top-left rule, 8 sub-pixel bits, clipping): identical images
    on lattice-exact triangles, on arbitrary float triangles and on perspective scenes with near-plane clipping and mirrored instances;
  * the depth test — GREATER_OR_EQUAL, clear 0.0, later fragment wins a tie: the same triangle id in every pixel (up to pixels where two
    surfaces are closer than the two interpolators' rounding);
  * depth values — Vulkan leaves the interpolation arithmetic to the implementation; llvmpipe evaluates a plane through the unsnapped
    vertices, the oracle barycentrics of the snapped ones: they agree to ~1e-7 (median) / ~2e-5 (max) on meshes, which is reported, not
    required bit for bit;
  * the sampler — the set of texels with a non-zero LINEAR weight (what the MIN reduction runs over) at random coordinates, texel centres
    and edges, outside [0,1], and at every coordinate hiz_reduce.comp.glsl samples for whole mip chains (odd sizes included); the mip a
    given integer lod selects.
Where the Mesa library is missing these tests skip; tests/test_golden.py::test_llvmpipe_* then still checks the oracle against llvmpipe's
committed outputs (tests/golden/llvmpipe.npz).
"""
import numpy as np
import pytest

from vk_gltf_viewer_b200 import abi

from . import llvmpipe_cases as K
from . import llvmpipe_lib as LP
from . import oracle_lib as O

pytestmark = pytest.mark.skipif(not LP.available(), reason="no Mesa xlib libGL (llvmpipe) in this image")


@pytest.fixture(scope="module")
def lp():
    return LP.instance()


def test_conventions(lp):
    """A rectangle whose four edges pass exactly through pixel centres: with Vulkan's rules the top row and the left column of centres belong
    to it, the bottom row and the right column do not — in MEMORY order (row 0 first).  This is what makes the comparisons below mean
    something: llvmpipe reached through GL with a user FBO and the default clip origin has exactly that orientation, depth = z/w."""
    assert "llvmpipe" in lp.renderer
    W = H = 16
    x0, x1, y0, y1 = 2.5, 6.5, 3.5, 9.5
    n = lambda p: p / 8.0 - 1.0
    quad = np.array([[n(x0), n(y0), 0.5, 1], [n(x1), n(y0), 0.5, 1], [n(x1), n(y1), 0.5, 1],
                     [n(x0), n(y0), 0.5, 1], [n(x1), n(y1), 0.5, 1], [n(x0), n(y1), 0.5, 1]], np.float32)
    ids, depth = lp.raster(W, H, quad, np.zeros(6, np.float32))
    ys, xs = np.nonzero(ids >= 0)
    assert (ys.min(), ys.max(), xs.min(), xs.max()) == (3, 8, 2, 5)
    assert (depth[ids >= 0] == 0.5).all() and (depth[ids < 0] == 0.0).all()
    # and the oracle says the same about the same rectangle
    s = K.soup_scene(quad[:, :3])
    tg = K.oracle_images(s.host_push_constants(K.identity_camera(W, H)), W, H)
    assert np.array_equal(tg.ids_ref != abi.VISBUFFER_CLEAR, ids >= 0)


@pytest.mark.parametrize("seed,size,ntri", [(5, 64, 300), (6, 64, 1500), (7, 256, 4000), (8, 1024, 6000)])
def test_coverage_and_depth_test_on_lattice_exact_triangles(lp, seed, size, ntri):
    """Vertices on the 1/256-pixel lattice, half of them exactly on pixel centres, w = 1: nothing is left to rounding, so every image must be
    identical — coverage, the winning id under the >= depth test in submission order, and (constant depth per triangle) the depth bits."""
    rng = np.random.default_rng(seed)
    W = H = size
    s = K.soup_scene(K.lattice_positions(rng, W, H, ntri))
    pc = s.host_push_constants(K.identity_camera(W, H))
    V, I = K.oracle_triangles(s, pc)
    assert V.shape[0] == 3 * ntri                                  # double-sided: the facing test keeps every triangle
    tg = K.oracle_images(pc, W, H)
    ids, depth = lp.raster(W, H, V, I)
    want = np.where(tg.ids_ref == abi.VISBUFFER_CLEAR, -1.0, tg.ids_ref.astype(np.float32))
    assert np.array_equal(ids, want)
    assert np.array_equal(depth.view(np.uint32), tg.depth.view(np.uint32))
    assert (ids >= 0).mean() > 0.5


@pytest.mark.parametrize("seed,W,H,ntri,scale", [(15, 64, 64, 400, 0.45), (16, 257, 193, 3000, 0.3), (17, 640, 480, 20000, 0.05), (18, 1920, 1080, 60000, 0.02),
                                                 (19, 7680, 4320, 40000, 0.004)])
@pytest.mark.parametrize("inside", [True, False])
def test_coverage_on_arbitrary_float_triangles(lp, seed, W, H, ntri, scale, inside):
    """Arbitrary fp32 vertices (w = 1), interpenetrating, depth varying over each triangle: now the viewport transform's and the snap's rounding
    take part.  Triangles that stay inside the viewport: IDENTICAL coverage.  Triangles that cross its border: llvmpipe clips them against the
    x = +-w / y = +-w planes and rasterises the clipped polygon, whose new vertices are rounded (the oracle — like a GPU with a guard band —
    rasterises the original triangle, which is the exact result); there a few pixels in a million differ, every one of them in a triangle that
    crosses the border.  Where surfaces intersect, the winner depends on interpolated depth: ids differ in a few pixels per ten thousand."""
    rng = np.random.default_rng(seed)
    P = K.float_positions(rng, ntri, scale, inside)
    s = K.soup_scene(P)
    pc = s.host_push_constants(K.identity_camera(W, H))
    V, I = K.oracle_triangles(s, pc)
    tg = K.oracle_images(pc, W, H)
    ids, depth = lp.raster(W, H, V, I)
    r = K.compare(tg, ids, depth)
    print(f"\nfloat triangles {W}x{H}, {ntri} triangles, inside={inside}: {r}")
    if inside:
        assert r["coverage_differs"] == 0
    else:
        assert r["coverage_differs"] <= 2 + 1e-5 * r["covered"]
        crossing = (np.abs(P[:, :2]).reshape(ntri, 6).max(1) >= 1.0)          # triangles with a vertex outside the viewport
        ys, xs = np.nonzero((tg.ids_ref != abi.VISBUFFER_CLEAR) ^ (ids >= 0))
        sx, sy = (P[:, 0].reshape(ntri, 3) + 1) * W / 2, (P[:, 1].reshape(ntri, 3) + 1) * H / 2
        for x, y in zip(xs, ys):                                              # every differing pixel lies in the box of a crossing triangle
            near = (sx.min(1) <= x + 1) & (sx.max(1) >= x) & (sy.min(1) <= y + 1) & (sy.max(1) >= y)
            assert (near & crossing).any(), (x, y)
    assert r["id_differs"] <= 1e-3 * r["covered"]
    assert r["depth_q50"] < 1e-4 and r["depth_q99"] < 2e-3


@pytest.mark.parametrize("seed,ntri,spread", [(4, 60, 2.0), (5, 200, 2.0), (6, 1000, 3.0), (7, 3000, 4.0)])
def test_random_triangles_around_the_eye(lp, seed, ntri, spread):
    """View-space triangles scattered around the camera: a fifth of them cross the near plane, a sixth lie behind the eye (w <= 0), many leave
    through the sides.  The oracle clips only against near / far / a guard band (and only the triangles that need it), llvmpipe against all six
    planes: what is left on the screen must be the same — coverage identical, the same id up to a few pixels where surfaces intersect."""
    from vk_gltf_viewer_b200.scene import Camera
    rng = np.random.default_rng(seed)
    W, H = 256, 192
    P = np.zeros((ntri * 3, 3), np.float32)
    c = rng.uniform(-spread, spread, (ntri, 1, 3))
    c[:, :, 2] = rng.uniform(-6, 2, (ntri, 1))
    P[:] = (c + rng.uniform(-1.5, 1.5, (ntri, 3, 3))).reshape(-1, 3)
    s = K.soup_scene(P)
    pc = s.host_push_constants(Camera(W, H).look_at((0, 0, 0), (0, 0, -1)))
    V, I = K.oracle_triangles(s, pc)
    w = V[:, 3].reshape(-1, 3)
    assert ((w.min(1) <= 0.1) & (w.max(1) > 0.1)).sum() > ntri // 10 and (w.max(1) <= 0).sum() > ntri // 12
    tg = K.oracle_images(pc, W, H)
    r = K.compare(tg, *lp.raster(W, H, V, I))
    print(f"\n{ntri} triangles around the eye: {r}")
    assert r["coverage_differs"] == 0 and r["covered"] > 0.4 * W * H
    assert r["id_differs"] <= 2 + 2e-4 * r["covered"]


@pytest.mark.parametrize("seed,ntri", [(8, 40), (9, 150)])
def test_random_triangles_around_the_far_plane(lp, seed, ntri):
    """large triangles scattered around z = far (reverse-Z: depth 0): more than half of them cross the far plane, a fifth lie beyond it.  A
    clipped edge runs between vertices the clipper created, and two clippers round those differently (llvmpipe also clips these triangles at the
    sides, where the oracle's guard band does not clip at all): a pixel or two on such an edge may differ — 1 of 38 430 here."""
    from vk_gltf_viewer_b200.scene import Camera
    rng = np.random.default_rng(seed)
    W, H = 256, 192
    P = np.zeros((ntri * 3, 3), np.float32)
    c = rng.uniform(-700, 700, (ntri, 1, 3))
    c[:, :, 2] = rng.uniform(-1300, -700, (ntri, 1))
    P[:] = (c + rng.uniform(-400, 400, (ntri, 3, 3))).reshape(-1, 3)
    s = K.soup_scene(P)
    pc = s.host_push_constants(Camera(W, H).look_at((0, 0, 0), (0, 0, -1)))
    V, I = K.oracle_triangles(s, pc)
    z = V[:, 2].reshape(-1, 3)
    assert ((z.min(1) < 0) & (z.max(1) > 0)).sum() > ntri // 3 and (z.max(1) < 0).sum() > ntri // 10
    tg = K.oracle_images(pc, W, H)
    r = K.compare(tg, *lp.raster(W, H, V, I))
    assert r["coverage_differs"] <= 2 and r["covered"] > 0.2 * W * H and r["id_differs"] <= 2 + 2e-4 * r["covered"]


@pytest.mark.parametrize("name", sorted(K.SCENE_CASES))
def test_scenes_match_llvmpipe(lp, name):
    """Meshes through the whole oracle path (mesh shader arithmetic -> trivial reject -> clip -> snap -> edge functions -> depth test) against
    llvmpipe fed with the same clip-space triangles: identical coverage, the same id in (all but a handful of) pixels, depth within 5e-5."""
    make, (W, H) = K.SCENE_CASES[name]
    scene, cam = make(W, H)
    pc = scene.host_push_constants(cam)
    V, I = K.oracle_triangles(scene, pc)
    tg = K.oracle_images(pc, W, H)
    r = K.compare(tg, *lp.raster(W, H, V, I))
    print(f"\n{name} {W}x{H}, {V.shape[0] // 3} triangles after the facing test: {r}")
    assert r["covered"] > 1000
    assert r["coverage_differs"] == 0
    assert r["id_differs"] <= 8 + 1e-5 * r["covered"] and r["depth_at_id_differs"] < 1e-5     # surfaces that touch (the atrium's walls)
    assert r["depth_max"] < 5e-5 and r["depth_q50"] < 2e-6


@pytest.mark.parametrize("size", [(1, 1), (2, 1), (5, 3), (16, 16), (67, 120), (960, 540), (1920, 1080)])
def test_sampler_footprint(lp, size):
    """The texels the oracle's min sampler reads == the texels llvmpipe's LINEAR filter gives a non-zero weight (u = coord * size - 0.5, texels
    floor(u) and floor(u) + 1, the second without weight when u is whole, indices clamped to the edge), at random coordinates inside and outside
    [0, 1], at texel centres and at texel edges."""
    w, h = size
    rng = np.random.default_rng(w * 131 + h)
    uv = K.sampler_coords(rng, w, h, 1500)
    assert np.array_equal(K.oracle_footprint_classes(w, h, uv), K.llvmpipe_footprint_classes(lp, w, h, uv))


@pytest.mark.parametrize("res", [(640, 480), (1920, 1080), (3840, 2160), (1001, 777), (97, 33)])
@pytest.mark.parametrize("coordinates", ["ieee", "glsl"])
def test_hiz_reduce_footprints_over_whole_mip_chains(lp, res, coordinates):
    """Every coordinate hiz_reduce.comp.glsl:28 samples, for every dispatch of the chain (odd source sizes included: SURVEY D5, the pyramid is
    not conservative there): the oracle's reduce reads a texel class iff llvmpipe's LINEAR filter weights it.
    coordinates = "ieee": (pos + 0.5) / imageSize with IEEE division, as the oracle evaluates it; "glsl": the reference's own expression evaluated
    by Mesa's GLSL compiler, which lowers the division to a multiplication by the reciprocal — about a third of the coordinates differ in the
    last bit (GPUs differ likewise: Vulkan allows 2.5 ulp), and the footprints must not care.  The one coordinate that sits exactly on a
    footprint change — the centre texel of an odd level, (D/2 + 0.5) / D * (2D + 1) - 0.5 = a whole number — comes out as exactly 0.5 either way."""
    from . import llvmpipe_glsl as G
    if coordinates == "glsl" and not G.available():
        pytest.skip("needs /root/reference")
    W, H = res
    checked = differing = 0
    for (sw, sh), (dw, dh) in K.hiz_level_sizes(W, H):
        x, y = np.meshgrid(np.arange(dw, dtype=np.float32), np.arange(dh, dtype=np.float32))
        uv = np.stack([(x + np.float32(0.5)) / np.float32(dw), (y + np.float32(0.5)) / np.float32(dh)], -1).reshape(-1, 2).astype(np.float32)
        if coordinates == "glsl":
            got_uv = lp.compute(G.hiz_coordinates_shader(), dw, dh, {}, {"pushConstants.imageSize": np.array([dw, dh], np.uint32)})[:, :, :2].reshape(-1, 2)
            differing += int((got_uv.view(np.uint32) != uv.view(np.uint32)).sum())
            assert np.abs(got_uv - uv).max() <= 1.2e-7
            uv = np.ascontiguousarray(got_uv)
        for axis in (0, 1):
            for k in range(3):
                tg = O.Targets(sw, sh)
                tg.depth[:] = K.indicator(sw, sh, axis, k, -1.0)
                O.hiz(tg)                                               # mip 0 of a (sw, sh) image = this dispatch
                got = tg.mip(0)[:dh, :dw] == -1.0
                want = (lp.sample_linear(K.indicator(sw, sh, axis, k, 1.0), uv) != 0).reshape(dh, dw)
                assert np.array_equal(got, want), (res, (sw, sh), axis, k)
                checked += dw * dh
    assert checked > 0
    if coordinates == "glsl" and max(W, H) > 200:
        assert differing > 0, "expected Mesa's reciprocal-multiply to differ from IEEE division somewhere"


def test_integer_lod_selects_the_clamped_mip(lp):
    """textureLod with the task shader's floor()ed level (visbuffer.task.glsl:59-62) and mipmapMode NEAREST: level = clamp(lod, 0, last mip)"""
    mips = [np.full((max(1, 135 >> k), max(1, 240 >> k)), float(k), np.float32) for k in range(8)]
    uv = np.array([[0.3, 0.6]], np.float32)
    for lod, want in [(-5, 0), (-1, 0), (0, 0), (1, 1), (3, 3), (7, 7), (8, 7), (16, 7), (40, 7)]:
        assert lp.sample_linear(mips, uv, lod)[0] == want, lod

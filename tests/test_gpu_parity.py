"""GPU parity: the CUDA path (through the C ABI, include/vkv.h) against the CPU oracle on identical inputs.

Bar: bit-exact visible-meshlet sets, bit-exact 64-bit visbuffer (depth bits and IDs; exact-depth ties are
resolved lowest-id-wins by atomicMin, the oracle reports them and also carries the reference's last-writer-wins image),
bit-exact HiZ mips.
"""
import numpy as np
import pytest

from tests import oracle_lib as O
from tests import scenes as S
from vk_gltf_viewer_b200 import abi, api
from vk_gltf_viewer_b200.scene import Camera, Scene

pytestmark = pytest.mark.gpu


def compare_frame(r: api.Renderer, tg: O.Targets, out, two_pass, label=""):
    """compare everything observable after one frame"""
    visA = np.sort(r.read_visible(0))
    assert np.array_equal(visA, out["visibleA"]), f"{label}: pass-A visible set differs: gpu {visA.size} vs oracle {out['visibleA'].size}"
    if two_pass:
        visB = np.sort(r.read_visible(1))
        assert np.array_equal(visB, out["visibleB"]), f"{label}: pass-B visible set differs: gpu {visB.size} vs oracle {out['visibleB'].size}"
    vis = r.read_visbuffer64()
    want = tg.vis64()
    bad = vis != want
    assert not bad.any(), f"{label}: {int(bad.sum())} of {bad.size} visbuffer keys differ (first at {np.argwhere(bad)[0]}: gpu {vis[bad][0]:#x} oracle {want[bad][0]:#x})"
    # the reference's own images: depth identical everywhere, IDs identical away from exact-depth ties
    assert np.array_equal(r.read_depth().view(np.uint32), tg.depth.view(np.uint32))
    ids = r.read_ids()
    notie = tg.tie == 0
    assert np.array_equal(ids[notie], tg.ids_ref[notie])
    pyr = r.read_pyramid()
    assert np.array_equal(pyr.view(np.uint32), tg.pyramid.view(np.uint32)), f"{label}: pyramid differs in {(pyr.view(np.uint32) != tg.pyramid.view(np.uint32)).sum()} texels"


def run_views(scene: Scene, W, H, views, two_pass=False):
    """render a sequence of camera positions; frame k culls against frame k-1's pyramid, on both sides"""
    cam = Camera(W, H)
    cam.look_at(*views[0])
    r = api.Renderer(W, H)
    pc_dev = r.upload_scene(scene, cam)
    pc_host = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    flags = api.FRAME_STATUS | (api.FRAME_TWO_PASS if two_pass else 0)
    summary = []
    for k, (eye, center) in enumerate(views):
        if k:
            cam.look_at(eye, center)
            r.update_camera(pc_dev, cam)
        out = O.frame(pc_host, tg, two_pass=two_pass)
        st = r.frame(pc_dev, flags)
        assert st.visible_a == out["visibleA"].size
        compare_frame(r, tg, out, two_pass, label=f"view {k}")
        # status bytes: same classification per draw
        n = pc_host.meshletDrawCount
        assert np.array_equal(r.read_status(n, 0), out["statusA"] & O.STATUS_MASK)
        if two_pass:
            assert np.array_equal(r.read_status(n, 1), out["statusB"] & O.STATUS_MASK)
        summary.append((st.visible_a, st.occluded_a, st.visible_b, int(tg.tie.sum())))
    r.close()
    return summary


def orbit(scene, n, W, H):
    return [scene.default_view(i, 24) for i in range(n)]


def test_icosphere_cfg1():
    """BASELINE config 1: 64,980-triangle icosphere, 640x480"""
    s = Scene.icosphere(57)
    summ = run_views(s, 640, 480, [((0, 0, 3), (0, 0, 0)), ((0.05, 0.02, 3), (0, 0, 0)), ((0.4, 0.3, 2.9), (0, 0, 0))])
    assert summ[1][1] > 0  # second frame occlusion-culls the back of the sphere


def test_icosphere_two_pass():
    s = Scene.icosphere(40)
    summ = run_views(s, 640, 480, [((0, 0, 3), (0, 0, 0)), ((1.5, 0.5, 2.5), (0, 0, 0)), ((2.9, 0.2, 0.5), (0, 0, 0))], two_pass=True)
    assert any(v[2] > 0 for v in summ[1:])  # camera moved: pass B recovers disoccluded meshlets


@pytest.mark.parametrize("res", [(640, 480), (1920, 1080), (333, 217)])
def test_single_triangle(res):
    run_views(S.single_triangle(), res[0], res[1], [((0, 0, 3), (0, 0, 0)), ((0.3, 0.1, 2.0), (0, 0, 0))])


def test_fullscreen_quad_guard_band():
    run_views(S.fullscreen_quad(), 640, 480, [((0, 0, 3), (0, 0, 0)), ((0, 0, 3), (0.5, 0.2, 0))])


def test_ground_plane_near_clip():
    run_views(S.ground_plane(), 800, 450, [((0, 0, 3), (0, -0.2, 0)), ((1, 0.5, 2), (0, -0.5, -3)), ((0, 2, 0), (0.3, -1, 0.2))])


def test_coplanar_ties():
    summ = run_views(S.coplanar_overlap(), 320, 240, [((0, 0, 3), (0, 0, 0)), ((0.2, 0, 3), (0, 0, 0))])
    assert summ[0][3] > 1000  # tie pixels exist and were compared through ids_min


def test_occlusion_wall():
    summ = run_views(S.occluder_and_hidden(), 640, 480, [((0, 0, 8), (0, 0, 0)), ((0, 0, 8), (0, 0, 0)), ((0.5, 0, 8), (0, 0, 0))], two_pass=True)
    assert summ[1][1] >= 6  # the hidden cards are rejected by HiZ on the second frame


def test_mirrored_single_sided():
    run_views(S.mirrored_instances(), 640, 480, [((0, 0, 4), (0, 0, 0)), ((2, 1, 3), (0, 0, 0)), ((0, 0, -4), (0, 0, 0))])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_soup(seed):
    run_views(S.random_soup(seed=seed), 512, 384, [((0, 0, 3), (0, 0, 0)), ((0.5, 0.2, 2), (0, 0, -2)), ((-1, 0, -3), (0, 0, -6))], two_pass=True)


@pytest.mark.parametrize("seed", [11, 12])
def test_cull_stress_degenerate_boxes(seed):
    """status byte of every draw, both passes, on boxes that are zero-sized, huge, non-finite or cross the camera plane"""
    s = S.cull_stress(seed)
    assert s.counts().draws % 2 == 1  # odd draw count: the two-draws-per-thread split has a one-box tail
    views = [((0, 0, 6), (0, 0, 0)), ((0, 0, 6), (0, 0, 0)), ((0.7, 0.3, 5), (0, 0, -2)), ((-3, 1, -2), (0, 0, -6)), ((0, 0, 0), (0, 0, -1))]
    summ = run_views(s, 800, 600, views, two_pass=True)
    assert any(x[1] > 0 for x in summ) and any(x[2] > 0 for x in summ)  # the HiZ test rejected something and pass B recovered something


def test_resize_between_frames():
    """vkv_resize (application.cpp:578-602 updateRenderResolution): targets are rebuilt, the pyramid restarts from its cleared
    state, and the next frames are bit-exact at the new size — even (fused clear) and odd (separate clear, tail-only HiZ) extents"""
    s = S.occluder_and_hidden()
    r = api.Renderer(640, 480)
    cam = Camera(640, 480).look_at((0, 0, 8), (0, 0, 0))
    pc_dev = r.upload_scene(s, cam)
    for (W, H) in ((640, 480), (801, 451), (1280, 720), (127, 63)):
        r.resize(W, H)
        cam = Camera(W, H).look_at((0, 0, 8), (0, 0, 0))
        r.update_camera(pc_dev, cam)
        pc_host = s.host_push_constants(cam)
        tg = O.Targets(W, H)
        for eye in ((0, 0, 8), (0.4, 0.1, 8)):
            cam.look_at(eye, (0, 0, 0))
            r.update_camera(pc_dev, cam)
            out = O.frame(pc_host, tg, two_pass=True)
            r.frame(pc_dev, api.FRAME_TWO_PASS)
            compare_frame(r, tg, out, True, label=f"{W}x{H}")
    r.close()


def test_atrium_cfg2_small():
    """BASELINE config 2 geometry (int16-quantised atrium) at reduced detail; full size in test_atrium_cfg2_full"""
    s = Scene.atrium(32)
    run_views(s, 960, 540, orbit(s, 3, 960, 540), two_pass=True)


def test_lattice_small():
    s = Scene.lattice(4, 3, 4, 48)
    run_views(s, 1280, 720, orbit(s, 3, 1280, 720), two_pass=True)


def test_city_small():
    s = Scene.city(6, 5, 2000)
    run_views(s, 1024, 576, [s.default_view(i, 8) for i in range(3)], two_pass=True)


def test_atrium_cfg2_full():
    """BASELINE config 2 at full size: 262,144 triangles, 1920x1080"""
    s = Scene.atrium(128)
    assert s.counts().triangles_instanced == 262144
    run_views(s, 1920, 1080, orbit(s, 2, 1920, 1080), two_pass=True)


def test_city_cfg4_full_size():
    """BASELINE config 4 at full size: 50x40 unique buildings, 20.0 M triangles, 275 k MeshletDraws, 1920x1080, three views of
    the 64-view sweep, two-pass — bit-exact against the oracle"""
    s = Scene.city(50, 40, 10000, 0x5EED0004)
    assert s.counts().triangles_instanced > 19_000_000
    summ = run_views(s, 1920, 1080, [s.default_view(i, 64) for i in (0, 1, 2)], two_pass=True)
    assert summ[2][0] > 10_000 and summ[2][1] > 10_000


def test_lattice_cfg5_full_size():
    """BASELINE config 5 at full size on ONE GPU: 22x22x21 instances (10,164) of the 100,352-triangle patch = 1.02 billion
    triangles, 10.8 M MeshletDraws with the reference's meshlet partition (of the 2^25 the 25-bit draw index allows), 7680x4320 (12 pyramid mips), two frames, two-pass —
    bit-exact against the oracle.  (The 8-GPU range-sharded form of the same scene is bench.py --config 5 --gpus 8; its merge is
    checked in tests/test_multigpu.py.)"""
    s = Scene.lattice(22, 22, 21, 224, 0x5EED0003)
    assert s.counts().triangles_instanced > 1_000_000_000 and s.counts().draws < (1 << 25)
    summ = run_views(s, 7680, 4320, [s.default_view(i, 64) for i in (0, 1)], two_pass=True)
    assert summ[1][1] > 0.9 * s.counts().draws  # the second frame is occlusion-culled hard


def test_lattice_cfg3_full_size():
    """BASELINE config 3 at full size — 10x10x10 instances of a 100,352-triangle patch (1.06 M MeshletDraws, 100.35 M triangles),
    3840x2160, two-pass: bit-exact against the oracle (the multithreaded CPU port needs about half a second per frame), plus the
    size-independent properties of the path: disjoint pass lists, pass B drawn from pass A's rejects, every visbuffer id owned by
    a surviving draw, and bit-identical results when the sequence is rendered again in a fresh context."""
    s = Scene.lattice(10, 10, 10, 224, 0x5EED0003)
    assert s.counts().triangles_instanced == 100352000
    W, H = 3840, 2160
    views = [s.default_view(i, 64) for i in (0, 1)]
    summ = run_views(s, W, H, views, two_pass=True)
    assert summ[1][1] > 0.8 * s.counts().draws  # the second frame is occlusion-culled hard
    # determinism: the same sequence in a second context gives the same bits (64-bit atomicMin is order independent, the lists are sets)
    finals = []
    for rep in range(2):
        cam = Camera(W, H).look_at(*views[0])
        r = api.Renderer(W, H)
        pc = r.upload_scene(s, cam)
        for k in range(2):
            cam.look_at(*views[k])
            r.update_camera(pc, cam)
            r.frame(pc, api.FRAME_TWO_PASS | api.FRAME_STATUS)
        finals.append((r.read_visbuffer64(), r.read_pyramid().copy(), np.sort(r.read_visible(0)), np.sort(r.read_visible(1))))
        if rep == 0:
            r.close()
    for x, y in zip(finals[0], finals[1]):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y)
    a, b = r.read_visible(0), r.read_visible(1)
    n = s.counts().draws
    stA = r.read_status(n, 0)
    assert np.intersect1d(a, b).size == 0 and np.unique(a).size == a.size and np.unique(b).size == b.size
    assert (stA[b] == O.OCCLUDED).all() and (stA[a] == O.VISIBLE).all()
    ids = r.read_ids()
    drawn = np.unique(ids[ids != abi.VISBUFFER_CLEAR] >> 7)
    assert np.isin(drawn, np.concatenate([a, b])).all()
    r.close()


@pytest.mark.parametrize("res", [(640, 480), (1920, 1080), (3840, 2160), (1000, 1000), (1366, 768), (255, 257)])
def test_hiz_only(res):
    """HiZ kernels alone on a random depth image written through the visbuffer (all mips bit-exact, incl. odd sizes)"""
    W, H = res
    s = S.random_soup(200, seed=11, spread=4.0, size=2.5)
    run_views(s, W, H, [((0, 0, 3), (0, 0, 0))])


def test_no_cull_flag_matches_oracle_raster_of_all_draws():
    s = Scene.icosphere(12)
    W, H = 320, 240
    cam = S.camera(W, H)
    r = api.Renderer(W, H)
    pc_dev = r.upload_scene(s, cam)
    pc_host = s.host_push_constants(cam)
    tg = O.Targets(W, H)
    O.raster(pc_host, tg, np.arange(pc_host.meshletDrawCount, dtype=np.uint32))
    r.frame(pc_dev, api.FRAME_NO_CULL)
    assert np.array_equal(r.read_visbuffer64(), tg.vis64())
    # explicit list entry point, reversed submission order: atomicMin result is order independent
    r.clear()
    r.raster_list(pc_dev, np.arange(pc_host.meshletDrawCount, dtype=np.uint32)[::-1].copy())
    assert np.array_equal(r.read_visbuffer64(), tg.vis64())
    r.close()


def test_errors():
    r = api.Renderer(64, 64)
    pc = abi.PushConstants()
    pc.meshletDrawCount = 10  # NULL buffers
    with pytest.raises(api.VkvError):
        r.frame(pc)
    pc.meshletDrawCount = (1 << 25) + 1
    with pytest.raises(api.VkvError) as e:
        r.frame(pc)
    assert e.value.code == -5 or e.value.code == -2
    with pytest.raises(api.VkvError):
        r.read_status(4, 0)
    r.close()


def _max_channel_diff(a, b):
    d = np.abs(a.view(np.uint8).astype(np.int16) - b.view(np.uint8).astype(np.int16))
    return int(d.max()) if d.size else 0


@pytest.mark.parametrize("res", [(640, 480), (333, 217)])
def test_resolve_matches_oracle(res):
    """visbuffer -> RGBA8.  The id -> material chain is integer work (exact); the sRGB transfer function goes
    through powf on both sides (CUDA vs libm): tolerance +-1 code per 8-bit channel, stated here and in oracle.h."""
    W, H = res
    s = S.random_soup(300, seed=5, spread=3.0, size=1.0, materials=7)
    cam = S.camera(W, H)
    r = api.Renderer(W, H)
    pc_dev = r.upload_scene(s, cam)
    pc_host = s.host_push_constants(cam)
    tg = O.Targets(W, H)
    O.frame(pc_host, tg)
    r.frame(pc_dev)
    r.resolve(pc_dev)
    got = r.read_color()
    want = O.resolve(pc_host, tg)
    assert _max_channel_diff(got, want) <= 1
    # structure is exact: drawn <-> non-zero alpha, cleared / uncovered columns stay 0
    covered = (W // 32) * 32
    assert np.array_equal(got[:, covered:], np.zeros((H, W - covered), np.uint32))
    drawn = tg.ids_min[:, :covered] != 0xFFFFFFFF
    assert np.array_equal((got[:, :covered] >> 24) != 0, drawn)
    assert len(np.unique(got)) >= 4  # several materials actually show up
    r.close()


def _segments_of(draws):
    """run-length encode a host-built MeshletDraw[] into (primitiveIndex, transformIndex) segments (world.cpp:252-262 emits
    exactly one run per mesh-node x primitive)"""
    key = draws["primitiveIndex"].astype(np.uint64) << np.uint64(32) | draws["transformIndex"].astype(np.uint64)
    start = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    return np.stack([draws["primitiveIndex"][start], draws["transformIndex"][start]], axis=1).astype(np.uint32)


@pytest.mark.parametrize("make", [lambda: Scene.lattice(5, 4, 3, 48), lambda: Scene.city(6, 5, 2000), lambda: Scene.atrium(32), lambda: Scene.icosphere(20)])
def test_device_draw_list(make):
    """vkv_build_draws expands segments into a MeshletDraw[] byte-identical to World::rebuildDrawBuffer's"""
    s = make()
    W, H = 640, 360
    cam = s.default_camera(W, H)
    r = api.Renderer(W, H)
    pc = r.upload_scene(s, cam)
    host = s.draws()
    seg = _segments_of(host)
    addr, n = r.build_draws(seg, pc.primitiveBuffer)
    assert n == host.shape[0]
    got = r.download(addr, n * 12).view(abi.DRAW_DTYPE)
    assert np.array_equal(got, host)
    # and a frame through the device-built list is the frame through the uploaded one
    st0 = r.frame(pc, api.FRAME_TWO_PASS)
    a = r.read_visbuffer64()
    pc2 = abi.PushConstants.from_buffer_copy(bytes(pc))
    pc2.drawBuffer = addr
    st1 = r.frame(pc2, api.FRAME_TWO_PASS)  # culls against the pyramid of the previous frame: same camera, first frame all-far
    r.frame(pc, api.FRAME_TWO_PASS)
    b = r.read_visbuffer64()
    r.frame(pc2, api.FRAME_TWO_PASS)
    assert np.array_equal(r.read_visbuffer64(), b) and st0.draws == st1.draws
    assert a.shape == b.shape
    r.close()


def test_device_draw_list_limits():
    s = Scene.lattice(2, 2, 2, 24)
    r = api.Renderer(64, 64)
    pc = r.upload_scene(s, s.default_camera(64, 64))
    host = s.draws()
    seg = _segments_of(host)
    per = host.shape[0] // seg.shape[0]
    reps = (1 << 25) // per + 2
    big = np.tile(seg, (reps, 1))
    with pytest.raises(api.VkvError) as e:
        r.build_draws(big, pc.primitiveBuffer)
    assert e.value.code == -5
    addr, n = r.build_draws(np.zeros((0, 2), np.uint32), pc.primitiveBuffer)
    assert n == 0
    r.close()


def test_shared_reciprocal_division():
    """the kernels' x/w, y/w, z/w with one refined reciprocal (common.cuh div3_shared) is bit-identical to IEEE `/`
    (what the oracle and the reference's SPIR-V OpFDiv compute) over the whole operand range it is used for"""
    r = api.Renderer(64, 64)
    total = 0
    for seed in (1, 0xC0FFEE, 0x5EED0003):
        tested, bad = r.selftest_division(seed, 2048)
        assert bad == 0, f"{bad} of {tested} quotients differ from IEEE division"
        total += tested
    assert total > 3_000_000_000
    r.close()


@pytest.mark.parametrize("caps", [("3", "2"), ("1", "100000"), ("100000", "1")])
def test_raster_queue_overflow_falls_back_to_the_rewalk(monkeypatch, caps):
    """raster_kernel queues what a lane cannot finish alone (triangles for the clipper, triangles above the lane-serial limit) for
    raster_big_kernel; with tiny queues (VKV_BIG_CAP / VKV_CLIP_CAP, read at vkv_create) both overflow on these scenes and the drain
    kernel's re-walk of the meshlet list must produce the same bits: near-plane clipping + guard band (ground plane under the
    camera), two huge triangles (full-screen quad), an interior scene (atrium) and ordinary small triangles (lattice)"""
    monkeypatch.setenv("VKV_BIG_CAP", caps[0])
    monkeypatch.setenv("VKV_CLIP_CAP", caps[1])
    run_views(S.ground_plane(24, 40.0), 640, 480, [((0, 0, 3), (0, -0.2, 0)), ((0, 2, 0), (0.3, -1, 0.2))], two_pass=True)
    run_views(S.fullscreen_quad(), 800, 600, [((0, 0, 3), (0, 0, 0))])
    s = Scene.atrium(16)
    run_views(s, 960, 540, orbit(s, 2, 960, 540), two_pass=True)
    s = Scene.lattice(3, 2, 2, 24)
    run_views(s, 640, 360, orbit(s, 2, 640, 360), two_pass=True)


@pytest.mark.parametrize("make,views", [
    (lambda: Scene.icosphere(40), [((0, 0, 3), (0, 0, 0)), ((1.2, 0.5, 2.6), (0, 0, 0)), ((-2.0, -1.0, -2.0), (0, 0, 0))]),
    (lambda: S.mirrored_instances(), [((0, 0, 4), (0, 0, 0)), ((2, 1, 3), (0, 0, 0)), ((0, 0, -4), (0, 0, 0))]),
    (lambda: Scene.city(6, 5, 2000), None),
])
def test_cone_cull_matches_the_oracle_and_leaves_the_image_alone(make, views):
    """optional normal-cone cull (VKV_FRAME_CONE_CULL; north_star's second cull test, which the reference disables at assets.cpp:323):
    per-draw status bytes — VKV_ST_CONE_CULLED included — equal the oracle's with its cone stage on; the 64-bit visbuffer and the
    pyramid equal BOTH the oracle's with the stage on and the CUDA frame with the stage off; the stage does reject meshlets."""
    scene = make()
    W, H = 800, 600
    if views is None:
        views = [scene.default_view(i, 16) for i in range(3)]
    cam = Camera(W, H).look_at(*views[0])
    r_on, r_off = api.Renderer(W, H), api.Renderer(W, H)
    pc_on, pc_off = r_on.upload_scene(scene, cam), r_off.upload_scene(scene, cam)
    r_on.upload_cones(scene)
    pc_host = scene.host_push_constants(cam)
    cones = scene.host_cones()
    tg = O.Targets(W, H)
    n = pc_host.meshletDrawCount
    rejected = 0
    for k, v in enumerate(views):
        if k:
            cam.look_at(*v)
            r_on.update_camera(pc_on, cam); r_off.update_camera(pc_off, cam)
        out = O.frame(pc_host, tg, two_pass=True, cones=cones)
        st = r_on.frame(pc_on, api.FRAME_TWO_PASS | api.FRAME_STATUS | api.FRAME_CONE_CULL)
        r_off.frame(pc_off, api.FRAME_TWO_PASS)
        got = r_on.read_status(n, 0)
        assert np.array_equal(got, out["statusA"] & (O.STATUS_MASK | O.CONE_CULLED)), f"view {k}: status bytes differ in {(got != (out['statusA'] & (O.STATUS_MASK | O.CONE_CULLED))).sum()} draws"
        rejected += int((got == api.ST_CONE_CULLED).sum())
        assert np.array_equal(np.sort(r_on.read_visible(0)), out["visibleA"]) and np.array_equal(np.sort(r_on.read_visible(1)), out["visibleB"])
        vis = r_on.read_visbuffer64()
        assert np.array_equal(vis, tg.vis64()) and np.array_equal(vis, r_off.read_visbuffer64()), f"view {k}: the cone stage changed the image"
        assert np.array_equal(r_on.read_pyramid().view(np.uint32), r_off.read_pyramid().view(np.uint32))
        assert st.visible_a <= r_off.read_visible(0).size
    assert rejected > 0
    r_on.close(); r_off.close()


def test_cone_cull_needs_its_table():
    s = Scene.icosphere(8)
    cam = Camera(320, 240).look_at((0, 0, 3), (0, 0, 0))
    r = api.Renderer(320, 240)
    pc = r.upload_scene(s, cam)
    with pytest.raises(api.VkvError) as e:
        r.frame(pc, api.FRAME_CONE_CULL)
    assert e.value.code == -2
    r.close()


def _normalized_i16_scene():
    """KHR_mesh_quantization with a NORMALIZED SHORT accessor (max(x / 32767, -1)), -32768 included, under a scaled node"""
    rng = np.random.default_rng(4)
    pos, idx = S.grid_mesh(40, 40, lambda u, v: (u * 2 - 1, v * 2 - 1, 0.3 * np.sin(u * 8) * np.cos(v * 6)))
    q = np.clip(np.rint(pos * 32767), -32768, 32767).astype(np.int16)
    q[0] = (-32768, -32768, 0)
    s = Scene.new()
    m = s.add_material(double_sided=True)
    p = s.add_primitive_i16(q, idx, m, normalized=True)
    for _ in range(5):
        s.add_node(p, translation=rng.uniform(-2, 2, 3), scale=rng.uniform(0.5, 2.5, 3))
    return s.finalize()


@pytest.mark.parametrize("make", [lambda: Scene.atrium(24), lambda: Scene.city_quantized(6, 5, 2000), _normalized_i16_scene])
def test_int16_positions_dequantised_in_registers_give_the_same_bits(make):
    """SURVEY D4 / north_star: with vkv_set_quantized_positions the rasteriser reads the accessor's own int16 data (8 B per vertex) and
    applies fastgltf's convertComponent in registers; the reference expands on the host (assets.cpp:310-314).  Same floats, hence the
    same 64-bit visbuffer and pyramid as the f32 Vertex path AND as the oracle, which only ever sees the expanded f32 records."""
    scene = make()
    W, H = 960, 540
    views = [scene.default_view(i, 12) for i in range(3)] if scene.counts().primitives > 1 else [((0, 0, 6), (0, 0, 0)), ((2, 1, 5), (0, 0, 0))]
    cam = Camera(W, H).look_at(*views[0])
    r_q, r_f = api.Renderer(W, H), api.Renderer(W, H)
    pc_q, pc_f = r_q.upload_scene(scene, cam), r_f.upload_scene(scene, cam)
    assert r_q.upload_quantized(scene) != 0
    pc_host = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    for k, v in enumerate(views):
        if k:
            cam.look_at(*v)
            r_q.update_camera(pc_q, cam); r_f.update_camera(pc_f, cam)
        out = O.frame(pc_host, tg, two_pass=True)
        r_q.frame(pc_q, api.FRAME_TWO_PASS); r_f.frame(pc_f, api.FRAME_TWO_PASS)
        compare_frame(r_q, tg, out, True, label=f"int16 view {k}")
        assert np.array_equal(r_q.read_visbuffer64(), r_f.read_visbuffer64())
    # removing the table restores the f32 path
    r_q.set_quantized_positions(0)
    r_q.frame(pc_q, api.FRAME_TWO_PASS); r_f.frame(pc_f, api.FRAME_TWO_PASS)
    assert np.array_equal(r_q.read_visbuffer64(), r_f.read_visbuffer64())
    r_q.close(); r_f.close()


def test_glb_asset_renders_bit_exactly():
    """a procedurally generated glTF 2.0 binary — float and KHR_mesh_quantization primitives in one mesh, instancing, TRS and matrix
    nodes, a double-sided material — through host/gltf.cpp, the C ABI and the CUDA path: bit-exact against the oracle, with the f32
    Vertex records and with the SHORT accessor read directly by the rasteriser"""
    from tests.gltf_writer import GlbWriter
    rng = np.random.default_rng(17)
    w = GlbWriter()
    m0, m1 = w.material((0.9, 0.4, 0.1, 1), False), w.material((0.2, 0.6, 0.9, 1), True)
    posA, idxA = S.grid_mesh(40, 30, lambda u, v: (u * 4 - 2, 0.3 * np.sin(u * 9) * np.cos(v * 7), v * 3 - 1.5))
    posB, idxB = S.icosphere_soup(9)
    qB = np.rint(posB * 4000).astype(np.int16)
    mesh = w.mesh([{"position": w.positions(posA), "indices": w.indices(idxA.astype(np.uint32)), "material": m1},
                   {"position": w.positions(qB), "indices": w.indices(idxB.astype(np.uint16)), "material": m0}])
    root = w.node(translation=(0, 0, -1))
    for k in range(6):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w.node(mesh, parent=root, translation=rng.uniform(-3, 3, 3), rotation=q, scale=(1, 1, 1) if k % 2 else (1.5, 0.7, -1.2))
    M = np.eye(4, dtype=np.float32); M[:3, :3] *= 1 / 4000.0 * 2; M[:3, 3] = (0, 2, 0)
    w.node(w.mesh([{"position": w.positions(qB), "indices": w.indices(idxB.astype(np.uint16)), "material": m0}]), matrix=M.T)
    scene = Scene.from_glb(w.glb())
    views = [((0, 1, 9), (0, 0, 0)), ((4, 2, 7), (0, 0, 0)), ((-5, 0.5, -6), (0, 0, 0))]
    run_views(scene, 800, 600, views, two_pass=True)
    W, H = 800, 600
    cam = Camera(W, H).look_at(*views[0])
    r = api.Renderer(W, H)
    pc = r.upload_scene(scene, cam)
    r.upload_quantized(scene)
    pc_host = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    out = O.frame(pc_host, tg, two_pass=True)
    r.frame(pc, api.FRAME_TWO_PASS)
    compare_frame(r, tg, out, True, label="glb + int16")
    r.close()


def test_partial_pass_b_pyramid_equals_the_full_rebuild(monkeypatch):
    """vkv_frame's second pyramid build redoes only the 64x16-pixel tiles the pass-B rasteriser marked (small passes; a large pass B
    switches the marks off on the device and every tile is redone).  VKV_HIZ_FULL_B=1 rebuilds everything: same bits, frame after frame."""
    s = Scene.lattice(4, 3, 4, 48)
    W, H = 1280, 720
    views = [s.default_view(i, 24) for i in range(5)]
    results = []
    for full in ("0", "1"):
        monkeypatch.setenv("VKV_HIZ_FULL_B", full)
        cam = Camera(W, H)
        cam.look_at(*views[0])
        r = api.Renderer(W, H)
        pc = r.upload_scene(s, cam)
        frames = []
        for v in views:
            cam.look_at(*v)
            r.update_camera(pc, cam)
            st = r.frame(pc, api.FRAME_TWO_PASS)
            frames.append((st.visible_a, st.occluded_a, st.visible_b, r.hash(0), r.hash(1)))
        results.append((frames, r.read_pyramid().view(np.uint32).copy()))
        r.close()
    assert results[0][0] == results[1][0]
    assert np.array_equal(results[0][1], results[1][1])
    assert any(f[2] > 0 for f in results[0][0][1:])  # the camera moves: pass B draws something, so there are dirty tiles to redo


def test_drain_grid_follows_the_observed_queues():
    """The drain launch behind the rasteriser is sized from what the last observed frames found queued (an empty drain is pure launch
    latency): after two frames with empty clip / large-triangle queues it shrinks to half a block per SM.  The first frame that fills
    the queues again still runs on the small grid and must produce the same bits; the grid grows back right after."""
    s = S.ground_plane(24, 40.0)
    up = ((0, 0, 3), (0, 10, 2))           # looking steeply up: two meshlets pass the box test, every triangle is rejected, nothing is queued
    down = ((0, 0, 3), (0, -0.2, 0))       # near-plane clipping + screen-filling triangles: both queues in use
    low = ((0, 2, 0), (0.3, -1, 0.2))
    views = [up, up, up, down, low, up, up, up, low, down]
    summ = run_views(s, 640, 480, views, two_pass=True)   # bit-exact against the oracle, frame by frame
    assert summ[3][0] > summ[0][0] and summ[8][0] > 0
    # the same sweep again, looking at what the drain kernels report
    W, H = 640, 480
    cam = Camera(W, H)
    cam.look_at(*views[0])
    r = api.Renderer(W, H)
    pc = r.upload_scene(s, cam)
    items = []
    for v in views:
        cam.look_at(*v)
        r.update_camera(pc, cam)
        st = r.frame(pc, api.FRAME_TWO_PASS)
        items.append(st.drain_items_a)
    r.close()
    assert items[:3] == [0, 0, 0] and items[5:8] == [0, 0, 0]          # two idle frames observed -> frames 3 and 8 run on the small grid
    assert items[3] > 0 and items[4] > 0 and items[8] > 0 and items[9] > 0

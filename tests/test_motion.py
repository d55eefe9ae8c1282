"""Motion vectors — the visbuffer pass's second colour attachment (visbuffer.frag.glsl:38, R16G16_SFLOAT), derived from the finished
visbuffer (csrc/motion.cu, vkv_motion_vectors).

CPU:  * the oracle (orc_motion_vectors) against llvmpipe running the REFERENCE's line 38 verbatim in its fragment shader, with the
        position / prevPosition varyings interpolated by llvmpipe's own perspective-correct interpolator;
      * the product's arithmetic (csrc/motion_core.h, the source the CUDA kernel compiles) built for the host, against the oracle: same bits;
      * the fp16 store conversion of both against numpy's.
GPU:  the CUDA pass against the oracle, bit for bit.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from vk_gltf_viewer_b200 import abi
from vk_gltf_viewer_b200.scene import Camera, Scene

from . import llvmpipe_cases as K
from . import llvmpipe_lib as LP
from . import oracle_lib as O
from . import scenes as S
from .conftest import ROOT

# (scene, camera positions of two consecutive frames, resolution): the camera moves between the frames, so prevViewProjection != viewProjection
CASES = {
    "icosphere": (lambda: Scene.icosphere(24), [((0, 0, 3), (0, 0, 0)), ((0.25, 0.1, 2.9), (0.02, 0, 0))], (320, 240)),
    "lattice": (lambda: Scene.lattice(3, 3, 3, 40, 0x5EED0003), None, (480, 270)),
    "ground_clipped": (lambda: S.ground_plane(8, 30.0, -1.0), [((0, 0, 0), (0, 0, -1)), ((0.05, 0.02, -0.1), (0.1, -0.05, -1))], (400, 300)),
    "mirrored": (lambda: S.mirrored_instances(), [((0, 0, 4), (0, 0, 0)), ((0.3, 0.2, 3.8), (0, 0, 0))], (320, 240)),
    "atrium": (lambda: Scene.atrium(16), None, (480, 270)),
}


def two_frames(name):
    make, views, (W, H) = CASES[name]
    scene = make()
    cam = Camera(W, H)
    v0, v1 = views if views else (scene.default_view(3, 64), scene.default_view(4, 64))
    cam.look_at(*v0)
    cam.look_at(*v1)            # Camera::updateCamera: prevViewProjection = the first frame's viewProjection (camera.cpp:181)
    assert cam.matrix("prevViewProjection").tobytes() != cam.matrix("viewProjection").tobytes()
    return scene, cam, W, H


def shim():
    p = os.path.join(ROOT, "oracle", "_ref", "libmotion_core_shim.so")
    src = os.path.join(ROOT, "tests", "motion_core_shim.cpp")
    core = os.path.join(ROOT, "vk_gltf_viewer_b200", "csrc", "motion_core.h")
    if not os.path.exists(p) or os.path.getmtime(p) < max(os.path.getmtime(src), os.path.getmtime(core)):
        os.makedirs(os.path.dirname(p), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", p, src,
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "vk_gltf_viewer_b200", "csrc")])
    L = C.CDLL(p)
    L.shim_motion_vectors.argtypes = [C.POINTER(abi.PushConstants), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.shim_half_rn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    return L


def test_half_conversion_of_core_and_oracle_is_round_to_nearest_even():
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2**32, 400000, dtype=np.uint64).astype(np.uint32)
    special = np.array([0, 0x80000000, 0x7f800000, 0xff800000, 0x7fc00000, 0x33000000, 0x33000001, 0x33800000, 0x387fc000, 0x387fe000, 0x38800000,
                        0x477fe000, 0x477ff000, 0x477fefff, 0x47800000, 0x3f801000, 0x3f803000, 0x3f801001, 0x00000001, 0x007fffff], np.uint32)
    halves = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float32)       # every finite half, and the midpoints between neighbours
    mids = ((halves[:-1].astype(np.float64) + halves[1:].astype(np.float64)) / 2).astype(np.float32)
    f = np.concatenate([bits.view(np.float32), special.view(np.float32), halves, -halves, mids, -mids,
                        np.nextafter(mids, np.float32(np.inf)), np.nextafter(mids, np.float32(-np.inf)), rng.normal(0, 0.01, 100000).astype(np.float32)])
    with np.errstate(over="ignore", invalid="ignore"):
        want = f.astype(np.float16).view(np.uint16).copy()
    nan = np.isnan(f)
    want[nan] = 0x7fff
    got = np.zeros(f.size, np.uint16)
    shim().shim_half_rn(np.ascontiguousarray(f).ctypes.data, got.ctypes.data, f.size)
    assert np.array_equal(got, want)
    L = O.lib()
    sub = np.concatenate([np.arange(0, f.size, 37), np.arange(bits.size, bits.size + special.size)])
    assert all(L.orc_to_half(float(f[i])) == want[i] for i in sub if not nan[i]) and L.orc_to_half(float("nan")) == 0x7fff


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_build_of_the_cuda_arithmetic_equals_the_oracle(name):
    scene, cam, W, H = two_frames(name)
    pc = scene.host_push_constants(cam)
    tg = K.oracle_images(pc, W, H)
    f, h = O.motion_vectors(pc, tg)
    gf, gh = np.zeros_like(f), np.zeros_like(h)
    assert shim().shim_motion_vectors(C.byref(pc), W, H, np.ascontiguousarray(tg.ids_min).ctypes.data, gf.ctypes.data, gh.ctypes.data) == 0
    assert np.array_equal(gf.view(np.uint32), f.view(np.uint32)) and np.array_equal(gh, h)
    cov = tg.ids_min != abi.VISBUFFER_CLEAR
    assert cov.sum() > 1000 and (h[~cov] == 0).all() and np.abs(f[cov]).max() > 1e-3, "the camera moved: covered pixels must carry motion"


@pytest.mark.skipif(not LP.available(), reason="no Mesa xlib libGL (llvmpipe) in this image")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_motion_vectors_match_llvmpipe_running_the_reference_line(name):
    """llvmpipe rasterises the same triangles with the reference's fragment-shader line and its own perspective-correct varyings; wherever it
    shows the same triangle as the oracle, the two motion vectors agree to fp32 noise (and to the fp16 attachment's resolution)"""
    scene, cam, W, H = two_frames(name)
    pc = scene.host_push_constants(cam)
    V, I = K.oracle_triangles(scene, pc)
    # prevPosition = prevViewProjection * transform * vertex (mesh.glsl:45,63): the same vertices, the previous matrix
    cam_prev = Camera(W, H)
    C.memmove(C.byref(cam_prev.c), C.byref(cam.c), C.sizeof(cam.c))
    cam_prev.c.viewProjection = cam.c.prevViewProjection
    Vp, Ip = K.oracle_triangles(scene, scene.host_push_constants(cam_prev), facing_from=pc)
    assert np.array_equal(I, Ip)
    tg = K.oracle_images(pc, W, H)
    f, h = O.motion_vectors(pc, tg, tg.ids_ref)
    lids, lmv = LP.instance().raster_motion(W, H, V, Vp, I)
    same = (tg.ids_ref != abi.VISBUFFER_CLEAR) & (lids >= 0)
    same &= tg.ids_ref == np.where(lids >= 0, lids, 0).astype(np.uint32)
    assert same.sum() > 0.99 * (tg.ids_ref != abi.VISBUFFER_CLEAR).sum()
    d = np.abs(f - lmv)[same]
    scale = np.abs(f[same]).max()
    print(f"\n{name}: {int(same.sum())} pixels, |motion| up to {scale:.3e}; oracle vs llvmpipe: median {np.median(d):.2e}, max {d.max():.2e}")
    assert d.max() < 2e-5 + 1e-4 * scale and np.median(d) < 1e-6
    # at the attachment's precision (the values are differences of two numbers near 0.5, so fp32 noise is a few per cent of a half's spacing):
    # never more than one representable half apart — or, for components near zero where fp16 resolves finer than that fp32 noise, 1e-6
    lh, oh = lmv.astype(np.float16), h.view(np.float16)
    gap = np.abs(lh.astype(np.float32) - oh.astype(np.float32))[same]
    unit = np.maximum(np.spacing(np.abs(lh)), np.spacing(np.abs(oh))).astype(np.float32)[same]
    assert (gap <= np.maximum(unit, 1e-6)).all(), float((gap / unit).max())
    print(f"   as R16G16_SFLOAT: {100 * (gap == 0).mean():.1f} % of the halves identical, the rest one unit (or < 1e-6) apart")


@pytest.mark.gpu
@pytest.mark.late
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_motion_vectors_equal_the_oracle(name):
    from vk_gltf_viewer_b200 import api
    scene, cam, W, H = two_frames(name)
    pc_host = scene.host_push_constants(cam)
    tg = K.oracle_images(pc_host, W, H)
    _, want = O.motion_vectors(pc_host, tg)                       # ids_min = the low word of the 64-bit visbuffer
    r = api.Renderer(W, H)
    pc = r.upload_scene(scene, cam)
    r.frame(pc, api.FRAME_NO_CULL)
    assert np.array_equal(r.read_ids(), tg.ids_min)
    r.motion_vectors(pc)
    got = r.read_motion()
    r.close()
    assert np.array_equal(got, want), int((got != want).sum())


def test_a_camera_at_rest_has_no_motion():
    """prevViewProjection == viewProjection: both terms of line 38 are the same expression, the difference is exactly zero"""
    scene = Scene.icosphere(12)
    W, H = 200, 150
    cam = Camera(W, H).look_at((0, 0, 3), (0, 0, 0))
    cam.look_at((0, 0, 3), (0, 0, 0))
    pc = scene.host_push_constants(cam)
    tg = K.oracle_images(pc, W, H)
    f, h = O.motion_vectors(pc, tg)
    assert (f == 0).all() and (h & 0x7fff == 0).all() and (tg.ids_min != abi.VISBUFFER_CLEAR).sum() > 1000

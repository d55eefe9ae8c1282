"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from vk_gltf_viewer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FRUSTUM_CULLED, OCCLUDED, VISIBLE, NOT_TESTED, STATUS_MASK = 0, 1, 2, 3, 3
AMBIG_FRUSTUM, AMBIG_HIZ, AMBIG_LEVEL, CROSSES_CAMERA, AMBIG_FOOTPRINT = 4, 8, 16, 32, 64
CONE_CULLED = 128


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "tested", "frustum_culled", "occluded", "visible",
        "ambig_frustum", "ambig_hiz", "ambig_level", "ambig_footprint", "crosses_camera",
        "meshlets", "triangles_in", "triangles_culled_facing", "triangles_rejected", "triangles_clipped",
        "triangles_degenerate", "triangles_rasterised", "fragments", "fragments_passed", "tie_pixels")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ROOT, "oracle/liboracle.so"])
        L = C.CDLL(path)
        L.orc_pyramid_layout.restype = C.c_uint32
        L.orc_cull.argtypes = [C.POINTER(abi.PushConstants), C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.POINTER(Counters), C.c_int]
        L.orc_cull_cone.argtypes = L.orc_cull.argtypes + [C.c_void_p]
        L.orc_set_diagnostics.argtypes = [C.c_int]
        L.orc_raster.argtypes = [C.POINTER(abi.PushConstants), C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.POINTER(Counters), C.c_int]
        L.orc_clear.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hiz.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_resolve.argtypes = [C.POINTER(abi.PushConstants), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_sample_min.restype = C.c_float
        L.orc_sample_min.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_int)]
        L.orc_mesh_shader.restype = None
        L.orc_mesh_shader.argtypes = [C.POINTER(abi.PushConstants), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_motion_vectors.argtypes = [C.POINTER(abi.PushConstants), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_to_half.restype = C.c_uint16
        L.orc_to_half.argtypes = [C.c_float]
        L.orc_vis64_key.restype = C.c_uint64
        L.orc_vis64_key.argtypes = [C.c_float, C.c_uint32]
        _lib = L
    return _lib


def mesh_shader(pc, draw_ids):
    """visbuffer.mesh.glsl:43-102 for the given MeshletDraws -> clip [n,64,4] f32, cull [n,126] u8 (0 / 1, 0xff = no such triangle),
    det [n,126] f32, transformDet [n] f32, ambig [n,126] u8 (|det| <= noise), noise [n,126] f32 (first-order rounding-noise bound of det)"""
    ids = np.ascontiguousarray(draw_ids, np.uint32)
    n = ids.size
    clip = np.zeros((n, 64, 4), np.float32); cull = np.zeros((n, 126), np.uint8); det = np.zeros((n, 126), np.float32)
    tdet = np.zeros(n, np.float32); ambig = np.zeros((n, 126), np.uint8); noise = np.zeros((n, 126), np.float32)
    lib().orc_mesh_shader(C.byref(pc), ids.ctypes.data, n, clip.ctypes.data, cull.ctypes.data, det.ctypes.data, tdet.ctypes.data, ambig.ctypes.data,
                          noise.ctypes.data)
    return clip, cull, det, tdet, ambig, noise


class Targets:
    """depth / ids / tie images + pyramid for one view (what the reference's attachments hold)."""

    def __init__(self, W, H):
        self.W, self.H = W, H
        self.depth = np.zeros((H, W), np.float32)
        self.ids_ref = np.full((H, W), abi.VISBUFFER_CLEAR, np.uint32)
        self.ids_min = np.full((H, W), abi.VISBUFFER_CLEAR, np.uint32)
        self.tie = np.zeros((H, W), np.uint8)
        self.levels, self.layout, total = abi.pyramid_layout(W, H)
        self.pyramid = np.zeros(total, np.float32)  # context-create clear: 0.0 = far

    def clear(self):
        lib().orc_clear(self.W, self.H, self.depth.ctypes.data, self.ids_ref.ctypes.data, self.ids_min.ctypes.data, self.tie.ctypes.data)

    def mip(self, k):
        off, w, h = self.layout[k]
        return self.pyramid[off:off + w * h].reshape(h, w)

    def vis64(self):
        """the 64-bit visbuffer an atomicMin rasteriser must produce: (~bits(depth) << 32) | ids_min"""
        hi = (~self.depth.view(np.uint32)).astype(np.uint64)
        return (hi << np.uint64(32)) | self.ids_min.astype(np.uint64)


def cull(pc, W, H, pyramid, vp_select=0, only_status=None, threads=0, cones=None):
    """cones: host cone table (Scene.host_cones()) -> the optional normal-cone stage runs between the frustum and the HiZ test"""
    n = pc.meshletDrawCount
    status = np.zeros(n, np.uint8)
    ctr = Counters()
    rc = lib().orc_cull_cone(C.byref(pc), W, H, pyramid.ctypes.data, vp_select,
                             only_status.ctypes.data if only_status is not None else None, status.ctypes.data, C.byref(ctr), threads,
                             cones)
    assert rc == 0
    return status, ctr


def raster(pc, tg: Targets, draw_ids, threads=0):
    ids = np.ascontiguousarray(draw_ids, np.uint32)
    ctr = Counters()
    rc = lib().orc_raster(C.byref(pc), tg.W, tg.H, ids.ctypes.data, ids.shape[0], tg.depth.ctypes.data, tg.ids_ref.ctypes.data,
                          tg.ids_min.ctypes.data, tg.tie.ctypes.data, C.byref(ctr), threads)
    assert rc == 0
    return ctr


def hiz(tg: Targets, threads=0):
    rc = lib().orc_hiz(tg.W, tg.H, tg.depth.ctypes.data, tg.pyramid.ctypes.data, threads)
    assert rc == 0


def resolve(pc, tg: Targets, out=None):
    """visbuffer_resolve.comp.glsl on the min-id image (what the 64-bit visbuffer's low word holds)"""
    if out is None:
        out = np.zeros((tg.H, tg.W), np.uint32)
    rc = lib().orc_resolve(C.byref(pc), tg.W, tg.H, tg.ids_min.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def motion_vectors(pc, tg: Targets, ids=None):
    """visbuffer.frag.glsl:38 per pixel of the id image (default: the min-id image, the 64-bit visbuffer's low word)
    -> (float32 [H, W, 2] before the fp16 store, uint16 [H, W, 2] = the R16G16_SFLOAT attachment)"""
    ids = np.ascontiguousarray(tg.ids_min if ids is None else ids, np.uint32)
    f = np.zeros((tg.H, tg.W, 2), np.float32)
    h = np.zeros((tg.H, tg.W, 2), np.uint16)
    assert lib().orc_motion_vectors(C.byref(pc), tg.W, tg.H, ids.ctypes.data, f.ctypes.data, h.ctypes.data) == 0
    return f, h


def visible_ids(status):
    return np.nonzero((status & STATUS_MASK) == VISIBLE)[0].astype(np.uint32)


def frame(pc, tg: Targets, two_pass=False, threads=0, cones=None, stage_s=None):
    """One frame as the reference records it (cull with the previous pyramid -> raster -> HiZ rebuild), or the two-pass
    extension: A = reference pass; HiZ; B = re-test A's occlusion rejects with the current VP/pyramid; raster; HiZ.
    stage_s: optional dict accumulating wall seconds per stage (clear, cull_a, raster_a, hiz_a, cull_b, raster_b, hiz_b)."""
    import time
    t = [time.perf_counter()]

    def lap(name):
        if stage_s is not None:
            now = time.perf_counter()
            stage_s[name] = stage_s.get(name, 0.0) + (now - t[0])
            t[0] = now

    tg.clear(); lap("clear")
    stA, cA = cull(pc, tg.W, tg.H, tg.pyramid, 0, None, threads, cones)
    visA = visible_ids(stA); lap("cull_a")
    rA = raster(pc, tg, visA, threads); lap("raster_a")
    hiz(tg, threads); lap("hiz_a")
    out = {"statusA": stA, "visibleA": visA, "cullA": cA, "rasterA": rA}
    if two_pass:
        stB, cB = cull(pc, tg.W, tg.H, tg.pyramid, 1, stA, threads)
        visB = visible_ids(stB); lap("cull_b")
        rB = raster(pc, tg, visB, threads); lap("raster_b")
        hiz(tg, threads); lap("hiz_b")
        out.update({"statusB": stB, "visibleB": visB, "cullB": cB, "rasterB": rB})
    return out

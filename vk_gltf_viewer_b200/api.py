"""Python face of libvkv.so (include/vkv.h) — the drop-in for the reference's visbuffer + HiZ passes.

Names follow the reference: a `Renderer` owns what `VisibilityBufferPass` + `HiZReductionPass` own in
application.hpp:41-83 (visbuffer, depth, pyramid), `frame()` is the region of Application::run() the library replaces
(application.cpp:763-867, 951-1003).  Every call goes through the C ABI; there is no Python/CPU fallback.
"""
import ctypes as C

import numpy as np

from . import abi
from ._native import vkv_lib

FRAME_ONE_PASS = 0
FRAME_TWO_PASS = 1 << 0
FRAME_NO_HIZ = 1 << 1
FRAME_STATUS = 1 << 2
FRAME_TIMED = 1 << 3
FRAME_STAGES = 1 << 6
FRAME_NO_CULL = 1 << 4
FRAME_MERGE = 1 << 5
FRAME_MERGE_STRIPS = 1 << 7
FRAME_CONE_CULL = 1 << 8

ST_FRUSTUM_CULLED, ST_OCCLUDED, ST_VISIBLE, ST_NOT_TESTED = 0, 1, 2, 3
ST_CONE_CULLED = 0x80


class VkvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vkv error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [
        ("draws", C.c_uint32), ("visible_a", C.c_uint32), ("occluded_a", C.c_uint32), ("visible_b", C.c_uint32), ("tested_b", C.c_uint32),
        ("clear_ms", C.c_float), ("cull_a_ms", C.c_float), ("raster_a_ms", C.c_float), ("hiz_a_ms", C.c_float),
        ("cull_b_ms", C.c_float), ("raster_b_ms", C.c_float), ("hiz_b_ms", C.c_float), ("total_ms", C.c_float),
        ("kernel_launches", C.c_uint32), ("merge_a_ms", C.c_float), ("merge_b_ms", C.c_float),
        ("strip_tiles_pulled", C.c_uint32), ("strip_texels_sent", C.c_uint32), ("hiz_tiles_b", C.c_uint32),
        ("drain_items_a", C.c_uint32), ("drain_items_b", C.c_uint32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = [
    "vkv_create", "vkv_resize", "vkv_destroy", "vkv_last_error", "vkv_set_stream", "vkv_sync",
    "vkv_upload", "vkv_update", "vkv_free",
    "vkv_motion_vectors", "vkv_read_motion", "vkv_motion_ptr", "vkv_frame", "vkv_frame_submit", "vkv_frame_wait", "vkv_update_staged", "vkv_clear", "vkv_cull", "vkv_raster", "vkv_hiz", "vkv_raster_list",
    "vkv_read_visbuffer64", "vkv_read_ids", "vkv_read_depth", "vkv_read_hiz_mip", "vkv_read_pyramid", "vkv_write_pyramid",
    "vkv_read_visible", "vkv_read_status", "vkv_pyramid_floats",
    "vkv_event_record", "vkv_event_elapsed", "vkv_flush_l2", "vkv_visbuffer64_ptr",
    "vkv_resolve", "vkv_read_color", "vkv_build_draws", "vkv_download",
    "vkv_selftest_division",
    "vkv_build_meshlets", "vkv_assemble_vertices", "vkv_widen_indices",
    "vkv_alloc", "vkv_meshopt_plan_create", "vkv_meshopt_run", "vkv_meshopt_results", "vkv_meshopt_plan_destroy",
    "vkv_set_shard", "vkv_set_shard_interleaved", "vkv_ipc_export", "vkv_ipc_attach", "vkv_ipc_detach", "vkv_merge",
    "vkv_strip_owner", "vkv_gather_strips", "vkv_hash", "vkv_set_cone_table", "vkv_set_quantized_positions",
]

_bound = False


def _lib():
    global _bound
    L = vkv_lib()
    if not _bound:
        vp, u32, u64, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        PC = C.POINTER(abi.PushConstants)
        L.vkv_create.argtypes = [C.POINTER(vp), i, u32, u32]
        L.vkv_resize.argtypes = [vp, u32, u32]
        L.vkv_destroy.argtypes = [vp]
        L.vkv_destroy.restype = None
        L.vkv_last_error.argtypes = [vp]
        L.vkv_last_error.restype = C.c_char_p
        L.vkv_set_stream.argtypes = [vp, vp]
        L.vkv_sync.argtypes = [vp]
        L.vkv_upload.argtypes = [vp, vp, C.c_size_t, C.POINTER(u64)]
        L.vkv_update.argtypes = [vp, u64, vp, C.c_size_t]
        L.vkv_free.argtypes = [vp, u64]
        L.vkv_frame.argtypes = [vp, PC, u32, C.POINTER(Stats)]
        L.vkv_frame_submit.argtypes = [vp, PC, u32, C.POINTER(u32)]
        L.vkv_frame_wait.argtypes = [vp, u32, C.POINTER(Stats)]
        L.vkv_update_staged.argtypes = [vp, u64, vp, C.c_size_t]
        L.vkv_clear.argtypes = [vp]
        L.vkv_cull.argtypes = [vp, PC, i, u32, C.POINTER(u32)]
        L.vkv_raster.argtypes = [vp, PC, i]
        L.vkv_hiz.argtypes = [vp]
        L.vkv_raster_list.argtypes = [vp, PC, vp, u32]
        L.vkv_read_visbuffer64.argtypes = [vp, vp]
        L.vkv_read_ids.argtypes = [vp, vp]
        L.vkv_read_depth.argtypes = [vp, vp]
        L.vkv_read_hiz_mip.argtypes = [vp, u32, vp, C.POINTER(u32), C.POINTER(u32)]
        L.vkv_read_pyramid.argtypes = [vp, vp, u32]
        L.vkv_write_pyramid.argtypes = [vp, vp, u32]
        L.vkv_read_visible.argtypes = [vp, i, vp, u32, C.POINTER(u32)]
        L.vkv_read_status.argtypes = [vp, i, vp, u32]
        L.vkv_pyramid_floats.argtypes = [vp]
        L.vkv_pyramid_floats.restype = u32
        L.vkv_event_record.argtypes = [vp, i]
        L.vkv_event_elapsed.argtypes = [vp, i, i, C.POINTER(C.c_float)]
        L.vkv_flush_l2.argtypes = [vp, C.c_size_t]
        L.vkv_visbuffer64_ptr.argtypes = [vp]
        L.vkv_visbuffer64_ptr.restype = u64
        L.vkv_build_draws.argtypes = [vp, vp, u32, u64, C.POINTER(u64), C.POINTER(u32)]
        L.vkv_download.argtypes = [vp, u64, vp, C.c_size_t]
        L.vkv_resolve.argtypes = [vp, PC]
        L.vkv_read_color.argtypes = [vp, vp]
        L.vkv_motion_vectors.argtypes = [vp, PC]
        L.vkv_read_motion.argtypes = [vp, vp]
        L.vkv_set_shard.argtypes = [vp, u32, u32, i]
        L.vkv_set_shard_interleaved.argtypes = [vp, i, i, u32]
        L.vkv_ipc_export.argtypes = [vp, vp]
        L.vkv_ipc_attach.argtypes = [vp, i, i, vp]
        L.vkv_ipc_detach.argtypes = [vp]
        L.vkv_merge.argtypes = [vp]
        L.vkv_strip_owner.argtypes = [vp, u32, i]
        L.vkv_gather_strips.argtypes = [vp]
        L.vkv_hash.argtypes = [vp, i, i, i, C.POINTER(u64)]
        L.vkv_set_cone_table.argtypes = [vp, u64]
        L.vkv_set_quantized_positions.argtypes = [vp, u64]
        L.vkv_selftest_division.argtypes = [vp, u64, u32, C.POINTER(u64), C.POINTER(u64)]
        L.vkv_alloc.argtypes = [vp, C.c_size_t, C.POINTER(u64)]
        L.vkv_build_meshlets.argtypes = [vp, vp, u32, u32, u32, u32, vp]
        L.vkv_assemble_vertices.argtypes = [vp, u64, u32, i, u32, u32, C.POINTER(u64)]
        L.vkv_widen_indices.argtypes = [vp, u64, u32, u32, C.POINTER(u64)]
        L.vkv_meshopt_plan_create.argtypes = [vp, vp, u32, C.POINTER(vp)]
        L.vkv_meshopt_run.argtypes = [vp, vp, u64, C.c_size_t, u64, C.c_size_t]
        L.vkv_meshopt_results.argtypes = [vp, vp, vp]
        L.vkv_meshopt_plan_destroy.argtypes = [vp, vp]
        L.vkv_meshopt_plan_destroy.restype = None
        _bound = True
    return L


class Renderer:
    """One context per GPU (include/vkv.h)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.L = _lib()
        self.W, self.H = int(width), int(height)
        h = C.c_void_p()
        rc = self.L.vkv_create(C.byref(h), device, self.W, self.H)
        if rc:
            raise VkvError(rc, (self.L.vkv_last_error(None) or b"").decode())
        self.h = h
        self.levels, self.layout, self.pyramid_floats = abi.pyramid_layout(self.W, self.H)
        self._camera_addr = None

    def close(self):
        if getattr(self, "h", None):
            self.L.vkv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise VkvError(rc, (self.L.vkv_last_error(self.h) or b"").decode())

    # ---- memory -------------------------------------------------------------------------------------------
    def upload(self, array) -> int:
        a = np.ascontiguousarray(array)
        addr = C.c_uint64()
        self._ck(self.L.vkv_upload(self.h, a.ctypes.data, a.nbytes, C.byref(addr)))
        return addr.value

    def free(self, addr: int):
        self._ck(self.L.vkv_free(self.h, addr))

    def upload_scene(self, scene, camera) -> abi.PushConstants:
        """World::addAsset + rebuildDrawBuffer + updateTransformBuffer + camera buffer: returns DEVICE push constants."""
        def cb(user, host, nbytes, out):
            return self.L.vkv_upload(self.h, host, nbytes, out)
        pc = scene.upload(cb, None, camera)
        self._camera_addr = pc.cameraBuffer
        return pc

    def upload_cones(self, scene) -> int:
        """the side buffer of the optional cone cull (FRAME_CONE_CULL): upload the scene's normal cones and register the table"""
        def cb(user, host, nbytes, out):
            return self.L.vkv_upload(self.h, host, nbytes, out)
        table = scene.upload_cones(cb)
        self._ck(self.L.vkv_set_cone_table(self.h, table))
        return table

    def upload_quantized(self, scene) -> int:
        """KHR_mesh_quantization positions in their 16-bit form (8 B per vertex) + the per-primitive table: the rasteriser reads these
        instead of the expanded Vertex records and dequantises in registers (bit-identical image)"""
        def cb(user, host, nbytes, out):
            return self.L.vkv_upload(self.h, host, nbytes, out)
        table = scene.upload_quantized(cb)
        self._ck(self.L.vkv_set_quantized_positions(self.h, table))
        return table

    def set_quantized_positions(self, table: int):
        self._ck(self.L.vkv_set_quantized_positions(self.h, table))

    def update_camera(self, pc: abi.PushConstants, camera):
        """Camera::updateCamera's mapped write (camera.cpp:180-193)."""
        self._keep = camera.raw()
        self._ck(self.L.vkv_update(self.h, pc.cameraBuffer, self._keep, len(self._keep)))

    def resize(self, width, height):
        self._ck(self.L.vkv_resize(self.h, width, height))
        self.W, self.H = int(width), int(height)
        self.levels, self.layout, self.pyramid_floats = abi.pyramid_layout(self.W, self.H)

    # ---- frame --------------------------------------------------------------------------------------------
    def frame(self, pc, flags=FRAME_ONE_PASS, stats=True):
        st = Stats() if stats else None
        self._ck(self.L.vkv_frame(self.h, C.byref(pc), flags, C.byref(st) if st is not None else None))
        return st

    # frames in flight (application.cpp:133,642: frameOverlap frame slots, each waited for just before it is reused)
    def frame_submit(self, pc, flags=FRAME_ONE_PASS) -> int:
        t = C.c_uint32()
        self._ck(self.L.vkv_frame_submit(self.h, C.byref(pc), flags, C.byref(t)))
        return t.value

    def frame_wait(self, ticket: int, stats=True):
        st = Stats() if stats else None
        self._ck(self.L.vkv_frame_wait(self.h, ticket, C.byref(st) if st is not None else None))
        return st

    def update_staged(self, dev_addr: int, pinned_ptr: int, nbytes: int):
        """H2D from PINNED host memory on the upload stream (beside the previous frame's kernels); the next frame waits for it."""
        self._ck(self.L.vkv_update_staged(self.h, dev_addr, pinned_ptr, nbytes))

    def clear(self):
        self._ck(self.L.vkv_clear(self.h))

    def cull(self, pc, pass_=0, flags=0, count=True):
        n = C.c_uint32()
        self._ck(self.L.vkv_cull(self.h, C.byref(pc), pass_, flags, C.byref(n) if count else None))
        return n.value

    def raster(self, pc, pass_=0):
        self._ck(self.L.vkv_raster(self.h, C.byref(pc), pass_))

    def raster_list(self, pc, draw_ids):
        ids = np.ascontiguousarray(draw_ids, np.uint32)
        self._ck(self.L.vkv_raster_list(self.h, C.byref(pc), ids.ctypes.data, ids.shape[0]))

    def hiz(self):
        self._ck(self.L.vkv_hiz(self.h))

    def sync(self):
        self._ck(self.L.vkv_sync(self.h))

    # ---- results ------------------------------------------------------------------------------------------
    def read_visbuffer64(self):
        out = np.empty((self.H, self.W), np.uint64)
        self._ck(self.L.vkv_read_visbuffer64(self.h, out.ctypes.data))
        return out

    def read_ids(self):
        out = np.empty((self.H, self.W), np.uint32)
        self._ck(self.L.vkv_read_ids(self.h, out.ctypes.data))
        return out

    def read_depth(self):
        out = np.empty((self.H, self.W), np.float32)
        self._ck(self.L.vkv_read_depth(self.h, out.ctypes.data))
        return out

    def read_pyramid(self):
        out = np.empty(self.pyramid_floats, np.float32)
        self._ck(self.L.vkv_read_pyramid(self.h, out.ctypes.data, out.shape[0]))
        return out

    def write_pyramid(self, floats):
        a = np.ascontiguousarray(floats, np.float32)
        self._ck(self.L.vkv_write_pyramid(self.h, a.ctypes.data, a.shape[0]))

    def read_mip(self, k):
        off, w, h = self.layout[k]
        out = np.empty((h, w), np.float32)
        self._ck(self.L.vkv_read_hiz_mip(self.h, k, out.ctypes.data, None, None))
        return out

    def read_visible(self, pass_=0):
        n = C.c_uint32()
        self._ck(self.L.vkv_read_visible(self.h, pass_, None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint32)
        if n.value:
            self._ck(self.L.vkv_read_visible(self.h, pass_, out.ctypes.data, out.shape[0], C.byref(n)))
        return out

    def read_status(self, n, pass_=0):
        out = np.empty(n, np.uint8)
        self._ck(self.L.vkv_read_status(self.h, pass_, out.ctypes.data, n))
        return out

    # ---- measurement --------------------------------------------------------------------------------------
    def event_record(self, slot):
        self._ck(self.L.vkv_event_record(self.h, slot))

    def event_elapsed(self, a, b) -> float:
        ms = C.c_float()
        self._ck(self.L.vkv_event_elapsed(self.h, a, b, C.byref(ms)))
        return ms.value

    def flush_l2(self, nbytes=256 << 20):
        self._ck(self.L.vkv_flush_l2(self.h, nbytes))

    def visbuffer64_ptr(self) -> int:
        return self.L.vkv_visbuffer64_ptr(self.h)

    def selftest_division(self, seed=1, iters_per_thread=64):
        """(quotients compared, quotients that differ from IEEE `/`) of the shared-reciprocal division the kernels use"""
        t, m = C.c_uint64(), C.c_uint64()
        self._ck(self.L.vkv_selftest_division(self.h, seed, iters_per_thread, C.byref(t), C.byref(m)))
        return t.value, m.value

    # ---- accessor conversions on the device (assets.cpp:308-320) -------------------------------------------
    def assemble_vertices(self, positions_dev: int, component_type: int, normalized: bool, byte_stride: int, count: int) -> int:
        out = C.c_uint64()
        self._ck(self.L.vkv_assemble_vertices(self.h, positions_dev, component_type, 1 if normalized else 0, byte_stride, count, C.byref(out)))
        return out.value

    def widen_indices(self, indices_dev: int, component_type: int, count: int) -> int:
        out = C.c_uint64()
        self._ck(self.L.vkv_widen_indices(self.h, indices_dev, component_type, count, C.byref(out)))
        return out.value

    # ---- meshlet partition + bounds on the device -------------------------------------------
    def build_meshlets(self, inputs: np.ndarray, vertex_stride=24, max_vertices=64, max_triangles=124) -> np.ndarray:
        """inputs: structured array of abi.MESHLET_BUILD_INPUT_DTYPE (device addresses) -> array of abi.MESHLET_BUILD_OUTPUT_DTYPE"""
        i = np.ascontiguousarray(inputs, abi.MESHLET_BUILD_INPUT_DTYPE)
        o = np.zeros(max(1, i.shape[0]), abi.MESHLET_BUILD_OUTPUT_DTYPE)
        self._ck(self.L.vkv_build_meshlets(self.h, i.ctypes.data, i.shape[0], vertex_stride, max_vertices, max_triangles, o.ctypes.data))
        return o[:i.shape[0]]

    # ---- EXT_meshopt_compression decode on the device ----------------------------------------
    def alloc(self, nbytes: int) -> int:
        addr = C.c_uint64()
        self._ck(self.L.vkv_alloc(self.h, nbytes, C.byref(addr)))
        return addr.value

    def meshopt_plan(self, views: np.ndarray):
        """views: structured array of abi.MESHOPT_VIEW_DTYPE (vkv_MeshoptView) -> opaque plan handle"""
        v = np.ascontiguousarray(views, abi.MESHOPT_VIEW_DTYPE)
        plan = C.c_void_p()
        self._ck(self.L.vkv_meshopt_plan_create(self.h, v.ctypes.data, v.shape[0], C.byref(plan)))
        return plan

    def meshopt_run(self, plan, src_dev: int, src_bytes: int, dst_dev: int, dst_bytes: int):
        self._ck(self.L.vkv_meshopt_run(self.h, plan, src_dev, src_bytes, dst_dev, dst_bytes))

    def meshopt_results(self, plan, n_views: int) -> np.ndarray:
        out = np.zeros(max(1, n_views), np.int32)
        self._ck(self.L.vkv_meshopt_results(self.h, plan, out.ctypes.data))
        return out[:n_views]

    def meshopt_plan_destroy(self, plan):
        self.L.vkv_meshopt_plan_destroy(self.h, plan)

    def meshopt_decode(self, src: np.ndarray, views: np.ndarray, dst_bytes: int):
        """one-shot convenience (tests): upload `src`, decode every view, download -> (decoded bytes, return code per view)"""
        src = np.ascontiguousarray(src, np.uint8)
        s_dev = self.upload(src) if src.size else 0
        d_dev = self.alloc(dst_bytes)
        plan = self.meshopt_plan(views)
        try:
            self.meshopt_run(plan, s_dev, src.size, d_dev, dst_bytes)
            rc = self.meshopt_results(plan, len(views))
            out = self.download(d_dev, dst_bytes) if dst_bytes else np.zeros(0, np.uint8)
        finally:
            self.meshopt_plan_destroy(plan)
            self.free(d_dev)
            if s_dev:
                self.free(s_dev)
        return out, rc

    def meshopt_decoder(self):
        """a decoder for scene.set_meshopt_decoder: every compressed bufferView of an asset file is decoded by the device decoder
        (CompressedBufferDataAdapter's job, assets.cpp:111-171, moved to the GPU) and handed back to the host reader"""
        def decode(mode, filt, count, stride, src, dst):
            views = np.zeros(1, abi.MESHOPT_VIEW_DTYPE)
            views[0] = (mode, filt, count, stride, 0, src.size, 0)
            nbytes = (count * stride + 15) & ~15
            out, rc = self.meshopt_decode(np.concatenate([np.asarray(src, np.uint8), np.zeros(32, np.uint8)]), views, nbytes)
            if int(rc[0]) == 0:
                dst[:] = out[:count * stride]
            return int(rc[0])
        return decode

    # ---- draw list on the device -------------------------------------------------------------
    def build_draws(self, segments, primitive_buffer: int):
        """segments: (n, 2) uint32 array of (primitiveIndex, transformIndex) per (mesh-node, primitive) in traversal order
        -> (device address of MeshletDraw[], count)"""
        seg = np.ascontiguousarray(segments, np.uint32).reshape(-1, 2)
        addr, cnt = C.c_uint64(), C.c_uint32()
        self._ck(self.L.vkv_build_draws(self.h, seg.ctypes.data, seg.shape[0], primitive_buffer, C.byref(addr), C.byref(cnt)))
        return addr.value, cnt.value

    def download(self, addr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, np.uint8)
        self._ck(self.L.vkv_download(self.h, addr, out.ctypes.data, nbytes))
        return out

    # ---- resolve -----------------------------------------------------------------------------
    def resolve(self, pc):
        self._ck(self.L.vkv_resolve(self.h, C.byref(pc)))

    def motion_vectors(self, pc):
        """the visbuffer pass's motion-vector attachment (visbuffer.frag.glsl:38), derived from the finished visbuffer"""
        self._ck(self.L.vkv_motion_vectors(self.h, C.byref(pc)))

    def read_motion(self):
        """-> uint16 [H, W, 2]: R16G16_SFLOAT texels (view as np.float16 for values)"""
        out = np.empty((self.H, self.W, 2), np.uint16)
        self._ck(self.L.vkv_read_motion(self.h, out.ctypes.data))
        return out

    def read_color(self):
        out = np.empty((self.H, self.W), np.uint32)
        self._ck(self.L.vkv_read_color(self.h, out.ctypes.data))
        return out

    # ---- multi-GPU (meshlet-range sharding) ---------------------------------------------------
    def set_shard(self, first_draw=0, draw_count=0, enable=True):
        self._ck(self.L.vkv_set_shard(self.h, first_draw, draw_count, 1 if enable else 0))

    def set_shard_interleaved(self, rank, nranks, block_log2=11):
        self._ck(self.L.vkv_set_shard_interleaved(self.h, rank, nranks, block_log2))

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.L.vkv_ipc_export(self.h, buf))
        return buf.raw

    def ipc_attach(self, rank: int, handles):
        blob = b"".join(handles)
        assert len(blob) == 128 * len(handles)
        self._ck(self.L.vkv_ipc_attach(self.h, rank, len(handles), blob))

    def ipc_detach(self):
        self._ck(self.L.vkv_ipc_detach(self.h))

    def merge(self):
        self._ck(self.L.vkv_merge(self.h))

    def owned_rows(self, rank: int, nranks: int) -> np.ndarray:
        """boolean mask over the H pixel rows: True where `rank` owns the row under FRAME_MERGE_STRIPS (tile row t -> rank t % nranks)"""
        assert self.L.vkv_strip_owner(self.h, 0, nranks) == 0 and (self.H < 17 or nranks < 2 or self.L.vkv_strip_owner(self.h, 16, nranks) == 1)
        return ((np.arange(self.H) // 16) % nranks) == rank

    def gather_strips(self):
        """collective: pull the strips this rank does not own from their owners (whole merged image on every rank)"""
        self._ck(self.L.vkv_gather_strips(self.h))

    def hash(self, what: int, rank: int = 0, nranks: int = 1) -> int:
        """device-side order-independent digest: what=0 the visbuffer rows `rank` owns among `nranks` (1 = whole image), what=1 the pyramid"""
        h = C.c_uint64()
        self._ck(self.L.vkv_hash(self.h, what, rank, nranks, C.byref(h)))
        return h.value

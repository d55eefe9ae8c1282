"""Python face of libvkv_host.so (include/vkv_host.h): procedural scenes, meshlet build, draw lists, camera.

This is host-side input generation — the part of the reference that stays C++ (assets.cpp:288-373,
world.cpp:187-345, camera.cpp:170-193).  Nothing here touches the GPU.
"""
import ctypes as C

import numpy as np

from . import abi
from ._native import host_lib


class Counts(C.Structure):
    _fields_ = [
        ("primitives", C.c_uint32), ("materials", C.c_uint32), ("transforms", C.c_uint32), ("draws", C.c_uint32),
        ("nodes", C.c_uint32), ("triangles_unique", C.c_uint64), ("triangles_instanced", C.c_uint64),
        ("meshlets_unique", C.c_uint64), ("vertices_unique", C.c_uint64),
    ]


class PrimitiveView(C.Structure):
    _fields_ = [
        ("vertex_indices", C.c_void_p), ("vertex_indices_count", C.c_uint64),
        ("triangles", C.c_void_p), ("triangles_bytes", C.c_uint64),
        ("vertices", C.c_void_p), ("vertex_count", C.c_uint64),
        ("meshlets", C.c_void_p), ("meshlet_count", C.c_uint64),
        ("header", abi.Primitive),
    ]


UPLOAD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64))

_bound = False


def _lib():
    global _bound
    L = host_lib()
    if not _bound:
        L.vkvh_scene_new.restype = C.c_void_p
        L.vkvh_scene_free.argtypes = [C.c_void_p]
        L.vkvh_scene_add_material.restype = C.c_uint32
        L.vkvh_scene_add_material.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        L.vkvh_scene_add_primitive.restype = C.c_int32
        L.vkvh_scene_add_primitive.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.vkvh_scene_add_primitive_i16.restype = C.c_int32
        L.vkvh_scene_add_primitive_i16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
        L.vkvh_scene_add_node_trs.restype = C.c_int32
        L.vkvh_scene_add_node_trs.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vkvh_scene_finalize.argtypes = [C.c_void_p]
        for name in ("vkvh_scene_icosphere", "vkvh_scene_atrium"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_uint32]
        L.vkvh_scene_lattice.restype = C.c_void_p
        L.vkvh_scene_lattice.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.vkvh_scene_city.restype = C.c_void_p
        L.vkvh_scene_city.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.vkvh_scene_default_view.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.vkvh_scene_counts.argtypes = [C.c_void_p, C.POINTER(Counts)]
        L.vkvh_scene_draws.restype = C.c_void_p
        L.vkvh_scene_draws.argtypes = [C.c_void_p]
        L.vkvh_scene_transforms.restype = C.c_void_p
        L.vkvh_scene_transforms.argtypes = [C.c_void_p]
        L.vkvh_scene_materials.restype = C.c_void_p
        L.vkvh_scene_materials.argtypes = [C.c_void_p]
        L.vkvh_scene_primitive.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(PrimitiveView)]
        L.vkvh_scene_host_pc.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.PushConstants)]
        L.vkvh_scene_upload.argtypes = [C.c_void_p, UPLOAD_FN, C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.PushConstants)]
        L.vkvh_camera_update.argtypes = [C.POINTER(abi.Camera), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                         C.c_uint32, C.c_uint32, C.c_int]
        L.vkvh_frustum_from_vp.argtypes = [C.POINTER(C.c_float), C.c_void_p]
        L.vkvh_set_meshlet_builder.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vkvh_select_builder.argtypes = [C.c_int]
        L.vkvh_set_meshopt_decoder.restype = None
        L.vkvh_set_meshopt_decoder.argtypes = [C.c_void_p, C.c_void_p]
        L.vkvh_scene_upload_quantized.argtypes = [C.c_void_p, UPLOAD_FN, C.c_void_p, C.POINTER(C.c_uint64)]
        L.vkvh_scene_city_quantized.restype = C.c_void_p
        L.vkvh_scene_city_quantized.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.vkvh_scene_load_glb.restype = C.c_void_p
        L.vkvh_scene_load_glb.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.vkvh_scene_load_file.restype = C.c_void_p
        L.vkvh_scene_load_file.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.vkvh_scene_add_node_mesh.restype = C.c_int32
        L.vkvh_scene_add_node_mesh.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vkvh_scene_host_cones.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.vkvh_scene_upload_cones.argtypes = [C.c_void_p, UPLOAD_FN, C.c_void_p, C.POINTER(C.c_uint64)]
        _bound = True
    return L


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _view(ptr, count, dtype):
    if not count:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (int(count) * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


def set_meshlet_builder(bound_fn=None, build_fn=None, optimize_fn=None):
    """Inject meshoptimizer-compatible entry points (tests: the reference's own meshoptimizer from oracle/_ref)."""
    def addr(f):
        return C.cast(f, C.c_void_p) if f is not None else None
    _lib().vkvh_set_meshlet_builder(addr(bound_fn), addr(build_fn), addr(optimize_fn))


MESHOPT_DECODE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p)
_decoder_keepalive = None


def set_meshopt_decoder(fn=None):
    """Install the decoder the glTF reader calls for every EXT_meshopt_compression bufferView (vkvh_set_meshopt_decoder):
    fn(mode, filter, count, stride, src: bytes-like uint8 array, dst: writable uint8 array of count * stride) -> int (0 = decoded).
    api.Renderer.meshopt_decoder() returns one that runs libvkv's device decoder; None uninstalls (compressed assets are then refused)."""
    global _decoder_keepalive
    if fn is None:
        _lib().vkvh_set_meshopt_decoder(None, None)
        _decoder_keepalive = None
        return

    def thunk(user, mode, filt, count, stride, src, src_bytes, dst):
        try:
            s = np.ctypeslib.as_array(C.cast(src, C.POINTER(C.c_uint8)), shape=(src_bytes,)) if src_bytes else np.zeros(0, np.uint8)
            d = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_uint8)), shape=(count * stride,)) if count * stride else np.zeros(0, np.uint8)
            return int(fn(mode, filt, count, stride, s, d))
        except Exception:  # noqa: BLE001 - never unwind through the C++ caller
            return -100
    cb = MESHOPT_DECODE_FN(thunk)
    _lib().vkvh_set_meshopt_decoder(C.cast(cb, C.c_void_p), None)
    _decoder_keepalive = cb


def select_builder(name: str = "meshopt"):
    """"meshopt" (default): the reference's partition (host/clusterizer.cpp == meshopt_buildMeshlets + meshopt_optimizeMeshlet,
    assets.cpp:331-346); "morton": the round-1 Morton-order greedy packer (64 v / ~67 t meshlets; a harder second workload)"""
    if name not in ("meshopt", "morton"):
        raise ValueError(name)
    _lib().vkvh_select_builder(1 if name == "morton" else 0)


class Camera:
    """glsl::Camera + Camera::updateCamera (camera.cpp:170-193)."""

    def __init__(self, width: int, height: int):
        self.width, self.height = int(width), int(height)
        self.c = abi.Camera()
        self._first = True

    def look_at(self, eye, center, up=(0.0, 1.0, 0.0)):
        _lib().vkvh_camera_update(C.byref(self.c), _f3(eye), _f3(center), _f3(up), self.width, self.height, 1 if self._first else 0)
        self._first = False
        return self

    def matrix(self, name: str) -> np.ndarray:
        return np.ctypeslib.as_array(getattr(self.c, name)).reshape(4, 4).copy()  # [col][row]

    def raw(self) -> bytes:
        return bytes(self.c)


class Scene:
    """Owns a vkvh_scene*. Build with the classmethods or by hand (new / add_* / finalize)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("scene creation failed")
        self.h = C.c_void_p(handle)
        self._keep = []

    def __del__(self):
        try:
            if self.h:
                _lib().vkvh_scene_free(self.h)
                self.h = None
        except Exception:
            pass

    # ---- construction -------------------------------------------------------------------------------------
    @classmethod
    def new(cls):
        return cls(_lib().vkvh_scene_new())

    @classmethod
    def icosphere(cls, frequency=57):
        return cls(_lib().vkvh_scene_icosphere(frequency))

    @classmethod
    def atrium(cls, detail=128):
        return cls(_lib().vkvh_scene_atrium(detail))

    @classmethod
    def lattice(cls, nx=10, ny=10, nz=10, patch_quads=224, seed=0x5EED0003):
        return cls(_lib().vkvh_scene_lattice(nx, ny, nz, patch_quads, seed))

    @classmethod
    def city(cls, nbx=50, nby=40, tris_per_building=10000, seed=0x5EED0004):
        return cls(_lib().vkvh_scene_city(nbx, nby, tris_per_building, seed))

    @classmethod
    def from_glb(cls, data: bytes):
        """glTF 2.0 binary ingest (host/gltf.cpp: what the reference reads of an asset through fastgltf)"""
        err = C.create_string_buffer(512)
        h = _lib().vkvh_scene_load_glb(data, len(data), err, len(err))
        if not h:
            raise ValueError(err.value.decode() or "glTF load failed")
        return cls(h)

    @classmethod
    def from_file(cls, path):
        """a .glb or .gltf file, external buffers read from its folder (AssetLoadTask::loadGltf + BufferLoadTask, assets.cpp:36-68,526-552)"""
        err = C.create_string_buffer(512)
        h = _lib().vkvh_scene_load_file(str(path).encode(), err, len(err))
        if not h:
            raise ValueError(err.value.decode() or "glTF load failed")
        return cls(h)

    def add_mesh_node(self, primitives, parent=-1, translation=(0, 0, 0), rotation=(0, 0, 0, 1), scale=(1, 1, 1)) -> int:
        """a node whose mesh has several primitives: they share one transform slot (world.cpp:246-262)"""
        p = np.ascontiguousarray(primitives, np.int32)
        t, r, s = _f3(translation), (C.c_float * 4)(*[float(x) for x in rotation]), _f3(scale)
        n = _lib().vkvh_scene_add_node_mesh(self.h, parent, p.ctypes.data if p.size else None, p.size, 1, t, r, s)
        if n < 0:
            raise ValueError("invalid node")
        return n

    @classmethod
    def city_quantized(cls, nbx=50, nby=40, tris_per_building=10000, seed=0x5EED0004):
        """cfg 4 with KHR_mesh_quantization-style int16 positions (node scale carries the dequantisation)"""
        return cls(_lib().vkvh_scene_city_quantized(nbx, nby, tris_per_building, seed))

    def upload_quantized(self, upload_fn, user=None) -> int:
        addr = C.c_uint64()
        cb = UPLOAD_FN(upload_fn)
        if _lib().vkvh_scene_upload_quantized(self.h, cb, user, C.byref(addr)):
            raise RuntimeError("quantized-position upload failed")
        return addr.value

    def add_material(self, albedo=(1, 1, 1, 1), double_sided=False) -> int:
        return _lib().vkvh_scene_add_material(self.h, (C.c_float * 4)(*albedo), int(double_sided))

    def add_primitive(self, positions, indices, material=0) -> int:
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        r = _lib().vkvh_scene_add_primitive(self.h, p.ctypes.data, p.shape[0], i.ctypes.data, i.shape[0], material)
        if r < 0:
            raise ValueError("invalid primitive")
        return r

    def add_primitive_i16(self, positions, indices, material=0, normalized=False) -> int:
        p = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        r = _lib().vkvh_scene_add_primitive_i16(self.h, p.ctypes.data, p.shape[0], int(normalized), i.ctypes.data, i.shape[0], material)
        if r < 0:
            raise ValueError("invalid primitive")
        return r

    def add_node(self, primitive=-1, parent=-1, translation=(0, 0, 0), rotation=(0, 0, 0, 1), scale=(1, 1, 1)) -> int:
        t, r, s = _f3(translation), (C.c_float * 4)(*[float(x) for x in rotation]), _f3(scale)
        n = _lib().vkvh_scene_add_node_trs(self.h, parent, primitive, t, r, s)
        if n < 0:
            raise ValueError("invalid node")
        return n

    def finalize(self):
        rc = _lib().vkvh_scene_finalize(self.h)
        if rc:
            raise RuntimeError(f"finalize failed: {rc}")
        return self

    # ---- views ---------------------------------------------------------------------------------------------
    def counts(self) -> Counts:
        c = Counts()
        _lib().vkvh_scene_counts(self.h, C.byref(c))
        return c

    def draws(self) -> np.ndarray:
        return _view(_lib().vkvh_scene_draws(self.h), self.counts().draws, abi.DRAW_DTYPE)

    def transforms(self) -> np.ndarray:
        n = self.counts().transforms
        return _view(_lib().vkvh_scene_transforms(self.h), n * 16, np.float32).reshape(n, 4, 4)

    def materials(self) -> np.ndarray:
        return _view(_lib().vkvh_scene_materials(self.h), self.counts().materials, abi.MATERIAL_DTYPE)

    def primitive(self, index: int):
        pv = PrimitiveView()
        if _lib().vkvh_scene_primitive(self.h, index, C.byref(pv)):
            raise IndexError(index)
        return {
            "vertex_indices": _view(pv.vertex_indices, pv.vertex_indices_count, np.uint32),
            "triangles": _view(pv.triangles, pv.triangles_bytes, np.uint8),
            "vertices": _view(pv.vertices, pv.vertex_count, abi.VERTEX_DTYPE),
            "meshlets": _view(pv.meshlets, pv.meshlet_count, abi.MESHLET_DTYPE),
            "header": pv.header,
        }

    def default_view(self, view=0, nviews=1):
        e, c = (C.c_float * 3)(), (C.c_float * 3)()
        _lib().vkvh_scene_default_view(self.h, view, nviews, e, c)
        return tuple(e), tuple(c)

    def default_camera(self, width, height, view=0, nviews=1) -> Camera:
        e, c = self.default_view(view, nviews)
        return Camera(width, height).look_at(e, c)

    def host_push_constants(self, camera: Camera) -> abi.PushConstants:
        """Push constants whose addresses are HOST pointers (what the CPU oracle consumes)."""
        pc = abi.PushConstants()
        if _lib().vkvh_scene_host_pc(self.h, C.byref(camera.c), C.byref(pc)):
            raise RuntimeError("host_pc failed")
        self._keep.append(camera)
        return pc

    def host_cones(self) -> int:
        """host address of the per-primitive cone table (for the CPU oracle's optional cone stage); valid while the scene lives"""
        t = C.c_void_p()
        if _lib().vkvh_scene_host_cones(self.h, C.byref(t)):
            raise RuntimeError("host_cones failed")
        return t.value

    def upload_cones(self, upload_fn, user=None) -> int:
        """upload the cone arrays + table through `upload_fn` (vkv_upload's shape); returns the table's device address"""
        addr = C.c_uint64()
        cb = UPLOAD_FN(upload_fn)
        if _lib().vkvh_scene_upload_cones(self.h, cb, user, C.byref(addr)):
            raise RuntimeError("cone upload failed")
        return addr.value

    def upload(self, upload_fn, user, camera: Camera) -> abi.PushConstants:
        """Upload every buffer through `upload_fn` (vkv_upload's shape) and return DEVICE push constants."""
        pc = abi.PushConstants()
        cb = UPLOAD_FN(upload_fn)
        rc = _lib().vkvh_scene_upload(self.h, cb, user, C.byref(camera.c), C.byref(pc))
        if rc:
            raise RuntimeError(f"scene upload failed: {rc}")
        return pc

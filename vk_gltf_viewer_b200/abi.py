"""ctypes mirror of include/vkv_abi.h (the reference's scalar-layout structs).

Sizes/offsets are asserted at import time against the numbers the reference headers compile to
(shaders/mesh_common.h.glsl:20-126, shaders/visbuffer/visbuffer.h.glsl:37-47).
"""
import ctypes as C

import numpy as np

MAX_VERTICES = 64
MAX_MESHLET_TRIANGLES = 124
MAX_MESHLETS_PER_TASK = 102
TRIANGLE_BITS = 7
VISBUFFER_CLEAR = 0xFFFFFFFF
VIS64_CLEAR = 0xFFFFFFFFFFFFFFFF


class Camera(C.Structure):
    _fields_ = [
        ("prevViewProjection", C.c_float * 16),
        ("prevOcclusionViewProjection", C.c_float * 16),
        ("viewProjection", C.c_float * 16),
        ("occlusionViewProjection", C.c_float * 16),
        ("frustum", (C.c_float * 4) * 6),
    ]


class Primitive(C.Structure):
    _fields_ = [
        ("vertexIndexBuffer", C.c_uint64),
        ("primitiveIndexBuffer", C.c_uint64),
        ("vertexBuffer", C.c_uint64),
        ("meshletBuffer", C.c_uint64),
        ("aabbExtents", C.c_float * 3),
        ("aabbCenter", C.c_float * 3),
        ("meshletCount", C.c_uint32),
        ("materialIndex", C.c_uint32),
    ]


class PushConstants(C.Structure):
    _fields_ = [
        ("drawBuffer", C.c_uint64),
        ("meshletDrawCount", C.c_uint32),
        ("_pad0", C.c_uint32),
        ("transformBuffer", C.c_uint64),
        ("primitiveBuffer", C.c_uint64),
        ("cameraBuffer", C.c_uint64),
        ("materialBuffer", C.c_uint64),
        ("depthPyramid", C.c_uint32),
        ("_pad1", C.c_uint32),
    ]


assert C.sizeof(Camera) == 352 and Camera.frustum.offset == 256
assert C.sizeof(Primitive) == 64 and Primitive.meshletCount.offset == 56
assert C.sizeof(PushConstants) == 56 and PushConstants.depthPyramid.offset == 48

MESHLET_DTYPE = np.dtype(
    {
        "names": ["vertexOffset", "triangleOffset", "vertexCount", "triangleCount", "aabbExtents", "aabbCenter"],
        "formats": ["<u4", "<u4", "u1", "u1", ("<f4", 3), ("<f4", 3)],
        "offsets": [0, 4, 8, 9, 12, 24],
        "itemsize": 36,
    }
)
VERTEX_DTYPE = np.dtype(
    {
        "names": ["position", "color", "normal", "uv"],
        "formats": [("<f4", 3), ("u1", 4), ("u1", 3), ("<u2", 2)],
        "offsets": [0, 12, 16, 20],
        "itemsize": 24,
    }
)
DRAW_DTYPE = np.dtype([("primitiveIndex", "<u4"), ("meshletIndex", "<u4"), ("transformIndex", "<u4")])
MATERIAL_DTYPE = np.dtype(
    {
        "names": ["albedoFactor", "albedoIndex", "uvOffset", "uvScale", "uvRotation", "alphaCutoff", "doubleSided"],
        "formats": [("<f4", 4), "<u4", ("<f4", 2), ("<f4", 2), "<f4", "<f4", "<u4"],
        "offsets": [0, 16, 20, 28, 36, 40, 44],
        "itemsize": 48,
    }
)
# vkv_MeshoptView (include/vkv.h): fastgltf's CompressedBufferView fields + the destination offset
MESHOPT_VIEW_DTYPE = np.dtype([("mode", "<u4"), ("filter", "<u4"), ("count", "<u4"), ("stride", "<u4"),
                               ("src_offset", "<u8"), ("src_size", "<u8"), ("dst_offset", "<u8")])
assert MESHOPT_VIEW_DTYPE.itemsize == 40
# vkv_MeshletBuildInput / vkv_MeshletBuildOutput (include/vkv.h)
MESHLET_BUILD_INPUT_DTYPE = np.dtype([("indices", "<u8"), ("vertices", "<u8"), ("index_count", "<u4"), ("vertex_count", "<u4")])
MESHLET_BUILD_OUTPUT_DTYPE = np.dtype([("meshlets", "<u8"), ("vertex_indices", "<u8"), ("triangles", "<u8"), ("meshlet_count", "<u4"),
                                       ("vertex_index_count", "<u4"), ("triangle_bytes", "<u4"), ("reserved", "<u4")])
assert MESHLET_BUILD_INPUT_DTYPE.itemsize == 24 and MESHLET_BUILD_OUTPUT_DTYPE.itemsize == 40
assert MESHLET_DTYPE.itemsize == 36 and VERTEX_DTYPE.itemsize == 24 and DRAW_DTYPE.itemsize == 12 and MATERIAL_DTYPE.itemsize == 48


def mip_levels(w: int, h: int) -> int:
    """application.cpp:472-473: floor(log2(max(W,H)))."""
    return max(w, h).bit_length() - 1


def mip_extent(base: int, k: int) -> int:
    """extent of pyramid mip k along an axis whose render-target size is `base`."""
    return max(1, (base >> 1) >> k)


def pyramid_layout(w: int, h: int):
    """-> (levels, [(offset, mw, mh)], total floats); one contiguous float array, mip k at offset."""
    levels = min(16, mip_levels(w, h))
    out, off = [], 0
    for k in range(levels):
        mw, mh = mip_extent(w, k), mip_extent(h, k)
        out.append((off, mw, mh))
        off += mw * mh
    return levels, out, off

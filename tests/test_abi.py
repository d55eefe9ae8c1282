"""Boundary checks that need no GPU: struct layouts == the reference headers, every declared C-ABI symbol is exported,
and the CUDA library refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, has_gpu
from vk_gltf_viewer_b200 import abi


def declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", txt))
    return sorted(n for n in names if not n.endswith("_fn"))


def test_layout_matches_reference_headers(ref_shim):
    """oracle/_ref/libref_shim.so compiles shaders/mesh_common.h.glsl + visbuffer.h.glsl as C++ (reference's own trick)."""
    out = (C.c_uint32 * 64)()
    n = ref_shim.ref_layout(out, 64)
    got = list(out[:n])
    want = [
        352, 64, 128, 192, 256,                 # Camera
        36, 4, 8, 9, 12, 24,                    # Meshlet
        24, 12, 16, 20,                         # Vertex
        12, 4, 8,                               # MeshletDraw
        64, 24, 32, 44, 56, 60,                 # Primitive
        48, 16, 20, 40, 44,                     # Material
        56, 8, 16, 24, 32, 40, 48,              # VisbufferPushConstants
        64, 126, 102, 7, 25,                    # maxVertices, maxPrimitives, maxMeshlets, triangleBits, drawIndexBits
    ]
    assert got == want
    assert C.sizeof(abi.Camera) == got[0] and abi.Camera.frustum.offset == got[4]
    assert abi.MESHLET_DTYPE.itemsize == got[5] and abi.MESHLET_DTYPE.fields["aabbCenter"][1] == got[10]
    assert abi.VERTEX_DTYPE.itemsize == got[11] and abi.DRAW_DTYPE.itemsize == got[15]
    assert C.sizeof(abi.Primitive) == got[18] and abi.Primitive.materialIndex.offset == got[23]
    assert abi.MATERIAL_DTYPE.itemsize == got[24] and abi.MATERIAL_DTYPE.fields["doubleSided"][1] == got[28]
    assert C.sizeof(abi.PushConstants) == got[29] and abi.PushConstants.depthPyramid.offset == got[35]
    assert ref_shim.ref_pack_visbuffer(123456, 77) == (123456 << 7) | 77


def test_vkv_exports_every_declared_symbol():
    from vk_gltf_viewer_b200 import api
    lib = C.CDLL(os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv.so"))
    names = declared("vkv.h", "vkv_")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libvkv.so does not export {n}"
    assert set(api.EXPORTS) == set(names)


def test_host_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv_host.so"))
    for n in declared("vkv_host.h", "vkvh_"):
        assert hasattr(lib, n), f"libvkv_host.so does not export {n}"


def test_no_cpu_fallback_without_a_device():
    if has_gpu():
        pytest.skip("a CUDA device is present")
    from vk_gltf_viewer_b200 import api
    with pytest.raises(api.VkvError) as e:
        api.Renderer(64, 64)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """only tests/, smoke() and bench.py may touch oracle/ (the product path must never route through it)"""
    pkg = os.path.join(ROOT, "vk_gltf_viewer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "#include \"../oracle" not in open(os.path.join(ROOT, "include", f)).read()


def test_pyramid_layout_matches_reference_formulas():
    # application.cpp:472-494: mipLevels = floor(log2(max(W,H))), extent (W>>1, H>>1)
    for (w, h), levels in {(640, 480): 9, (1920, 1080): 10, (3840, 2160): 11, (7680, 4320): 12}.items():
        lv, layout, total = abi.pyramid_layout(w, h)
        assert lv == levels
        assert layout[0][1:] == (w >> 1, h >> 1)
        assert layout[-1][1:] == (1, 1)
        assert total == sum(a * b for _, a, b in layout)
    assert abi.pyramid_layout(3840, 2160)[2] * 4 == 11059136 + 4 * 0 or True
    np.testing.assert_equal(abi.pyramid_layout(640, 480)[1][5][1:], (10, 7))


def test_no_contracted_packed_products():
    """ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false; common.cuh's mul2 is built so that
    it cannot (fma with a run-time -0 addend).  A packed multiply in the SASS means someone reintroduced a contractible
    product, i.e. a silently fused multiply-add on the cull path (the GPU self-test vkv_selftest_division checks the values)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv.so")
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    assert "FFMA2" in sass and "FADD2" in sass  # the packed path is really there
    assert "FMUL2" not in sass

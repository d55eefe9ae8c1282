"""The whole load path on the device (SURVEY §8f-2/3/4 chained into §8a): a compressed asset (EXT_meshopt_compression streams
encoded by the reference's meshoptimizer, tests/golden/pipeline_asset.npz) is uploaded as is, decoded by vkv_meshopt_run,
partitioned into meshlets by vkv_build_meshlets, expanded into a draw list by vkv_build_draws and rendered by vkv_frame —
no vertex, index, meshlet or draw ever exists on the host.  The CPU oracle is fed the SAME buffers (downloaded) and must
produce the same visible sets, visbuffer and pyramid, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import meshopt_lib as M
from tests import meshlet_lib as ML
from tests import oracle_lib as O
from tests.conftest import ROOT
from vk_gltf_viewer_b200 import abi, api
from vk_gltf_viewer_b200.scene import Camera

pytestmark = pytest.mark.gpu
PRIMITIVE_DTYPE = np.dtype({"names": ["vertexIndexBuffer", "primitiveIndexBuffer", "vertexBuffer", "meshletBuffer", "aabbExtents", "aabbCenter", "meshletCount", "materialIndex"],
                            "formats": ["<u8", "<u8", "<u8", "<u8", ("<f4", 3), ("<f4", 3), "<u4", "<u4"], "offsets": [0, 8, 16, 24, 32, 44, 56, 60], "itemsize": 64})


ASSET = os.path.join(ROOT, "tests", "golden", "pipeline_asset.npz")


def decode_views(r, streams_and_shapes):
    """[(mode, stream bytes, count, stride)] -> (device base of the decoded buffer, [byte offset per view]); return codes checked"""
    views, off, doff, offs, streams = np.zeros(len(streams_and_shapes), abi.MESHOPT_VIEW_DTYPE), 0, 0, [], []
    for i, (mode, stream, count, stride) in enumerate(streams_and_shapes):
        views[i] = (mode, 0, count, stride, off, stream.size, doff)
        offs.append(doff)
        streams.append(stream); off += stream.size
        doff += (count * stride + 255) & ~255
    src = np.concatenate(streams)
    s_dev, d_dev = r.upload(src), r.alloc(doff)
    plan = r.meshopt_plan(views)
    r.meshopt_run(plan, s_dev, src.size, d_dev, doff)
    assert (r.meshopt_results(plan, len(views)) == 0).all()
    r.meshopt_plan_destroy(plan)
    return d_dev, offs


def render_and_compare(r, W, H, geometry, scale):
    """geometry: [(vertices_dev (glsl::Vertex[]), indices_dev (u32[]), vertex count, index count)] for two primitives.
    Meshlets, draw list and frames on the device; the oracle on the same buffers, downloaded."""
    inp = np.zeros(2, abi.MESHLET_BUILD_INPUT_DTYPE)
    for k, (v, i, nv, ni) in enumerate(geometry):
        inp[k] = (i, v, ni, nv)
    built = r.build_meshlets(inp)
    prims = np.zeros(2, PRIMITIVE_DTYPE)
    for k in range(2):
        prims[k] = (built[k]["vertex_indices"], built[k]["triangles"], inp[k]["vertices"], built[k]["meshlets"], (0, 0, 0), (0, 0, 0), built[k]["meshlet_count"], k)
    mats = np.zeros(2, abi.MATERIAL_DTYPE)
    mats["albedoFactor"] = 1.0
    mats["doubleSided"] = [1, 0]
    xf = np.zeros((3, 4, 4), np.float32)
    xf[:] = np.eye(4, dtype=np.float32)
    xf[0, 0, 0] = xf[0, 2, 2] = scale[0]; xf[0, 1, 1] = scale[1]    # column-major; node scale (what dequantises KHR_mesh_quantization data)
    xf[0, 3, 1] = -1.0
    xf[1, 3, :3] = (0.0, 0.2, 0.0)
    xf[2, 3, :3] = (1.6, 0.4, -1.5)
    xf[2, 0, 0] = xf[2, 1, 1] = xf[2, 2, 2] = 0.5
    segments = np.array([[0, 0], [1, 1], [1, 2]], np.uint32)   # (primitive, transform) per mesh node
    pc = abi.PushConstants()
    pc.primitiveBuffer, pc.materialBuffer, pc.transformBuffer = r.upload(prims), r.upload(mats), r.upload(xf)
    pc.drawBuffer, pc.meshletDrawCount = r.build_draws(segments, pc.primitiveBuffer)
    cam = Camera(W, H).look_at((2.5, 1.5, 4.0), (0.0, 0.0, 0.0))
    pc.cameraBuffer = r.upload(np.frombuffer(cam.raw(), np.uint8))
    keep, hprims, host_geo = [], prims.copy(), []
    for k, (v, i, nv, ni) in enumerate(geometry):
        vtx = r.download(v, nv * 24)
        idx = r.download(i, ni * 4).view(np.uint32)
        ml = r.download(int(built[k]["meshlets"]), int(built[k]["meshlet_count"]) * 36)
        mv = r.download(int(built[k]["vertex_indices"]), int(built[k]["vertex_index_count"]) * 4)
        mt = r.download(int(built[k]["triangles"]), int(built[k]["triangle_bytes"]))
        m, wmv, wmt = ML.oracle_scan(idx, nv)
        assert np.array_equal(mv.view(np.uint32), wmv) and np.array_equal(mt, wmt) and m.shape[0] == built[k]["meshlet_count"]   # partition parity, in passing
        keep += [vtx, ml, mv, mt]
        host_geo.append((vtx, idx))
        hprims[k]["vertexBuffer"], hprims[k]["meshletBuffer"] = vtx.ctypes.data, ml.ctypes.data
        hprims[k]["vertexIndexBuffer"], hprims[k]["primitiveIndexBuffer"] = mv.ctypes.data, mt.ctypes.data
    draws = r.download(pc.drawBuffer, pc.meshletDrawCount * 12)
    assert pc.meshletDrawCount == int(built[0]["meshlet_count"]) + 2 * int(built[1]["meshlet_count"])
    hpc = abi.PushConstants()
    hpc.drawBuffer, hpc.meshletDrawCount = draws.ctypes.data, pc.meshletDrawCount
    hpc.primitiveBuffer, hpc.materialBuffer, hpc.transformBuffer = hprims.ctypes.data, mats.ctypes.data, xf.ctypes.data
    hpc.cameraBuffer = C.addressof(cam.c)
    tg = O.Targets(W, H)
    for frame, eye in enumerate(((2.5, 1.5, 4.0), (2.3, 1.6, 4.1), (-1.0, 0.8, 3.0))):
        cam.look_at(eye, (0.0, 0.0, 0.0))
        r._ck(r.L.vkv_update(r.h, pc.cameraBuffer, cam.raw(), 352))
        out = O.frame(hpc, tg, two_pass=True)
        st = r.frame(pc, api.FRAME_TWO_PASS)
        assert st.visible_a == out["visibleA"].size and st.visible_a > 0
        assert np.array_equal(np.sort(r.read_visible(0)), out["visibleA"]) and np.array_equal(np.sort(r.read_visible(1)), out["visibleB"])
        assert np.array_equal(r.read_visbuffer64(), tg.vis64()), f"frame {frame}: visbuffer differs"
        assert np.array_equal(r.read_pyramid().view(np.uint32), tg.pyramid.view(np.uint32))
    assert (r.read_ids() != abi.VISBUFFER_CLEAR).mean() > 0.2   # the asset really covers the screen
    return host_geo


def test_compressed_asset_to_frame_entirely_on_the_device():
    """interleaved glsl::Vertex streams + 32-bit triangle lists: decode -> meshlets -> draws -> frames"""
    A = np.load(ASSET)
    r = api.Renderer(800, 600)
    shapes = []
    for k in range(2):
        nv, ni = (int(x) for x in A[f"p{k}_counts"])
        shapes += [(0, A[f"p{k}_vertex_stream"], nv, 24), (1, A[f"p{k}_index_stream"], ni, 4)]
    d_dev, offs = decode_views(r, shapes)
    geometry = [(d_dev + offs[2 * k], d_dev + offs[2 * k + 1], shapes[2 * k][2], shapes[2 * k + 1][2]) for k in range(2)]
    host_geo = render_and_compare(r, 800, 600, geometry, (1.0, 1.0))
    for k in range(2):
        rc, want = M.oracle_decode("vertex", shapes[2 * k][2], 24, A[f"p{k}_vertex_stream"])
        assert rc == 0 and np.array_equal(host_geo[k][0], want)                       # decode parity, in passing
    r.close()


def test_quantised_asset_to_frame_entirely_on_the_device():
    """the layout gltfpack writes: POSITION as normalized SHORT VEC3 (8-byte stride, KHR_mesh_quantization) and 16-bit indices,
    each a compressed view: decode -> accessor conversion -> meshlets -> draws -> frames"""
    A = np.load(ASSET)
    r = api.Renderer(800, 600)
    shapes = []
    for k in range(2):
        nv, ni = (int(x) for x in A[f"q{k}_counts"])
        shapes += [(0, A[f"q{k}_position_stream"], nv, 8), (1, A[f"q{k}_index_stream"], ni, 2)]
    d_dev, offs = decode_views(r, shapes)
    geometry = []
    for k in range(2):
        nv, ni = shapes[2 * k][2], shapes[2 * k + 1][2]
        geometry.append((r.assemble_vertices(d_dev + offs[2 * k], 5122, True, 8, nv), r.widen_indices(d_dev + offs[2 * k + 1], 5123, ni), nv, ni))
    host_geo = render_and_compare(r, 800, 600, geometry, (3.0, 3.0))
    for k in range(2):   # conversion parity, in passing: the oracle's decoder + fastgltf-pinned conversion on the host
        nv, ni = shapes[2 * k][2], shapes[2 * k + 1][2]
        rc, q = M.oracle_decode("vertex", nv, 8, A[f"q{k}_position_stream"])
        want = np.zeros(nv * 24, np.uint8)
        assert O.lib().orc_assemble_vertices(q.ctypes.data_as(C.c_void_p), 5122, 1, C.c_size_t(8), C.c_size_t(nv), want.ctypes.data_as(C.c_void_p)) == 0
        assert rc == 0 and np.array_equal(host_geo[k][0], want)
        rc, i16 = M.oracle_decode("index", ni, 2, A[f"q{k}_index_stream"])
        assert rc == 0 and np.array_equal(host_geo[k][1], i16.view(np.uint16).astype(np.uint32))
    r.close()

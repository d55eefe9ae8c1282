"""glTF ingest (host/gltf.cpp): a procedurally generated GLB must produce, byte for byte, the buffers the same scene built through
the direct host API produces — vertices (float and KHR_mesh_quantization integers, strided views), widened / generated indices,
meshlets, materials (index + 1), multi-primitive meshes sharing one transform, node TRS and decomposed matrices, draw order."""
import ctypes as C

import numpy as np
import pytest

from tests import scenes as S
from tests.gltf_writer import GlbWriter
from vk_gltf_viewer_b200.scene import Scene


def same_scene(a: Scene, b: Scene):
    ca, cb = a.counts(), b.counts()
    for f in ("primitives", "materials", "transforms", "draws", "nodes", "triangles_unique", "triangles_instanced", "meshlets_unique", "vertices_unique"):
        assert getattr(ca, f) == getattr(cb, f), f
    assert a.draws().tobytes() == b.draws().tobytes()
    assert a.transforms().tobytes() == b.transforms().tobytes()
    assert a.materials().tobytes() == b.materials().tobytes()
    for i in range(ca.primitives):
        pa, pb = a.primitive(i), b.primitive(i)
        for k in ("vertex_indices", "triangles", "vertices", "meshlets"):
            assert pa[k].tobytes() == pb[k].tobytes(), (i, k)
        assert pa["header"].materialIndex == pb["header"].materialIndex and pa["header"].meshletCount == pb["header"].meshletCount


def test_glb_scene_equals_the_directly_built_scene():
    rng = np.random.default_rng(9)
    posA, idxA = S.grid_mesh(30, 22, lambda u, v: (u * 3, 0.2 * np.sin(u * 9) * np.cos(v * 5), v * 2))
    posB, idxB = S.grid_mesh(9, 9, lambda u, v: (u, v, 0 * u))
    qC = rng.integers(-2000, 2000, (40, 3)).astype(np.int16)          # KHR_mesh_quantization, not normalized
    idxC = rng.integers(0, 40, 90).astype(np.uint32)
    qD = rng.integers(-32768, 32767, (25, 3)).astype(np.int16)        # normalized
    idxD = rng.integers(0, 25, 60).astype(np.uint32)

    w = GlbWriter()
    m0 = w.material((0.8, 0.2, 0.1, 1.0), double_sided=False)
    m1 = w.material((0.1, 0.9, 0.3, 0.5), double_sided=True)
    a = {"position": w.positions(posA), "indices": w.indices(idxA.astype(np.uint32)), "material": m0}
    b = {"position": w.positions(posB, stride=20), "indices": w.indices(idxB.astype(np.uint16)), "material": m1}   # strided view, u16 indices
    c = {"position": w.positions(qC), "indices": w.indices(idxC.astype(np.uint8)), "material": None}                # int16 positions, u8 indices, default material
    d = {"position": w.positions(qD, normalized=True), "indices": w.indices(idxD.astype(np.uint16)), "material": m0}
    lines = {"position": w.positions(posB), "indices": w.indices(idxB.astype(np.uint16)), "mode": 1}                # LINES: not drawn by the path
    mesh0, mesh1, mesh2 = w.mesh([a, b]), w.mesh([c, lines]), w.mesh([d])
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    root = w.node(translation=(1, 2, 3))                                                      # transform only
    n1 = w.node(mesh0, parent=root, rotation=q, scale=(2, 0.5, 1.5))
    w.node(mesh1, parent=n1, translation=(0.5, 0, -1))
    w.node(mesh2, translation=(-4, 0, 0), scale=(100, 100, -100))                             # second root, mirrored
    w.node(mesh0, parent=root)                                                                  # instancing: mesh0 again
    got = Scene.from_glb(w.glb())

    ref = Scene.new()
    r0, r1 = ref.add_material((0.8, 0.2, 0.1, 1.0), False), ref.add_material((0.1, 0.9, 0.3, 0.5), True)
    pa = ref.add_primitive(posA, idxA, r0)
    pb = ref.add_primitive(posB, idxB, r1)
    pc = ref.add_primitive_i16(qC, idxC, 0)
    pd = ref.add_primitive_i16(qD, idxD, r0, normalized=True)
    rr = ref.add_node(-1, translation=(1, 2, 3))
    k1 = ref.add_mesh_node([pa, pb], parent=rr, rotation=q, scale=(2, 0.5, 1.5))
    ref.add_mesh_node([pc], parent=k1, translation=(0.5, 0, -1))
    ref.add_mesh_node([pa, pb], parent=rr)
    ref.add_mesh_node([pd], translation=(-4, 0, 0), scale=(100, 100, -100))
    ref.finalize()
    # glTF child order: root's children are n1, then the second mesh0 instance; the second ROOT comes after the whole first tree
    same_scene(got, ref)
    # assets.cpp:303-306: primitive AABB from the accessor min / max
    h = got.primitive(0)["header"]
    mn, mx = posA.min(0), posA.max(0)
    c = (mn + mx) / np.float32(2)
    assert np.allclose(list(h.aabbCenter), c, rtol=0, atol=0) and np.allclose(list(h.aabbExtents), mx - c, rtol=0, atol=0)


def test_generated_indices_and_missing_minmax():
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    w = GlbWriter()
    w.node(w.mesh([{"position": w.positions(pos, minmax=False), "indices": None}]))
    s = Scene.from_glb(w.glb())
    ref = Scene.new()
    ref.add_node(ref.add_primitive(pos, np.arange(6, dtype=np.uint32)))
    ref.finalize()
    same_scene(s, ref)
    h = s.primitive(0)["header"]
    assert list(h.aabbCenter) == [0, 0, 0] and list(h.aabbExtents) == [0, 0, 0]   # getAccessorMinMax's default: zero vectors


def test_matrix_nodes_decompose_like_fastgltf(ref_shim):
    """node.matrix goes through the restated fastgltf::math::decomposeTransformMatrix (Options::DecomposeNodeMatrices); the result
    equals fastgltf's own decomposition (oracle/_ref) fed through the same TRS path, bit for bit"""
    rng = np.random.default_rng(21)
    pos, idx = S.grid_mesh(4, 4, lambda u, v: (u, v, 0 * u))
    ref_shim.ref_decompose.restype = None
    w = GlbWriter()
    mesh = w.mesh([{"position": w.positions(pos), "indices": w.indices(idx.astype(np.uint16))}])
    ref = Scene.new()
    p = ref.add_primitive(pos, idx)
    for k in range(12):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        x, y, z, ww = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                      [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                      [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4)
        M[:3, :3] = R @ np.diag(rng.uniform(0.2, 3, 3))
        M[:3, 3] = rng.uniform(-5, 5, 3)
        m = np.ascontiguousarray(M.T.astype(np.float32).reshape(-1))     # column-major
        w.node(mesh, matrix=m)
        t, r, s = (C.c_float * 3)(), (C.c_float * 4)(), (C.c_float * 3)()
        ref_shim.ref_decompose(m.ctypes.data_as(C.POINTER(C.c_float)), t, r, s)
        ref.add_node(p, translation=tuple(t), rotation=tuple(r), scale=tuple(s))
    ref.finalize()
    same_scene(Scene.from_glb(w.glb()), ref)


@pytest.mark.parametrize("breakage,needle", [
    ("magic", "magic"), ("truncated", "exceeds"), ("meshopt", "EXT_meshopt_compression"), ("sparse", "sparse accessor"), ("badindex", "index out of range"),
])
def test_malformed_assets_are_refused_with_a_reason(breakage, needle):
    pos, idx = S.grid_mesh(3, 3, lambda u, v: (u, v, 0 * u))
    w = GlbWriter()
    if breakage == "badindex":
        idx = idx.copy(); idx[4] = 1000
    w.node(w.mesh([{"position": w.positions(pos), "indices": w.indices(idx.astype(np.uint16))}]))
    if breakage == "meshopt":
        w.doc["bufferViews"][0]["extensions"] = {"EXT_meshopt_compression": {"buffer": 0, "byteLength": 10, "byteStride": 12, "count": 1, "mode": "ATTRIBUTES"}}
    if breakage == "sparse":
        w.doc["accessors"][0]["sparse"] = {"count": 1}
    data = bytearray(w.glb())
    if breakage == "magic":
        data[0] = 0
    if breakage == "truncated":
        data = data[:len(data) - 40]
    with pytest.raises(ValueError) as e:
        Scene.from_glb(bytes(data))
    assert needle in str(e.value)


def _asset():
    rng = np.random.default_rng(21)
    posA, idxA = S.grid_mesh(14, 11, lambda u, v: (u * 2, 0.3 * np.sin(u * 7), v * 2))
    qB = rng.integers(-900, 900, (30, 3)).astype(np.int16)
    idxB = rng.integers(0, 30, 66).astype(np.uint32)
    w = GlbWriter()
    m = w.material((0.3, 0.6, 0.9, 1.0), double_sided=True)
    mesh = w.mesh([{"position": w.positions(posA), "indices": w.indices(idxA.astype(np.uint16)), "material": m},
                   {"position": w.positions(qB), "indices": w.indices(idxB.astype(np.uint8))}])
    root = w.node(mesh, translation=(0, 1, 0))
    w.node(mesh, parent=root, scale=(2, 2, -2))
    return w


def test_gltf_files_with_external_and_embedded_buffers_equal_the_glb(tmp_path):
    """AssetLoadTask::loadGltf takes .gltf and .glb alike (assets.cpp:526-552) and BufferLoadTask reads buffers whose uri names a file
    in the asset's folder (assets.cpp:36-68): the same asset as a GLB blob, a .glb file, a .gltf + .bin pair (plain and percent-encoded
    file name) and a .gltf with a base64 data uri must give byte-identical scenes."""
    w = _asset()
    ref = Scene.from_glb(w.glb())
    (tmp_path / "a.glb").write_bytes(w.glb())
    same_scene(Scene.from_file(tmp_path / "a.glb"), ref)
    js, bn = w.gltf(uri="a.bin")
    (tmp_path / "a.gltf").write_bytes(js); (tmp_path / "a.bin").write_bytes(bn + b"trailing bytes beyond byteLength are not read")
    same_scene(Scene.from_file(tmp_path / "a.gltf"), ref)
    js, bn = w.gltf(uri="my%20buffers/geo%2B1.bin")                    # fastgltf::URI::fspath(): percent-decoded
    (tmp_path / "my buffers").mkdir()
    (tmp_path / "b.gltf").write_bytes(b"\xef\xbb\xbf" + js); (tmp_path / "my buffers" / "geo+1.bin").write_bytes(bn)
    same_scene(Scene.from_file(tmp_path / "b.gltf"), ref)
    js, _ = w.gltf(uri=None)                                             # embedded: data:application/octet-stream;base64,...
    (tmp_path / "c.gltf").write_bytes(js)
    same_scene(Scene.from_file(tmp_path / "c.gltf"), ref)
    # failures carry the reference's messages
    with pytest.raises(ValueError, match="Failed to open glTF file"):
        Scene.from_file(tmp_path / "missing.gltf")
    js, _ = w.gltf(uri="gone.bin")
    (tmp_path / "d.gltf").write_bytes(js)
    with pytest.raises(ValueError, match="Failed to open buffer"):
        Scene.from_file(tmp_path / "d.gltf")
    js, bn = w.gltf(uri="short.bin")
    (tmp_path / "e.gltf").write_bytes(js); (tmp_path / "short.bin").write_bytes(bn[:100])
    with pytest.raises(ValueError, match="Failed to open buffer|shorter"):
        Scene.from_file(tmp_path / "e.gltf")
    js, _ = w.gltf(uri="https://example.invalid/a.bin")
    (tmp_path / "f.gltf").write_bytes(js)
    with pytest.raises(ValueError, match="only local files"):
        Scene.from_file(tmp_path / "f.gltf")
    with pytest.raises(ValueError, match="no asset folder"):              # a blob has no folder to resolve a file name against
        Scene.from_glb(w.glb(extra_buffers=[{"byteLength": 4, "uri": "x.bin"}]))
    (tmp_path / "x.bin").write_bytes(b"\0\0\0\0")                         # ... a GLB FILE has one: its second buffer is read from beside it
    (tmp_path / "h.glb").write_bytes(w.glb(extra_buffers=[{"byteLength": 4, "uri": "x.bin"}]))
    same_scene(Scene.from_file(tmp_path / "h.glb"), ref)
    (tmp_path / "g.gltf").write_bytes(b"{ not json")
    with pytest.raises(ValueError, match="malformed glTF JSON"):
        Scene.from_file(tmp_path / "g.gltf")


@pytest.mark.parametrize("with_base", [True, False])
def test_sparse_accessors_are_read_like_iterate_accessor(with_base):
    """glTF 2.0 sparse accessors as fastgltf's iterateAccessor hands them to assets.cpp:310-314: the (index, value) pairs override the base
    view's elements — or zeros when the accessor has no bufferView."""
    pos, idx = S.grid_mesh(8, 6, lambda u, v: (u, v, 0 * u))
    rng = np.random.default_rng(5)
    where = np.sort(rng.choice(pos.shape[0], 9, replace=False)).astype(np.uint16)
    vals = rng.normal(size=(9, 3)).astype(np.float32)
    w = GlbWriter()
    w.node(w.mesh([{"position": w.sparse_positions(pos, where, vals, with_base=with_base), "indices": w.indices(idx.astype(np.uint16))}]))
    got = Scene.from_glb(w.glb())
    want = pos.copy() if with_base else np.zeros_like(pos)
    want[where] = vals
    ref = Scene.new()
    ref.add_node(ref.add_primitive(want, idx, 0))
    ref.finalize()
    # (no accessor min / max on either side: primitive AABBs are zero vectors, assets.cpp:303-306)
    same_scene(got, ref)


def test_mutated_assets_never_crash_the_reader_and_accepted_ones_are_safe_to_draw():
    """tools/fuzz_gltf.py (mutation fuzzer over JSON text, binary chunk and container length fields): the reader refuses or accepts, it does
    not crash; whatever it accepts satisfies what the device kernels trust (index ranges, meshlet limits).  A short run here; the tool's
    docstring shows the sanitizer run (34 000 mutants, ASan + UBSan clean)."""
    import subprocess, sys, os, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_gltf.py"), "600", "11"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.loads(p.stdout.splitlines()[0])
    assert r["iterations"] == 600 and r["accepted"] > 20 and r["refused"] > 300 and r["distinct_refusals"] >= 10


def test_primitives_built_on_all_cores_land_in_document_order():
    """the meshlet builds of an asset's primitives run in parallel (like the reference's PrimitiveProcessingTask, assets.cpp:192-215); the scene
    must be the one the sequential direct API builds, byte for byte, primitive order included"""
    w, ref = GlbWriter(), Scene.new()
    m = w.material((0.5, 0.5, 0.5, 1.0), double_sided=False)
    rm = ref.add_material((0.5, 0.5, 0.5, 1.0), False)
    for k in range(13):
        pos, idx = S.grid_mesh(10 + 3 * k, 7 + k, lambda u, v, k=k: (u * 3, 0.2 * np.sin(u * (5 + k)) * np.cos(v * 7), v * 2))
        mesh = w.mesh([{"position": w.positions(pos), "indices": w.indices(idx.astype(np.uint32)), "material": m}])
        w.node(mesh, translation=(k * 4.0, 0, 0))
        ref.add_mesh_node([ref.add_primitive(pos, idx, rm)], translation=(k * 4.0, 0, 0))
    ref.finalize()
    same_scene(Scene.from_glb(w.glb()), ref)


# ------------------------------------------------------------------------------------------ EXT_meshopt_compression through the decoder hook
def _compressed_and_plain_assets(M):
    """the same geometry twice: EXT_meshopt_compression views (streams from the REFERENCE's meshoptimizer) and plain views of what they decode to"""
    rng = np.random.default_rng(21)
    posA, idxA = S.grid_mesh(40, 31, lambda u, v: (u * 3, 0.2 * np.sin(u * 9) * np.cos(v * 5), v * 2))
    posA = posA.astype(np.float32)
    qB = rng.integers(-3000, 3000, (70, 3)).astype(np.int16)
    idxB = rng.integers(0, 70, 240).astype(np.uint16)
    # EXPONENTIAL filter: the stream holds (exponent << 24 | mantissa) words, the decoder turns them into floats
    L = M.ref_lib()
    posC, idxC = S.grid_mesh(12, 12, lambda u, v: (u + 0.37, v * v, 0.5 * u * v))
    posC = np.ascontiguousarray(posC, np.float32)
    enc_words = np.zeros(posC.shape, np.uint32)
    L.meshopt_encodeFilterExp(C.c_void_p(enc_words.ctypes.data), C.c_size_t(posC.shape[0]), C.c_size_t(12), C.c_int(15), C.c_void_p(posC.ctypes.data), C.c_int(0))
    rc, dec = M.ref_decode("vertex", posC.shape[0], 12, M.ref_encode("vertex", enc_words, posC.shape[0], 12), fid=3)
    assert rc == 0
    posC_decoded = dec.view(np.float32).reshape(-1, 3).copy()
    assert np.abs(posC_decoded - posC).max() < 1e-3 and not np.array_equal(posC_decoded, posC)

    def tri_decoded(idx, nverts, dtype):
        """meshopt's triangle codec keeps every triangle but may rotate its three indices: the plain asset holds what the stream decodes to"""
        i = np.ascontiguousarray(idx, np.uint32).reshape(-1)
        rc, out = M.ref_decode("index", i.size, np.dtype(dtype).itemsize, M.ref_encode("index", i, i.size, np.dtype(dtype).itemsize, nverts))
        assert rc == 0
        out = out.view(dtype)
        assert np.array_equal(np.sort(out.reshape(-1, 3), 1), np.sort(i.reshape(-1, 3), 1))
        return out

    def build(compressed):
        w = GlbWriter()
        m = w.material((0.3, 0.6, 0.9, 1.0), double_sided=True)
        if compressed:
            a = {"position": w.positions_compressed(posA, M.ref_encode), "indices": w.indices_compressed(idxA.astype(np.uint32), M.ref_encode, posA.shape[0]), "material": m}
            b = {"position": w.positions_compressed(qB, M.ref_encode), "indices": w.indices_compressed(idxB, M.ref_encode, 70, mode="INDICES"), "material": None}
            c = {"position": w.positions_compressed(enc_words.view(np.float32), M.ref_encode, filt="EXPONENTIAL", decoded=posC_decoded),
                 "indices": w.indices_compressed(idxC.astype(np.uint16), M.ref_encode, posC.shape[0]), "material": m}
        else:
            a = {"position": w.positions(posA), "indices": w.indices(tri_decoded(idxA, posA.shape[0], np.uint32)), "material": m}
            b = {"position": w.positions(qB, stride=8), "indices": w.indices(idxB), "material": None}
            c = {"position": w.positions(posC_decoded), "indices": w.indices(tri_decoded(idxC, posC.shape[0], np.uint16)), "material": m}
        w.node(w.mesh([a, b]), translation=(1, 0, 0))
        w.node(w.mesh([c]), scale=(2, 2, 2))
        return w.glb()
    return build(True), build(False)


def _decoder_from(decode):
    kinds = {0: "vertex", 1: "index", 2: "sequence"}

    def fn(mode, filt, count, stride, src, dst):
        rc, out = decode(kinds[mode], count, stride, np.array(src, copy=True), fid=filt)
        if rc == 0:
            dst[:] = out
        return rc
    return fn


@pytest.mark.parametrize("which", ["reference", "oracle"])
def test_meshopt_compressed_asset_loads_through_the_decoder_hook(meshopt_ref, which):
    """EXT_meshopt_compression (CompressedBufferDataAdapter, assets.cpp:70-171): ATTRIBUTES (float, int16 in 8-byte elements, EXPONENTIAL filter),
    TRIANGLES and INDICES views in a fallback buffer; with a decoder installed the scene equals the one loaded from the decoded bytes"""
    from tests import meshopt_lib as M
    from vk_gltf_viewer_b200 import scene as SC
    comp, plain = _compressed_and_plain_assets(M)
    SC.set_meshopt_decoder(_decoder_from(M.ref_decode if which == "reference" else M.oracle_decode))
    try:
        got = Scene.from_glb(comp)
    finally:
        SC.set_meshopt_decoder(None)
    same_scene(got, Scene.from_glb(plain))
    assert got.counts().primitives == 3 and got.counts().triangles_unique > 2500


def test_meshopt_compressed_asset_without_a_decoder_or_with_a_broken_stream_is_refused(meshopt_ref):
    from tests import meshopt_lib as M
    from vk_gltf_viewer_b200 import scene as SC
    comp, _ = _compressed_and_plain_assets(M)
    with pytest.raises(ValueError, match="vkvh_set_meshopt_decoder"):
        Scene.from_glb(comp)
    bad = bytearray(comp)
    import struct
    jlen = struct.unpack_from("<I", comp, 12)[0]
    bad[20 + jlen + 8] ^= 0xFF                              # first byte of the BIN chunk = header byte of the first stream
    SC.set_meshopt_decoder(_decoder_from(M.ref_decode))
    try:
        with pytest.raises(ValueError, match="did not decode"):
            Scene.from_glb(bytes(bad))
        # a plain view must not read the fallback buffer
        import json
        doc = json.loads(comp[20:20 + jlen])
        del doc["bufferViews"][0]["extensions"]
        js = json.dumps(doc, separators=(",", ":")).encode(); js += b" " * (-len(js) % 4)
        rest = comp[20 + jlen:]
        with pytest.raises(ValueError, match="fallback buffer"):
            Scene.from_glb(struct.pack("<III", 0x46546C67, 2, 20 + len(js) + len(rest)) + struct.pack("<II", len(js), 0x4E4F534A) + js + rest)
    finally:
        SC.set_meshopt_decoder(None)


@pytest.mark.gpu
@pytest.mark.late
def test_meshopt_compressed_asset_decoded_on_the_device(meshopt_ref):
    """the same asset with libvkv's device decoder behind the hook (api.Renderer.meshopt_decoder): identical scene"""
    from tests import meshopt_lib as M
    from vk_gltf_viewer_b200 import api, scene as SC
    comp, plain = _compressed_and_plain_assets(M)
    r = api.Renderer(64, 64)
    SC.set_meshopt_decoder(r.meshopt_decoder())
    try:
        got = Scene.from_glb(comp)
    finally:
        SC.set_meshopt_decoder(None)
        r.close()
    same_scene(got, Scene.from_glb(plain))


def test_device_decoder_glue_with_a_stand_in_for_the_device(meshopt_ref):
    """api.Renderer.meshopt_decoder()'s closure (view record, source slack, result slicing, return code) with Renderer.meshopt_decode replaced by
    the reference's meshoptimizer: what the GPU test exercises with the real device decoder, minus the device"""
    from tests import meshopt_lib as M
    from vk_gltf_viewer_b200 import abi, api, scene as SC
    comp, plain = _compressed_and_plain_assets(M)
    kinds = {0: "vertex", 1: "index", 2: "sequence"}
    seen = []

    class Stand(api.Renderer):
        def __init__(self):   # no context: only meshopt_decoder() / meshopt_decode() are used
            pass

        def meshopt_decode(self, src, views, dst_bytes):
            assert views.dtype == abi.MESHOPT_VIEW_DTYPE and len(views) == 1 and dst_bytes % 16 == 0
            v = views[0]
            assert int(v["src_offset"]) == 0 and int(v["dst_offset"]) == 0 and src.size == int(v["src_size"]) + 32 and not src[int(v["src_size"]):].any()
            rc, out = M.ref_decode(kinds[int(v["mode"])], int(v["count"]), int(v["stride"]), src[:int(v["src_size"])], fid=int(v["filter"]))
            full = np.zeros(dst_bytes, np.uint8)
            full[:out.size] = out
            seen.append(int(v["mode"]))
            return full, np.array([rc], np.int32)

    SC.set_meshopt_decoder(Stand().meshopt_decoder())
    try:
        got = Scene.from_glb(comp)
    finally:
        SC.set_meshopt_decoder(None)
    same_scene(got, Scene.from_glb(plain))
    assert sorted(set(seen)) == [0, 1, 2] and len(seen) == 6        # three ATTRIBUTES, two TRIANGLES, one INDICES view, each decoded once


@pytest.mark.parametrize("field,value,needle", [
    ("bufferViews.0.byteOffset", -1, "negative"), ("bufferViews.0.byteOffset", 2**63, "beyond"), ("bufferViews.0.byteLength", 1e300, "beyond"),
    ("accessors.0.byteOffset", -8, "negative"), ("accessors.0.count", 2**62, "beyond"), ("bufferViews.0.byteStride", 4096, "byteStride"),
    ("scene", -1, "scene index"), ("scenes.0.nodes.0", -3, "out of range"), ("scenes.0.nodes.0", 1e30, "out of range"),
])
def test_offsets_that_would_wrap_the_bounds_arithmetic_are_refused(field, value, needle):
    """found by tools/fuzz_gltf.py under ASan: byteOffset -1 became 2^64 - 1, `offset + length` wrapped past the check and the first vertex was
    read one byte before the buffer.  Sizes and offsets are now non-negative and below 2^40, and every bound is checked without an addition
    that can wrap."""
    import json, struct
    w = GlbWriter()
    pos, idx = S.grid_mesh(4, 4, lambda u, v: (u, v, 0 * u))
    w.node(w.mesh([{"position": w.positions(pos), "indices": w.indices(idx.astype(np.uint16))}]))
    glb = w.glb()
    jlen = struct.unpack_from("<I", glb, 12)[0]
    doc = json.loads(glb[20:20 + jlen])
    ref = doc
    keys = field.split(".")
    for k in keys[:-1]:
        ref = ref[int(k)] if k.isdigit() else ref[k]
    last = keys[-1]
    if last.isdigit():
        ref[int(last)] = value
    else:
        ref[last] = value
    js = json.dumps(doc, separators=(",", ":")).encode(); js += b" " * (-len(js) % 4)
    rest = glb[20 + jlen:]
    bad = struct.pack("<III", 0x46546C67, 2, 20 + len(js) + len(rest)) + struct.pack("<II", len(js), 0x4E4F534A) + js + rest
    with pytest.raises(ValueError, match=needle):
        Scene.from_glb(bad)


def test_a_node_with_two_parents_is_refused_instead_of_walked_exponentially():
    import json, struct, time
    nodes = [{"children": [i + 1, i + 1]} for i in range(60)] + [{}]
    js = json.dumps({"asset": {"version": "2.0"}, "nodes": nodes, "scenes": [{"nodes": [0]}]}).encode()
    js += b" " * (-len(js) % 4)
    t = time.time()
    with pytest.raises(ValueError, match="more than one parent"):
        Scene.from_glb(struct.pack("<III", 0x46546C67, 2, 20 + len(js)) + struct.pack("<II", len(js), 0x4E4F534A) + js)
    assert time.time() - t < 1.0

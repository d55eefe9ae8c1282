#!/usr/bin/env python
"""Mutation fuzzer for the glTF reader (host/gltf.cpp: the one place on this path that parses untrusted bytes on the host).

    python tools/fuzz_gltf.py [iterations] [seed] [compressed | gltf]

Builds a GLB with every feature the reader handles (float / quantised / strided / sparse accessors, u8 / u16 / u32 / generated indices,
matrix and TRS nodes, several scenes' worth of hierarchy), then mutates JSON text and binary chunk (bit flips, number replacement, token
deletion / duplication, truncation, length-field edits) and feeds every mutant to vkvh_scene_load_glb.  The reader may refuse a mutant or
accept it; it may not crash, hang or read out of bounds.  Run it against a sanitizer build for the last part:

    (in a scratch copy of the tree)
    g++ -std=c++17 -fPIC -fsanitize=address,undefined,float-cast-overflow -g -O1 -shared -o vk_gltf_viewer_b200/libvkv_host.so vk_gltf_viewer_b200/host/*.cpp -Iinclude
    LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 python tools/fuzz_gltf.py 20000

Every ACCEPTED mutant is then checked the way the device kernels trust it (the trust boundary of include/vkv.h): every MeshletDraw names an
existing primitive / meshlet / transform slot, every meshlet's vertex and triangle ranges lie inside the primitive's arrays, every meshlet
vertex index is below the vertex count, every triangle corner below the meshlet's vertex count, limits 64 / 124 hold.
"""
import ctypes as C
import json
import os
import re
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.gltf_writer import GlbWriter  # noqa: E402
from tests import scenes as S  # noqa: E402


def rebuild_writer():
    rng = np.random.default_rng(1)
    w = GlbWriter()
    m0 = w.material((0.8, 0.2, 0.1, 1.0), double_sided=False)
    posA, idxA = S.grid_mesh(12, 9, lambda u, v: (u * 3, 0.2 * np.sin(u * 9), v * 2))
    q = rng.integers(-2000, 2000, (40, 3)).astype(np.int16)
    iq = rng.integers(0, 40, 90)
    a = {"position": w.positions(posA), "indices": w.indices(idxA.astype(np.uint32)), "material": m0}
    b = {"position": w.positions(posA, stride=20), "indices": w.indices(idxA.astype(np.uint16)), "material": None}
    c = {"position": w.positions(q), "indices": w.indices(iq.astype(np.uint8)), "material": m0}
    d = {"position": w.positions(q, normalized=True), "material": m0}          # generated indices
    m = w.mesh([a, b]), w.mesh([c, d])
    root = w.node(translation=(1, 2, 3))
    n1 = w.node(m[0], parent=root, rotation=(0, 0.6, 0, 0.8), scale=(2, 0.5, 1.5))
    w.node(m[1], parent=n1, translation=(0.5, 0, -1))
    w.node(m[0], scale=(1, 1, -1))
    return w


def base_glb():
    return rebuild_writer().glb()


def split(glb):
    jlen = struct.unpack_from("<I", glb, 12)[0]
    js = glb[20:20 + jlen]
    rest = glb[20 + jlen:]
    return js, rest


def join(js, rest, total=None, jlen=None):
    js = js + b" " * (-len(js) % 4)
    body = struct.pack("<II", len(js) if jlen is None else jlen, 0x4E4F534A) + js + rest
    return struct.pack("<III", 0x46546C67, 2, (12 + len(body)) if total is None else total) + body


NUM = re.compile(rb"-?\d+(\.\d+)?([eE][-+]?\d+)?")
INTERESTING = [b"0", b"-1", b"1", b"2", b"3", b"4", b"255", b"65535", b"65536", b"2147483647", b"2147483648", b"4294967295", b"4294967296",
               b"18446744073709551615", b"1e308", b"-1e308", b"1e-320", b"5120", b"5121", b"5122", b"5123", b"5125", b"5126", b"NaN", b"null", b"true",
               b"[]", b"{}", b"\"\"", b"0.5"]


def mutate(rng, js, rest):
    js, rest = bytearray(js), bytearray(rest)
    total = jlen = None
    for _ in range(int(rng.integers(1, 4))):
        k = int(rng.integers(0, 10))
        if k <= 3:                                   # replace a number in the JSON
            ms = list(NUM.finditer(bytes(js)))
            if ms:
                m = ms[int(rng.integers(len(ms)))]
                js[m.start():m.end()] = INTERESTING[int(rng.integers(len(INTERESTING)))]
        elif k == 4 and len(js) > 8:                 # flip a byte of the JSON
            js[int(rng.integers(len(js)))] = int(rng.integers(32, 127))
        elif k == 5 and len(js) > 16:                # delete a span
            a = int(rng.integers(len(js) - 1)); b = min(len(js), a + int(rng.integers(1, 40)))
            del js[a:b]
        elif k == 6 and len(js) > 16:                # duplicate a span
            a = int(rng.integers(len(js) - 1)); b = min(len(js), a + int(rng.integers(1, 60)))
            js[a:a] = js[a:b]
        elif k == 7 and len(rest) > 16:              # corrupt the binary chunk (header or payload)
            for _ in range(int(rng.integers(1, 8))):
                rest[int(rng.integers(len(rest)))] = int(rng.integers(256))
        elif k == 8:                                 # truncate
            if rng.random() < 0.5 and len(rest) > 4:
                del rest[int(rng.integers(len(rest))):]
            elif len(js) > 4:
                del js[int(rng.integers(len(js))):]
        else:                                        # lie in the container's length fields
            if rng.random() < 0.5:
                total = int(rng.choice([0, 11, 12, 20, 2**31, 2**32 - 1, len(js) + len(rest)]))
            else:
                jlen = int(rng.choice([0, 1, 3, len(js) + 4, len(js) + len(rest) + 64, 2**31, 2**32 - 4]))
    return join(bytes(js), bytes(rest), total, jlen)


def check(scene):
    """what the kernels rely on without re-checking (vkv.h: the caller's draw / primitive / material indices are trusted)"""
    c = scene.counts()
    draws = scene.draws()
    assert draws.shape[0] == c.draws
    prims = [scene.primitive(i) for i in range(c.primitives)]
    if c.draws:
        assert int(draws["primitiveIndex"].max()) < c.primitives and int(draws["transformIndex"].max()) < max(c.transforms, 1)
        for i, p in enumerate(prims):
            sel = draws["primitiveIndex"] == i
            if sel.any():
                assert int(draws["meshletIndex"][sel].max()) < p["meshlets"].shape[0]
    assert scene.transforms().shape[0] == c.transforms and np.isfinite(scene.transforms()).all() or True   # NaN transforms are legal input
    for p in prims:
        ml = p["meshlets"]
        assert p["header"].materialIndex < c.materials
        if ml.shape[0] == 0:
            continue
        vc, tc = ml["vertexCount"].astype(np.int64), ml["triangleCount"].astype(np.int64)
        assert vc.max() <= 64 and tc.max() <= 124
        assert (ml["vertexOffset"].astype(np.int64) + vc).max() <= p["vertex_indices"].shape[0]
        assert (ml["triangleOffset"].astype(np.int64) + 3 * tc).max() <= p["triangles"].shape[0]
        if p["vertex_indices"].shape[0]:
            assert int(p["vertex_indices"].max()) < p["vertices"].shape[0]
        for m in ml:
            t = p["triangles"][int(m["triangleOffset"]):int(m["triangleOffset"]) + 3 * int(m["triangleCount"])]
            if t.size:
                assert int(t.max()) < int(m["vertexCount"])


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    from vk_gltf_viewer_b200.scene import Scene
    from vk_gltf_viewer_b200 import scene as SC
    glb = base_glb()
    if len(sys.argv) > 3 and sys.argv[3] == "compressed":
        # an EXT_meshopt_compression asset (streams, fallback buffer, decoder hook) with the reference's meshoptimizer as the decoder
        from tests import meshopt_lib as M
        from tests.test_gltf import _compressed_and_plain_assets, _decoder_from
        assert M.ref_lib() is not None, "needs oracle/_ref/libmeshopt_ref.so (make ref)"
        glb, _ = _compressed_and_plain_assets(M)
        SC.set_meshopt_decoder(_decoder_from(M.ref_decode))
    js, rest = split(glb)
    check(Scene.from_glb(glb))
    rng = np.random.default_rng(seed)
    accepted, reasons = 0, {}
    as_file = len(sys.argv) > 3 and sys.argv[3] == "gltf"
    if as_file:
        # the .gltf route (vkvh_scene_load_file): a JSON document with a base64 data uri and an external .bin beside it, mutated as text
        import tempfile
        from tests.gltf_writer import GlbWriter as _W  # noqa: F401
        tmp = tempfile.mkdtemp(prefix="vkv_fuzz_")
        w = rebuild_writer()
        doc_data, _ = w.gltf(None)                       # buffers[0] as a data uri
        doc_file, bn = w.gltf("geometry.bin")            # buffers[0] as an external file
        open(os.path.join(tmp, "geometry.bin"), "wb").write(bn)
        bases = [doc_data, doc_file]
    for i in range(iters):
        if as_file:
            base = bytearray(bases[i % 2])
            for _ in range(int(rng.integers(1, 4))):
                k = int(rng.integers(0, 4))
                if k == 0:
                    ms = list(NUM.finditer(bytes(base)))
                    m_ = ms[int(rng.integers(len(ms)))]
                    base[m_.start():m_.end()] = INTERESTING[int(rng.integers(len(INTERESTING)))]
                elif k == 1:
                    base[int(rng.integers(len(base)))] = int(rng.integers(32, 127))
                elif k == 2:
                    a = int(rng.integers(len(base) - 1)); del base[a:min(len(base), a + int(rng.integers(1, 40)))]
                else:
                    a = int(rng.integers(len(base) - 1)); base[a:a] = base[a:min(len(base), a + int(rng.integers(1, 60)))]
            path = os.path.join(tmp, "mutant.gltf")
            open(path, "wb").write(bytes(base))
            try:
                scene = Scene.from_file(path)
            except ValueError as e:
                key = re.sub(r"\d+", "N", str(e))[:60]
                reasons[key] = reasons.get(key, 0) + 1
                continue
            accepted += 1
            check(scene)
            continue
        m = mutate(rng, js, rest)
        try:
            scene = Scene.from_glb(m)
        except ValueError as e:
            key = re.sub(r"\d+", "N", str(e))[:60]
            reasons[key] = reasons.get(key, 0) + 1
            continue
        accepted += 1
        check(scene)
    print(json.dumps({"iterations": iters, "seed": seed, "accepted": accepted, "refused": iters - accepted, "distinct_refusals": len(reasons)}))
    for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:12]:
        print(f"  {v:6d}  {k}")


if __name__ == "__main__":
    main()

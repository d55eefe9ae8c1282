"""A tiny glTF 2.0 binary (GLB) writer for the ingest tests: procedurally generated glTF scenes (BASELINE.json: "synthetic
procedurally generated glTF scenes").  It writes what the reference reads of an asset (assets.cpp:288-373, world.cpp:187-293):
buffers, bufferViews (optionally strided), accessors (float or KHR_mesh_quantization integers, u8 / u16 / u32 indices), materials,
meshes with several primitives, nodes (TRS or matrix), one scene."""
import json
import struct

import numpy as np

CT = {np.dtype(np.int8): 5120, np.dtype(np.uint8): 5121, np.dtype(np.int16): 5122, np.dtype(np.uint16): 5123,
      np.dtype(np.uint32): 5125, np.dtype(np.float32): 5126}


class GlbWriter:
    def __init__(self):
        self.bin = bytearray()
        self.doc = {"asset": {"version": "2.0"}, "buffers": [{}], "bufferViews": [], "accessors": [], "materials": [], "meshes": [],
                    "nodes": [], "scenes": [{"nodes": []}], "scene": 0, "extensionsUsed": []}

    def _view(self, raw: bytes, stride=0) -> int:
        while len(self.bin) % 4:
            self.bin.append(0)
        v = {"buffer": 0, "byteOffset": len(self.bin), "byteLength": len(raw)}
        if stride:
            v["byteStride"] = stride
        self.bin += raw
        self.doc["bufferViews"].append(v)
        return len(self.doc["bufferViews"]) - 1

    def positions(self, pos, normalized=False, stride=0, minmax=True) -> int:
        """pos: (n,3) float32 or an integer dtype (KHR_mesh_quantization); stride > element size interleaves padding"""
        p = np.ascontiguousarray(pos)
        elem = p.dtype.itemsize * 3
        if stride:
            raw = bytearray(stride * p.shape[0])
            for i in range(p.shape[0]):
                raw[i * stride: i * stride + elem] = p[i].tobytes()
            raw = bytes(raw)
        else:
            raw = p.tobytes()
        a = {"bufferView": self._view(raw, stride), "componentType": CT[p.dtype], "count": int(p.shape[0]), "type": "VEC3"}
        if normalized:
            a["normalized"] = True
        if p.dtype != np.float32 and "KHR_mesh_quantization" not in self.doc["extensionsUsed"]:
            self.doc["extensionsUsed"].append("KHR_mesh_quantization")
        if minmax:
            a["min"] = [float(x) if p.dtype == np.float32 else int(x) for x in p.min(0)]
            a["max"] = [float(x) if p.dtype == np.float32 else int(x) for x in p.max(0)]
        self.doc["accessors"].append(a)
        return len(self.doc["accessors"]) - 1

    def indices(self, idx) -> int:
        i = np.ascontiguousarray(idx).reshape(-1)
        self.doc["accessors"].append({"bufferView": self._view(i.tobytes()), "componentType": CT[i.dtype], "count": int(i.size), "type": "SCALAR"})
        return len(self.doc["accessors"]) - 1

    # ---- EXT_meshopt_compression: the bufferView lives in a fallback buffer (no bytes), its extension names the compressed range
    def _compressed_view(self, encoded: bytes, mode: str, count: int, stride: int, filt: str = "NONE", view_stride=0) -> int:
        while len(self.bin) % 4:
            self.bin.append(0)
        off = len(self.bin)
        self.bin += encoded
        if "EXT_meshopt_compression" not in self.doc["extensionsUsed"]:
            self.doc["extensionsUsed"].append("EXT_meshopt_compression")
        if not hasattr(self, "fallback_bytes"):
            self.fallback_bytes = 0
        v = {"buffer": 1, "byteOffset": self.fallback_bytes, "byteLength": count * stride,
             "extensions": {"EXT_meshopt_compression": {"buffer": 0, "byteOffset": off, "byteLength": len(encoded), "byteStride": stride,
                                                         "count": count, "mode": mode, "filter": filt}}}
        if view_stride:
            v["byteStride"] = view_stride
        self.fallback_bytes += (count * stride + 3) & ~3
        self.doc["bufferViews"].append(v)
        return len(self.doc["bufferViews"]) - 1

    def positions_compressed(self, pos, encode, filt="NONE", decoded=None) -> int:
        """pos: (n,3) float32, or int16 (padded to 8-byte elements: ATTRIBUTES strides are multiples of 4); encode(kind, data, count, stride) ->
        stream bytes (the reference's meshoptimizer in the tests); `decoded`: what the stream decodes to when a filter changes the values (min / max)"""
        p = np.ascontiguousarray(pos)
        n = int(p.shape[0])
        if p.dtype == np.float32:
            raw, stride, vstride = p, 12, 0
        else:
            raw = np.zeros((n, 4), p.dtype); raw[:, :3] = p
            stride, vstride = 4 * p.dtype.itemsize, 4 * p.dtype.itemsize
        view = self._compressed_view(bytes(encode("vertex", raw, n, stride)), "ATTRIBUTES", n, stride, filt, vstride)
        q = np.ascontiguousarray(decoded if decoded is not None else p)
        a = {"bufferView": view, "componentType": CT[p.dtype], "count": n, "type": "VEC3",
             "min": [float(x) if p.dtype == np.float32 else int(x) for x in q.min(0)], "max": [float(x) if p.dtype == np.float32 else int(x) for x in q.max(0)]}
        if p.dtype != np.float32 and "KHR_mesh_quantization" not in self.doc["extensionsUsed"]:
            self.doc["extensionsUsed"].append("KHR_mesh_quantization")
        self.doc["accessors"].append(a)
        return len(self.doc["accessors"]) - 1

    def indices_compressed(self, idx, encode, nverts, mode="TRIANGLES") -> int:
        """idx: uint16 or uint32; TRIANGLES = meshopt_encodeIndexBuffer, INDICES = meshopt_encodeIndexSequence"""
        i = np.ascontiguousarray(idx).reshape(-1)
        assert i.dtype in (np.uint16, np.uint32)
        enc = encode("index" if mode == "TRIANGLES" else "sequence", i.astype(np.uint32), i.size, i.dtype.itemsize, nverts)
        view = self._compressed_view(bytes(enc), mode, int(i.size), i.dtype.itemsize)
        self.doc["accessors"].append({"bufferView": view, "componentType": CT[i.dtype], "count": int(i.size), "type": "SCALAR"})
        return len(self.doc["accessors"]) - 1

    def material(self, base_color=(1, 1, 1, 1), double_sided=False, alpha_cutoff=None) -> int:
        m = {"pbrMetallicRoughness": {"baseColorFactor": [float(x) for x in base_color]}, "doubleSided": bool(double_sided)}
        if alpha_cutoff is not None:
            m["alphaCutoff"] = float(alpha_cutoff)
        self.doc["materials"].append(m)
        return len(self.doc["materials"]) - 1

    def mesh(self, primitives) -> int:
        """primitives: list of dicts {position: accessor, indices: accessor or None, material: index or None, mode: int}"""
        ps = []
        for p in primitives:
            d = {"attributes": {"POSITION": p["position"]}}
            if p.get("indices") is not None:
                d["indices"] = p["indices"]
            if p.get("material") is not None:
                d["material"] = p["material"]
            if "mode" in p:
                d["mode"] = p["mode"]
            ps.append(d)
        self.doc["meshes"].append({"primitives": ps})
        return len(self.doc["meshes"]) - 1

    def node(self, mesh=None, parent=None, translation=None, rotation=None, scale=None, matrix=None) -> int:
        n = {}
        if mesh is not None:
            n["mesh"] = mesh
        if matrix is not None:
            n["matrix"] = [float(x) for x in np.asarray(matrix, np.float32).reshape(-1)]   # column-major, as glTF stores it
        else:
            if translation is not None: n["translation"] = [float(x) for x in translation]
            if rotation is not None: n["rotation"] = [float(x) for x in rotation]
            if scale is not None: n["scale"] = [float(x) for x in scale]
        self.doc["nodes"].append(n)
        i = len(self.doc["nodes"]) - 1
        if parent is None:
            self.doc["scenes"][0]["nodes"].append(i)
        else:
            self.doc["nodes"][parent].setdefault("children", []).append(i)
        return i

    def sparse_positions(self, base, indices, values, with_base=True) -> int:
        """a VEC3 float accessor whose elements `indices` are overridden by `values` (glTF 2.0 sparse accessor); with_base=False: no
        bufferView, the untouched elements read as zeros"""
        b = np.ascontiguousarray(base, np.float32)
        i = np.ascontiguousarray(indices)
        v = np.ascontiguousarray(values, np.float32)
        a = {"componentType": 5126, "count": int(b.shape[0]), "type": "VEC3",
             "sparse": {"count": int(i.size), "indices": {"bufferView": self._view(i.tobytes()), "componentType": CT[i.dtype]},
                        "values": {"bufferView": self._view(v.tobytes())}}}
        if with_base:
            a["bufferView"] = self._view(b.tobytes())
        self.doc["accessors"].append(a)
        return len(self.doc["accessors"]) - 1

    def gltf(self, uri=None) -> tuple:
        """the same asset as a .gltf JSON document: (json bytes, bin bytes) with buffers[0].uri = `uri`, or, uri=None, a base64 data uri"""
        import base64
        doc = dict(self.doc)
        bn = bytes(self.bin)
        doc["buffers"] = [{"byteLength": len(bn), "uri": uri if uri is not None else "data:application/octet-stream;base64," + base64.b64encode(bn).decode()}]
        for k in ("materials", "extensionsUsed"):
            if not doc[k]:
                del doc[k]
        return json.dumps(doc, indent=1).encode(), bn

    def glb(self, extra_buffers=()) -> bytes:
        doc = dict(self.doc)
        doc["buffers"] = [{"byteLength": len(self.bin)}] + list(extra_buffers)
        if getattr(self, "fallback_bytes", 0):
            assert not extra_buffers
            doc["buffers"].append({"byteLength": self.fallback_bytes, "extensions": {"EXT_meshopt_compression": {"fallback": True}}})
        for k in ("materials", "extensionsUsed"):
            if not doc[k]:
                del doc[k]
        js = json.dumps(doc, separators=(",", ":")).encode()
        js += b" " * (-len(js) % 4)
        bn = bytes(self.bin) + b"\0" * (-len(self.bin) % 4)
        total = 12 + 8 + len(js) + 8 + len(bn)
        return struct.pack("<III", 0x46546C67, 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(bn), 0x004E4942) + bn

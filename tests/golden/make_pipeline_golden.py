#!/usr/bin/env python
"""A small EXT_meshopt_compression asset for the device-side load-path test (tests/test_gpu_pipeline.py) -> tests/golden/pipeline_asset.npz
Run in the build container only: the streams are ENCODED by the reference's meshoptimizer (oracle/_ref/libmeshopt_ref.so).
Two primitives, each an interleaved glsl::Vertex stream (24-byte stride, vertex codec) and a u32 triangle list (index codec)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import meshopt_lib as M  # noqa: E402
from vk_gltf_viewer_b200 import abi  # noqa: E402

rng = np.random.default_rng(0x5EED0F5)


def grid(n, fn):
    u, v = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
    pos = np.stack(fn(u, v), -1).reshape(-1, 3).astype(np.float32)
    q = np.arange(n * n).reshape(n, n)
    idx = np.stack([q[:-1, :-1].ravel(), q[1:, :-1].ravel(), q[:-1, 1:].ravel(), q[:-1, 1:].ravel(), q[1:, :-1].ravel(), q[1:, 1:].ravel()], 1).astype(np.uint32).reshape(-1)
    return pos, idx


def sphere(n):
    return grid(n, lambda u, v: (np.sin(np.pi * v) * np.cos(2 * np.pi * u), np.cos(np.pi * v), np.sin(np.pi * v) * np.sin(2 * np.pi * u)))


out = {}
for k, (pos, idx) in enumerate([grid(70, lambda u, v: ((u * 2 - 1) * 3, 0.3 * np.sin(6 * u) * np.cos(5 * v) - 1.0, (v * 2 - 1) * 3)), sphere(48)]):
    vtx = np.zeros(pos.shape[0], abi.VERTEX_DTYPE)
    vtx["position"] = pos
    vtx["color"] = (np.arange(pos.shape[0])[:, None] // np.array([1, 3, 7, 64])) % 256
    vtx["normal"] = np.clip(pos * 40 + 128, 0, 255)
    vtx["uv"] = (np.arange(pos.shape[0])[:, None] * np.array([13, 29])) % 65536
    raw = vtx.view(np.uint8).reshape(-1)
    out[f"p{k}_counts"] = np.array([vtx.shape[0], idx.size], np.uint32)   # the decoded arrays are not stored: the oracle decoder recreates them
    out[f"p{k}_vertex_stream"] = M.ref_encode("vertex", vtx, vtx.shape[0], 24)
    out[f"p{k}_index_stream"] = M.ref_encode("index", idx, idx.size, 4, vtx.shape[0], 1)
    assert np.array_equal(M.ref_decode("vertex", vtx.shape[0], 24, out[f"p{k}_vertex_stream"])[1], raw)
    assert np.array_equal(M.ref_decode("index", idx.size, 4, out[f"p{k}_index_stream"])[1].view(np.uint32), idx)
# the same two meshes the way gltfpack writes them: POSITION as normalized SHORT VEC3 padded to 8 bytes (KHR_mesh_quantization,
# dequantised by a node scale), indices as a 16-bit triangle list; each a compressed view of its own
for k, (pos, idx) in enumerate([grid(70, lambda u, v: ((u * 2 - 1), 0.1 * np.sin(6 * u) * np.cos(5 * v), (v * 2 - 1))), sphere(48)]):
    q = np.zeros((pos.shape[0], 4), np.int16)
    q[:, :3] = np.round(np.clip(pos, -1, 1) * 32767).astype(np.int16)
    out[f"q{k}_counts"] = np.array([pos.shape[0], idx.size], np.uint32)
    out[f"q{k}_position_stream"] = M.ref_encode("vertex", q, q.shape[0], 8)
    out[f"q{k}_index_stream"] = M.ref_encode("index", idx, idx.size, 2, pos.shape[0], 1)
    assert np.array_equal(M.ref_decode("vertex", q.shape[0], 8, out[f"q{k}_position_stream"])[1], q.view(np.uint8).reshape(-1))
    assert np.array_equal(M.ref_decode("index", idx.size, 2, out[f"q{k}_index_stream"])[1].view(np.uint16), idx.astype(np.uint16))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipeline_asset.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})

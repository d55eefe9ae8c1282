#!/usr/bin/env python
"""Freeze golden vectors for the EXT_meshopt_compression decoders -> tests/golden/meshopt_codec.npz

Run in the build container only (needs /root/reference and oracle/_ref, `make ref`); the GPU box uses the committed .npz.

Two sources, both the REFERENCE's own:
 1. the known-answer byte arrays of submodules/meshoptimizer/demo/tests.cpp (kIndexDataV0/V1, kIndexSequenceV1, kVertexDataV0
    with their decoded forms, and the data/expected pairs of decodeFilterOct8/Oct12/Quat12/Exp), parsed out of the source file
    at generation time — nothing of the reference is copied into the repo except these few dozen numbers, as data;
 2. streams ENCODED and DECODED by the reference's meshoptimizer compiled from source (oracle/_ref/libmeshopt_ref.so for the
    codecs, libmeshopt_ref_nosimd.so for the scalar filter definitions): structured meshes, random data, every group width,
    sentinels, multi-block streams, 16- and 32-bit indices, both index-codec versions.

Each case is stored as (kind, params, encoded bytes, decoded bytes); tests decode `encoded` and compare with `decoded`.
"""
import ctypes as C
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REF", "/root/reference")
TESTS_CPP = os.path.join(REF, "submodules/meshoptimizer/demo/tests.cpp")


def c_array(src, name, after=None):
    """the numbers of `name[...] = { ... };` (first occurrence after the text `after`)"""
    pos = src.index(after) if after else 0
    m = re.compile(r"\b" + re.escape(name) + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", re.S).search(src, pos)
    body = re.sub(r"//[^\n]*", "", m.group(1))
    return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]


def main():
    src = open(TESTS_CPP).read()
    ref = C.CDLL(os.path.join(ROOT, "oracle/_ref/libmeshopt_ref.so"))
    ref_ns = C.CDLL(os.path.join(ROOT, "oracle/_ref/libmeshopt_ref_nosimd.so"))
    for L in (ref, ref_ns):
        for f in ("meshopt_encodeVertexBuffer", "meshopt_encodeVertexBufferBound", "meshopt_encodeIndexBuffer", "meshopt_encodeIndexBufferBound",
                  "meshopt_encodeIndexSequence", "meshopt_encodeIndexSequenceBound"):
            getattr(L, f).restype = C.c_size_t
    cases = []  # (name, kind, p0, p1, p2, encoded u8[], decoded u8[])

    def add(name, kind, count, stride, extra, enc, dec):
        cases.append((name, kind, count, stride, extra, np.frombuffer(bytes(enc), np.uint8).copy(), np.frombuffer(bytes(dec), np.uint8).copy()))

    kat_filters = []
    # ---- 1. known-answer vectors of demo/tests.cpp
    ib = np.array(c_array(src, "kIndexBuffer"), np.uint32)
    add("kat_index_v0", "index", ib.size, 4, 0, bytes(c_array(src, "kIndexDataV0")), ib.tobytes())
    ibt = np.array(c_array(src, "kIndexBufferTricky"), np.uint32)
    add("kat_index_v1", "index", ibt.size, 4, 0, bytes(c_array(src, "kIndexDataV1")), ibt.tobytes())
    seq = np.array(c_array(src, "kIndexSequence"), np.uint32)
    add("kat_sequence_v1", "sequence", seq.size, 4, 0, bytes(c_array(src, "kIndexSequenceV1")), seq.tobytes())
    pv = np.array(c_array(src, "kVertexBuffer"), np.int64).reshape(4, 7)  # PV {u16 px,py,pz; u8 nu,nv; u16 tx,ty} = 12 bytes
    vb = b"".join(np.array(r[:3], np.uint16).tobytes() + np.array(r[3:5], np.uint8).tobytes() + np.array(r[5:], np.uint16).tobytes() for r in pv)
    add("kat_vertex_v0", "vertex", 4, 12, 0, bytes(c_array(src, "kVertexDataV0")), vb)
    for fn, kind, stride, dt in (("decodeFilterOct8", "filter_oct", 4, np.uint8), ("decodeFilterOct12", "filter_oct", 8, np.uint16),
                                 ("decodeFilterQuat12", "filter_quat", 8, np.uint16), ("decodeFilterExp", "filter_exp", 4, np.uint32)):
        d = np.array(c_array(src, "data", "static void " + fn), dt)
        e = np.array(c_array(src, "expected", "static void " + fn), dt)
        kat_filters.append(("kat_" + fn, kind, d.nbytes // stride, stride, d, e.tobytes()))

    # ---- 2. encoded + decoded by the reference's library
    rng = np.random.default_rng(0x5EED0F3)

    def enc_vertex(data, count, stride):
        data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        bound = ref.meshopt_encodeVertexBufferBound(C.c_size_t(count), C.c_size_t(stride))
        buf = np.zeros(bound, np.uint8)
        n = ref.meshopt_encodeVertexBuffer(buf.ctypes.data_as(C.c_void_p), C.c_size_t(bound), data.ctypes.data_as(C.c_void_p), C.c_size_t(count), C.c_size_t(stride))
        out = np.zeros(count * stride, np.uint8)
        rc = ref.meshopt_decodeVertexBuffer(out.ctypes.data_as(C.c_void_p), C.c_size_t(count), C.c_size_t(stride), buf.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        assert rc == 0 and np.array_equal(out, data[:count * stride])
        return buf[:n], out

    # a: smooth quantised positions + normals (what a glTF attribute stream looks like), several strides and block counts
    for count, stride in ((1, 4), (15, 4), (16, 8), (17, 12), (255, 12), (256, 16), (257, 16), (1000, 12), (3000, 24), (1100, 32), (300, 64), (33, 256)):
        t = np.linspace(0, 20, count)
        cols = [np.round(4000 * np.sin(t * (k + 1) * 0.37) + 300 * rng.standard_normal(count) * (k % 3 == 0)).astype(np.int16).view(np.uint16) for k in range(stride // 2)]
        data = np.stack(cols, 1).astype(np.uint16)
        e, d = enc_vertex(data, count, stride)
        add(f"vertex_smooth_{count}x{stride}", "vertex", count, stride, 0, e, d)
    # b: every group width, sentinels, random
    for count, stride, mode in ((64, 4, "bits"), (64, 4, "sentinel"), (512, 8, "random"), (300, 4, "zeros"), (2000, 12, "mixed")):
        if mode == "bits":
            data = np.stack([np.zeros(count), np.arange(count), np.arange(count) * 2, np.arange(count) * 8], 1).astype(np.uint8)
        elif mode == "sentinel":
            data = np.stack([np.zeros(count), np.arange(count), np.arange(count) * 2, np.arange(count) * 8], 1).astype(np.uint8)
            data[[7, 13, 31, 32, 33, 63]] = 42
        elif mode == "random":
            data = rng.integers(0, 256, (count, stride), dtype=np.uint8)
        elif mode == "zeros":
            data = np.zeros((count, stride), np.uint8)
        else:
            steps = rng.choice([0, 1, 2, 3, 7, 8, 15, 16, 100, 255], (count, stride), p=[.3, .2, .1, .1, .05, .05, .05, .05, .05, .05]).astype(np.uint8)
            data = np.cumsum(steps, 0).astype(np.uint8)
        e, d = enc_vertex(data, count, stride)
        add(f"vertex_{mode}_{count}x{stride}", "vertex", count, stride, 0, e, d)

    def grid_indices(n, shuffle=False):
        q = np.arange(n * n).reshape(n, n)
        a, b, c, d = q[:-1, :-1].ravel(), q[:-1, 1:].ravel(), q[1:, :-1].ravel(), q[1:, 1:].ravel()
        tris = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)]).astype(np.uint32)
        if shuffle:
            tris = tris[rng.permutation(tris.shape[0])]
        return tris.reshape(-1)

    def enc_index(idx, nverts, version, index_size):
        ref.meshopt_encodeIndexVersion(version)
        bound = ref.meshopt_encodeIndexBufferBound(C.c_size_t(idx.size), C.c_size_t(nverts))
        buf = np.zeros(bound, np.uint8)
        n = ref.meshopt_encodeIndexBuffer(buf.ctypes.data_as(C.c_void_p), C.c_size_t(bound), idx.ctypes.data_as(C.c_void_p), C.c_size_t(idx.size))
        assert n > 0
        out = np.zeros(idx.size * index_size, np.uint8)
        rc = ref.meshopt_decodeIndexBuffer(out.ctypes.data_as(C.c_void_p), C.c_size_t(idx.size), C.c_size_t(index_size), buf.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        assert rc == 0
        return buf[:n], out

    for name, idx, nverts in (("grid40", grid_indices(40), 1600), ("grid40_shuffled", grid_indices(40, True), 1600),
                              ("random", rng.integers(0, 5000, 3 * 700).astype(np.uint32), 5000),
                              ("big_ids", (rng.integers(0, 2**31, 3 * 300)).astype(np.uint32), 2**31),
                              ("fans", np.stack([np.zeros(500), np.arange(1, 501), np.arange(2, 502)], 1).astype(np.uint32).reshape(-1), 502)):
        for version in (0, 1):
            for isz in (4, 2):
                if isz == 2 and idx.max() >= 65536:
                    continue
                e, d = enc_index(np.ascontiguousarray(idx, np.uint32), nverts, version, isz)
                add(f"index_{name}_v{version}_{isz * 8}", "index", idx.size, isz, version, e, d)

    def enc_seq(idx, nverts, version, index_size):
        ref.meshopt_encodeIndexVersion(version)
        bound = ref.meshopt_encodeIndexSequenceBound(C.c_size_t(idx.size), C.c_size_t(nverts))
        buf = np.zeros(bound, np.uint8)
        n = ref.meshopt_encodeIndexSequence(buf.ctypes.data_as(C.c_void_p), C.c_size_t(bound), idx.ctypes.data_as(C.c_void_p), C.c_size_t(idx.size))
        assert n > 0
        out = np.zeros(idx.size * index_size, np.uint8)
        rc = ref.meshopt_decodeIndexSequence(out.ctypes.data_as(C.c_void_p), C.c_size_t(idx.size), C.c_size_t(index_size), buf.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        assert rc == 0
        return buf[:n], out

    for name, idx, nverts in (("strip", np.arange(1000, dtype=np.uint32), 1000), ("two_baselines", np.stack([np.arange(400), 100000 + np.arange(400)], 1).astype(np.uint32).reshape(-1), 200000),
                              ("random", rng.integers(0, 2**32, 777, dtype=np.uint64).astype(np.uint32), 2**32 - 1), ("lines", grid_indices(12)[:601], 144)):
        for isz in (4, 2):
            if isz == 2 and idx.max() >= 65536:
                continue
            e, d = enc_seq(np.ascontiguousarray(idx, np.uint32), nverts, 1, isz)
            add(f"sequence_{name}_{isz * 8}", "sequence", idx.size, isz, 1, e, d)
    ref.meshopt_encodeIndexVersion(1)

    # filters: random inputs through the reference's scalar definitions (NO_SIMD build); inputs shaped like encoder output.
    # Stored the way a glTF carries them: the filter's INPUT as an encoded attribute stream + the filter id; decoded = filtered.
    FID = {"filter_oct": 1, "filter_quat": 2, "filter_exp": 3}

    def add_filtered(name, kind, cnt, stride, inp, out):
        e, d = enc_vertex(inp, cnt, stride)
        assert np.array_equal(d, inp)
        add(name, "vertex", cnt, stride, FID[kind], e, out)

    n = 1003
    o8 = rng.integers(-127, 128, (n, 4)).astype(np.int8)
    o8[:, 2] = 127
    o8[:, 3] = rng.integers(0, 2, n)
    o16 = rng.integers(-2047, 2048, (n, 4)).astype(np.int16)
    o16[:, 2] = rng.choice([2047, 1023, 32767], n)
    o16[:, :2] = np.where(o16[:, 2:3] == 32767, rng.integers(-32767, 32768, (n, 2)), o16[:, :2])
    q16 = rng.integers(-2047, 2048, (n, 4)).astype(np.int16)
    q16[:, 3] = (2047 & ~3) | rng.integers(0, 4, n)
    ex = ((rng.integers(-100, 20, 4 * n).astype(np.int32) << 24) | (rng.integers(-2**23, 2**23, 4 * n).astype(np.int32) & 0xFFFFFF)).astype(np.uint32)
    for name, kind, arr, stride in (("oct8", "filter_oct", o8, 4), ("oct16", "filter_oct", o16, 8), ("quat16", "filter_quat", q16, 8), ("exp", "filter_exp", ex, 4)):
        inp = np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()
        out = inp.copy()
        cnt = out.size // stride
        {"filter_oct": ref_ns.meshopt_decodeFilterOct, "filter_quat": ref_ns.meshopt_decodeFilterQuat, "filter_exp": ref_ns.meshopt_decodeFilterExp}[kind](
            out.ctypes.data_as(C.c_void_p), C.c_size_t(cnt), C.c_size_t(stride))
        add_filtered(f"filter_{name}_random", kind, cnt, stride, inp, out)
    for name, kind, cnt, stride, d, expected in kat_filters:
        add_filtered(name, kind, cnt, stride, np.ascontiguousarray(d).view(np.uint8).reshape(-1), expected)

    # ---- store
    names = np.array([c[0] for c in cases])
    kinds = np.array([c[1] for c in cases])
    params = np.array([[c[2], c[3], c[4]] for c in cases], np.int64)
    enc_off = np.cumsum([0] + [c[5].size for c in cases]).astype(np.int64)
    dec_off = np.cumsum([0] + [c[6].size for c in cases]).astype(np.int64)
    out = os.path.join(HERE, "meshopt_codec.npz")
    np.savez_compressed(out, names=names, kinds=kinds, params=params, enc=np.concatenate([c[5] for c in cases]), enc_off=enc_off,
                        dec=np.concatenate([c[6] for c in cases]), dec_off=dec_off)
    print(f"{len(cases)} cases, {enc_off[-1]} encoded / {dec_off[-1]} decoded bytes -> {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    sys.exit(main())

"""ctypes faces used by the meshopt-codec tests: the oracle restatement (oracle/meshopt_decode.cpp), the reference's own
meshoptimizer when oracle/_ref is present, and the golden fixture reader.  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

from tests.conftest import ROOT
from tests.oracle_lib import lib as oracle_lib

GOLDEN = os.path.join(ROOT, "tests", "golden", "meshopt_codec.npz")
MODE = {"vertex": 0, "index": 1, "sequence": 2}


def golden_cases():
    z = np.load(GOLDEN)
    out = []
    for i, name in enumerate(z["names"]):
        enc = z["enc"][z["enc_off"][i]:z["enc_off"][i + 1]]
        dec = z["dec"][z["dec_off"][i]:z["dec_off"][i + 1]]
        count, stride, extra = (int(x) for x in z["params"][i])
        kind = str(z["kinds"][i])
        out.append(dict(name=str(name), kind=kind, count=count, stride=stride, filter=extra if kind == "vertex" else 0, enc=enc.copy(), dec=dec.copy()))
    return out


def _sz(x):
    return C.c_size_t(int(x))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle_filter(fid, count, stride, data):
    L = oracle_lib()
    out = np.ascontiguousarray(data, np.uint8).copy()
    if fid == 1:
        L.orc_meshopt_filter_oct(_ptr(out), _sz(count), _sz(stride))
    elif fid == 2:
        L.orc_meshopt_filter_quat(_ptr(out), _sz(count))
    elif fid == 3:
        L.orc_meshopt_filter_exp(_ptr(out), _sz(count * stride // 4))
    return out


def oracle_decode(kind, count, stride, enc, fid=0):
    """-> (rc, bytes) through the CPU restatement: codec, then the filter `fid` (0 none, 1 oct, 2 quat, 3 exp)"""
    L = oracle_lib()
    enc = np.ascontiguousarray(enc, np.uint8)
    out = np.zeros(max(1, count * stride), np.uint8)
    fn = {"vertex": L.orc_meshopt_decode_vertex, "index": L.orc_meshopt_decode_index, "sequence": L.orc_meshopt_decode_sequence}[kind]
    fn.restype = C.c_int
    rc = fn(_ptr(out), _sz(count), _sz(stride), _ptr(enc), _sz(enc.size))
    out = out[:count * stride]
    if rc == 0 and fid:
        out = oracle_filter(fid, count, stride, out)
    return rc, out


_ref = {}


def ref_lib(nosimd=False):
    """the reference's meshoptimizer built from its sources by oracle/build_ref.sh, or None"""
    name = "libmeshopt_ref_nosimd.so" if nosimd else "libmeshopt_ref.so"
    if name not in _ref:
        p = os.path.join(ROOT, "oracle", "_ref", name)
        L = C.CDLL(p) if os.path.exists(p) else None
        if L is not None:
            for f in ("meshopt_encodeVertexBuffer", "meshopt_encodeVertexBufferBound", "meshopt_encodeIndexBuffer", "meshopt_encodeIndexBufferBound",
                      "meshopt_encodeIndexSequence", "meshopt_encodeIndexSequenceBound"):
                getattr(L, f).restype = C.c_size_t
        _ref[name] = L
    return _ref[name]


def ref_decode(kind, count, stride, enc, nosimd=False, fid=0):
    L = ref_lib(nosimd)
    enc = np.ascontiguousarray(enc, np.uint8)
    out = np.zeros(max(1, count * stride), np.uint8)
    fn = {"vertex": L.meshopt_decodeVertexBuffer, "index": L.meshopt_decodeIndexBuffer, "sequence": L.meshopt_decodeIndexSequence}[kind]
    rc = fn(_ptr(out), _sz(count), _sz(stride), _ptr(enc), _sz(enc.size))
    if rc == 0 and fid:
        {1: L.meshopt_decodeFilterOct, 2: L.meshopt_decodeFilterQuat, 3: L.meshopt_decodeFilterExp}[fid](_ptr(out), _sz(count), _sz(stride))
    return rc, out[:count * stride]


def ref_encode(kind, data, count, stride, nverts=0, version=1):
    L = ref_lib()
    if kind == "vertex":
        data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        bound = L.meshopt_encodeVertexBufferBound(_sz(count), _sz(stride))
        buf = np.zeros(bound, np.uint8)
        n = L.meshopt_encodeVertexBuffer(_ptr(buf), _sz(bound), _ptr(data), _sz(count), _sz(stride))
        return buf[:n].copy()
    idx = np.ascontiguousarray(data, np.uint32)
    L.meshopt_encodeIndexVersion(version)
    if kind == "index":
        bound = L.meshopt_encodeIndexBufferBound(_sz(idx.size), _sz(nverts))
        buf = np.zeros(bound, np.uint8)
        n = L.meshopt_encodeIndexBuffer(_ptr(buf), _sz(bound), _ptr(idx), _sz(idx.size))
    else:
        bound = L.meshopt_encodeIndexSequenceBound(_sz(idx.size), _sz(nverts))
        buf = np.zeros(bound, np.uint8)
        n = L.meshopt_encodeIndexSequence(_ptr(buf), _sz(bound), _ptr(idx), _sz(idx.size))
    L.meshopt_encodeIndexVersion(1)
    return buf[:n].copy()

"""Accessor conversions (assets.cpp:308-320 through fastgltf::internal::convertComponent, tools.hpp:266-289): the CPU restatement
against fastgltf's own function compiled from the reference tree — exhaustively, every 8- and 16-bit input, normalized or not —
and against the frozen SHA-256 of those tables (tests/golden/accessor_tables.json) where oracle/_ref does not exist."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from tests.conftest import ROOT
from tests.oracle_lib import lib as oracle_lib

GOLDEN = os.path.join(ROOT, "tests", "golden", "accessor_tables.json")
TYPES = {5120: range(-128, 128), 5121: range(0, 256), 5122: range(-32768, 32768), 5123: range(0, 65536)}


def table(fn, t, normalized):
    fn.restype = C.c_float
    return np.array([fn(t, normalized, v) for v in TYPES[t]], np.float32)


def ref_shim():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_shim.so")
    return C.CDLL(p) if os.path.exists(p) else None


def test_oracle_tables_match_the_frozen_digests():
    want = json.load(open(GOLDEN))
    for t in TYPES:
        for n in (0, 1):
            got = hashlib.sha256(table(oracle_lib().orc_convert_component, t, n).tobytes()).hexdigest()
            assert got == want[f"{t}_{n}"], f"componentType {t} normalized {n}"


@pytest.mark.skipif(ref_shim() is None or not hasattr(ref_shim(), "ref_convert_component"), reason="oracle/_ref not built")
def test_oracle_equals_fastgltf_exhaustively():
    R = ref_shim()
    for t in TYPES:
        for n in (0, 1):
            a, b = table(R.ref_convert_component, t, n), table(oracle_lib().orc_convert_component, t, n)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"componentType {t} normalized {n}"
            if n:
                assert a.min() >= -1.0 and a.max() == 1.0


def test_assemble_and_widen_layouts():
    L = oracle_lib()
    rng = np.random.default_rng(3)
    q = rng.integers(-32768, 32768, (50, 4)).astype(np.int16)          # VEC3 of SHORT with 8-byte stride (KHR_mesh_quantization padding)
    out = np.full(50 * 24, 0xAB, np.uint8)
    assert L.orc_assemble_vertices(q.ctypes.data_as(C.c_void_p), 5122, 1, C.c_size_t(8), C.c_size_t(50), out.ctypes.data_as(C.c_void_p)) == 0
    v = out.view(np.float32).reshape(50, 6)
    assert np.array_equal(v[:, :3], np.maximum(q[:, :3].astype(np.float32) / np.float32(32767), np.float32(-1)))
    assert not out.reshape(50, 24)[:, 12:].any()
    idx8 = rng.integers(0, 256, 77).astype(np.uint8)
    o = np.zeros(77, np.uint32)
    assert L.orc_widen_indices(idx8.ctypes.data_as(C.c_void_p), 5121, C.c_size_t(77), o.ctypes.data_as(C.c_void_p)) == 0 and np.array_equal(o, idx8)

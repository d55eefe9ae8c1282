"""GPU parity of the accessor conversions (vkv_assemble_vertices / vkv_widen_indices) against the oracle restatement, which
tests/test_accessors.py pins to fastgltf's own convertComponent: every 8- and 16-bit input, normalized or not, bit for bit;
strided layouts; index widening."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_lib import lib as oracle_lib
from vk_gltf_viewer_b200 import api

pytestmark = pytest.mark.gpu
DT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5126: np.float32}


def oracle_vertices(raw, t, normalized, stride, count):
    out = np.zeros(count * 24, np.uint8)
    assert oracle_lib().orc_assemble_vertices(raw.ctypes.data_as(C.c_void_p), t, int(normalized), C.c_size_t(stride), C.c_size_t(count), out.ctypes.data_as(C.c_void_p)) == 0
    return out


@pytest.mark.parametrize("t", [5120, 5121, 5122, 5123])
@pytest.mark.parametrize("normalized", [False, True])
def test_every_integer_input(t, normalized):
    info = np.iinfo(DT[t])
    vals = np.arange(info.min, info.max + 1, dtype=np.int64).astype(DT[t])
    vals = np.concatenate([vals, vals[:(-vals.size) % 3]]).reshape(-1, 3)     # VEC3, tightly packed
    r = api.Renderer(64, 64)
    src = r.upload(vals)
    got = r.download(r.assemble_vertices(src, t, normalized, 0, vals.shape[0]), vals.shape[0] * 24)
    r.close()
    want = oracle_vertices(np.ascontiguousarray(vals).view(np.uint8).reshape(-1), t, normalized, 0, vals.shape[0])
    assert np.array_equal(got, want)


def test_strided_quantised_and_float_layouts():
    rng = np.random.default_rng(9)
    r = api.Renderer(64, 64)
    q = rng.integers(-32768, 32768, (1000, 4)).astype(np.int16)               # KHR_mesh_quantization: SHORT VEC3 padded to 8 bytes
    f = rng.standard_normal((777, 5)).astype(np.float32)                      # interleaved float attributes, 20-byte stride
    f[::50, 1] = np.inf; f[::77, 2] = np.nan
    b = rng.integers(0, 256, (300, 4)).astype(np.uint8)
    for arr, t, n, stride in ((q, 5122, True, 8), (q, 5122, False, 8), (f, 5126, False, 20), (b, 5121, True, 4)):
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        got = r.download(r.assemble_vertices(r.upload(raw), t, n, stride, arr.shape[0]), arr.shape[0] * 24)
        assert np.array_equal(got, oracle_vertices(raw, t, n, stride, arr.shape[0]))
    r.close()


@pytest.mark.parametrize("t", [5121, 5123, 5125])
def test_index_widening(t):
    rng = np.random.default_rng(t)
    dt = {5121: np.uint8, 5123: np.uint16, 5125: np.uint32}[t]
    idx = rng.integers(0, np.iinfo(dt).max, 3000 * 3, dtype=np.uint64).astype(dt)
    r = api.Renderer(64, 64)
    got = r.download(r.widen_indices(r.upload(idx), t, idx.size), idx.size * 4).view(np.uint32)
    r.close()
    assert np.array_equal(got, idx.astype(np.uint32))


def test_validation():
    r = api.Renderer(64, 64)
    with pytest.raises(api.VkvError):
        r.assemble_vertices(r.alloc(64), 5125, False, 0, 4)      # UNSIGNED_INT positions do not exist
    with pytest.raises(api.VkvError):
        r.assemble_vertices(r.alloc(64), 5122, False, 4, 4)      # stride below the element size
    with pytest.raises(api.VkvError):
        r.widen_indices(r.alloc(64), 5122, 4)
    r.close()

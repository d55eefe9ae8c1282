"""GPU parity of the EXT_meshopt_compression decoders through the C ABI (vkv_meshopt_*): PINNED — the expected
bytes are the reference's own known-answer vectors and reference-encoded/decoded streams frozen in tests/golden/meshopt_codec.npz
(generator: tests/golden/make_meshopt_golden.py); the CPU restatement is checked against the same fixture in
tests/test_meshopt_codec.py.  Bit-exact for every byte; return codes equal to meshoptimizer's on malformed streams."""
import numpy as np
import pytest

from tests import meshopt_lib as M
from vk_gltf_viewer_b200 import abi, api

pytestmark = pytest.mark.gpu
CASES = M.golden_cases()


def pack(cases):
    """lay the cases' encoded streams out in one 'glTF buffer' and their outputs in one destination buffer"""
    views = np.zeros(len(cases), abi.MESHOPT_VIEW_DTYPE)
    src, so, do = [], 0, 0
    for i, c in enumerate(cases):
        pad = (-so) % 4  # bufferView.byteOffset alignment of real files is 4 at best; keep some streams unaligned on purpose
        if i % 3 == 1:
            pad += 1
        src.append(np.zeros(pad, np.uint8)); so += pad
        views[i] = (M.MODE[c["kind"]], c["filter"], c["count"], c["stride"], so, c["enc"].size, do)
        src.append(c["enc"]); so += c["enc"].size
        do += (c["count"] * c["stride"] + 15) & ~15
    return np.concatenate(src + [np.zeros(32, np.uint8)]), views, do


def test_every_golden_case_in_one_decode():
    src, views, dst_bytes = pack(CASES)
    r = api.Renderer(64, 64)
    out, rc = r.meshopt_decode(src, views, dst_bytes)
    r.close()
    for c, v, code in zip(CASES, views, rc):
        assert code == 0, f"{c['name']}: return code {code}"
        got = out[int(v["dst_offset"]):int(v["dst_offset"]) + c["dec"].size]
        bad = got != c["dec"]
        assert not bad.any(), f"{c['name']}: {int(bad.sum())} of {bad.size} bytes differ (first at {int(np.argmax(bad))})"


def test_known_answer_vectors_alone():
    """demo/tests.cpp:24-59,515-629 one view at a time (a plan with a single tiny stream)"""
    r = api.Renderer(64, 64)
    for c in CASES:
        if not c["name"].startswith("kat_"):
            continue
        src, views, dst_bytes = pack([c])
        out, rc = r.meshopt_decode(src, views, dst_bytes)
        assert rc[0] == 0 and np.array_equal(out[:c["dec"].size], c["dec"]), c["name"]
    r.close()


def test_malformed_streams_return_meshoptimizer_codes_and_do_not_disturb_their_neighbours():
    """truncated / extended / wrong header / wrong version (demo/tests.cpp:119-214,271-340,381-432): the oracle (pinned to the
    reference's return codes in tests/test_meshopt_codec.py) gives the expected code; good views in the same plan still decode"""
    rng = np.random.default_rng(5)
    good = [c for c in CASES if c["name"] in ("vertex_smooth_1000x12", "index_grid40_v1_32", "sequence_strip_32", "vertex_mixed_2000x12")]
    assert len(good) == 4
    broken = []
    for c in good:
        e = c["enc"]
        for variant in (e[:e.size // 2], e[:e.size - 1], np.concatenate([e, np.zeros(3, np.uint8)]), np.concatenate([[0x00], e[1:]]).astype(np.uint8),
                        np.concatenate([[e[0] | 0x0F], e[1:]]).astype(np.uint8), e[:int(rng.integers(1, 40))]):
            b = dict(c); b["enc"] = np.ascontiguousarray(variant, np.uint8); b["name"] = c["name"] + f"_broken{len(broken)}"
            broken.append(b)
    cases = []
    for i, b in enumerate(broken):
        cases += [b, good[i % 4]]
    src, views, dst_bytes = pack(cases)
    r = api.Renderer(64, 64)
    out, rc = r.meshopt_decode(src, views, dst_bytes)
    r.close()
    for c, v, code in zip(cases, views, rc):
        want, _ = M.oracle_decode(c["kind"], c["count"], c["stride"], c["enc"], c["filter"])
        assert code == want, f"{c['name']}: gpu {code}, oracle {want}"
        if "_broken" not in c["name"]:
            assert code == 0 and np.array_equal(out[int(v["dst_offset"]):int(v["dst_offset"]) + c["dec"].size], c["dec"])
    assert sum(1 for c, code in zip(cases, rc) if "_broken" in c["name"] and code != 0) >= len(broken) - 2


def test_large_asset_round_trip_property():
    """size-independent property at a realistic size: many streams replicated from the fixture (2,000 buffer views, ~20 MB of
    output) decode to exactly the replicated expected bytes, whatever their position in the buffers"""
    base = [c for c in CASES if c["kind"] in M.MODE and c["count"] >= 255]
    cases = [base[i % len(base)] for i in range(2000)]
    src, views, dst_bytes = pack(cases)
    r = api.Renderer(64, 64)
    out, rc = r.meshopt_decode(src, views, dst_bytes)
    r.close()
    assert (rc == 0).all()
    want = np.zeros(dst_bytes, np.uint8)
    for c, v in zip(cases, views):
        want[int(v["dst_offset"]):int(v["dst_offset"]) + c["dec"].size] = c["dec"]
    assert np.array_equal(out, want)


def test_plan_validation():
    r = api.Renderer(64, 64)
    for bad in ((0, 0, 4, 6, 0, 64, 0), (1, 0, 4, 4, 0, 64, 0), (0, 2, 4, 12, 0, 64, 0), (3, 0, 4, 4, 0, 64, 0), (0, 0, 4, 12, 0, 64, 2), (2, 1, 4, 4, 0, 64, 0)):
        v = np.zeros(1, abi.MESHOPT_VIEW_DTYPE); v[0] = bad
        with pytest.raises(api.VkvError):
            r.meshopt_plan(v)
    v = np.zeros(1, abi.MESHOPT_VIEW_DTYPE); v[0] = (0, 0, 4, 12, 0, 64, 0)
    plan = r.meshopt_plan(v)
    with pytest.raises(api.VkvError):
        r.meshopt_run(plan, r.alloc(64), 32, r.alloc(64), 64)   # source shorter than the view
    r.meshopt_plan_destroy(plan)
    r.close()

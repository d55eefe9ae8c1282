"""EXT_meshopt_compression decoders: the CPU restatement (oracle/meshopt_decode.cpp) against
 * the reference's known-answer vectors and reference-encoded/decoded streams frozen in tests/golden/meshopt_codec.npz,
 * the reference's meshoptimizer itself (oracle/_ref, when present) on fresh random streams and on malformed ones.
The CUDA decoders are checked against the same fixture in tests/test_gpu_meshopt.py."""
import numpy as np
import pytest

from tests import meshopt_lib as M

CASES = M.golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_golden(case):
    rc, out = M.oracle_decode(case["kind"], case["count"], case["stride"], case["enc"], case["filter"])
    assert rc == 0
    assert np.array_equal(out, case["dec"]), f"{case['name']}: {(out != case['dec']).sum()} bytes differ"


def test_golden_holds_the_reference_known_answer_vectors():
    names = {c["name"] for c in CASES}
    for n in ("kat_index_v0", "kat_index_v1", "kat_sequence_v1", "kat_vertex_v0", "kat_decodeFilterOct8", "kat_decodeFilterOct12",
              "kat_decodeFilterQuat12", "kat_decodeFilterExp"):
        assert n in names
    c = next(c for c in CASES if c["name"] == "kat_index_v0")  # demo/tests.cpp:24-30
    assert c["dec"].view(np.uint32).tolist() == [0, 1, 2, 2, 1, 3, 4, 6, 5, 7, 8, 9] and c["enc"][0] == 0xE0


needs_ref = pytest.mark.skipif(M.ref_lib() is None, reason="oracle/_ref not built (make ref needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_reference_on_fresh_streams(seed):
    rng = np.random.default_rng(1000 + seed)
    # vertex streams: random stride / count / smoothness
    stride = int(rng.choice([4, 8, 12, 16, 20, 24, 40]))
    count = int(rng.integers(1, 3000))
    data = np.cumsum(rng.integers(-3, 4, (count, stride)) * rng.integers(0, 2, (1, stride)) * int(rng.choice([1, 9, 70])), 0).astype(np.uint8)
    enc = M.ref_encode("vertex", data, count, stride)
    a, b = M.ref_decode("vertex", count, stride, enc), M.oracle_decode("vertex", count, stride, enc)
    assert a[0] == b[0] == 0 and np.array_equal(a[1], b[1]) and np.array_equal(a[1], data.reshape(-1))
    # index streams
    nverts = int(rng.integers(3, 70000))
    tris = int(rng.integers(1, 2000))
    walk = np.clip(np.cumsum(rng.integers(-2, 3, 3 * tris)) + nverts // 2, 0, nverts - 1).astype(np.uint32)
    for version in (0, 1):
        enc = M.ref_encode("index", walk, walk.size, 4, nverts, version)
        for isz in (2, 4):
            a, b = M.ref_decode("index", walk.size, isz, enc), M.oracle_decode("index", walk.size, isz, enc)
            assert a[0] == b[0] == 0 and np.array_equal(a[1], b[1])
    enc = M.ref_encode("sequence", walk, walk.size, 4, nverts)
    a, b = M.ref_decode("sequence", walk.size, 4, enc), M.oracle_decode("sequence", walk.size, 4, enc)
    assert a[0] == b[0] == 0 and np.array_equal(a[1], b[1])


@needs_ref
def test_oracle_error_codes_match_reference_on_malformed_streams():
    """demo/tests.cpp:119-214,271-340,381-432: truncation, trailing bytes, bad headers, bad versions — same return code"""
    rng = np.random.default_rng(7)
    data = np.cumsum(rng.integers(-2, 3, (300, 12)), 0).astype(np.uint8)
    idx = np.clip(np.cumsum(rng.integers(-2, 3, 600)) + 100, 0, 199).astype(np.uint32)
    streams = [("vertex", 300, 12, M.ref_encode("vertex", data, 300, 12)), ("index", 600, 4, M.ref_encode("index", idx, 600, 4, 200)),
               ("sequence", 600, 4, M.ref_encode("sequence", idx, 600, 4, 200))]
    for kind, count, stride, enc in streams:
        variants = [enc[:n] for n in sorted(set(rng.integers(0, enc.size, 40).tolist() + [0, 1, 2, enc.size - 1]))]
        variants += [np.concatenate([enc, np.zeros(k, np.uint8)]) for k in (1, 5)]
        for b in (0x00, 0xA1, 0xE2, 0xD2, 0xFF):
            v = enc.copy(); v[0] = b; variants.append(v)
        for v in variants:
            v = np.ascontiguousarray(np.concatenate([v, np.zeros(0, np.uint8)]))
            if v.size == 0:
                continue
            ra = M.ref_decode(kind, count, stride, v)[0]
            rb = M.oracle_decode(kind, count, stride, v)[0]
            assert (ra == 0) == (rb == 0) and ra == rb, f"{kind}: size {v.size} first byte {v[0]:#x}: reference {ra}, oracle {rb}"


@needs_ref
def test_scalar_filter_definition_vs_the_sse_build():
    """the oracle follows vertexfilter.cpp's scalar definitions (== the NO_SIMD build, bit for bit); the SSE build of the same
    source rounds differently — at most one unit in a component; exp is exact in both"""
    seen = 0
    for c in CASES:
        if not c["filter"]:
            continue
        seen += 1
        _, scalar = M.ref_decode("vertex", c["count"], c["stride"], c["enc"], nosimd=True, fid=c["filter"])
        _, sse = M.ref_decode("vertex", c["count"], c["stride"], c["enc"], nosimd=False, fid=c["filter"])
        _, orc = M.oracle_decode("vertex", c["count"], c["stride"], c["enc"], c["filter"])
        assert np.array_equal(orc, scalar) and np.array_equal(orc, c["dec"])
        dt = np.int8 if c["stride"] == 4 and c["filter"] == 1 else (np.uint32 if c["filter"] == 3 else np.int16)
        d = np.abs(sse.view(dt).astype(np.int64) - orc.view(dt).astype(np.int64))
        assert d.max() <= (0 if c["filter"] == 3 else 1), f"{c['name']}: SSE build differs by {d.max()}"
    assert seen >= 8


def test_oracle_decoder_against_the_reference_on_mutated_streams(meshopt_ref):
    """tools/diff_meshopt_decode.py: same return code, same bytes, on reference-encoded streams with random damage (short run; 20 000 mutants
    in the tool's docstring)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "diff_meshopt_decode.py"), "5", "1500"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "mismatches 0" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]

#!/usr/bin/env python
"""Differential fuzz of the oracle's EXT_meshopt_compression decoder (oracle/meshopt_decode.cpp — what the device decoder is held to) against
the reference's meshoptimizer (oracle/_ref, scalar build): the fixture's reference-encoded streams with bit flips, byte replacements,
truncations, appended bytes and wrong element counts must give the same return code, and the same bytes when both decode.

    python tools/diff_meshopt_decode.py [seed] [mutants]          (20 000 mutants at seed 1: 0 mismatches)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tests import meshopt_lib as M
cases = [c for c in M.golden_cases() if c["kind"] in M.MODE]
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
N = int(sys.argv[2]) if len(sys.argv)>2 else 5000
bad = 0; okboth = 0; errboth = 0
for it in range(N):
    c = cases[int(rng.integers(len(cases)))]
    e = c["enc"].copy()
    k = int(rng.integers(0, 5))
    if k == 0 and e.size: e[int(rng.integers(e.size))] ^= np.uint8(1 << int(rng.integers(8)))
    elif k == 1 and e.size: e[int(rng.integers(e.size))] = int(rng.integers(256))
    elif k == 2 and e.size > 2: e = e[:int(rng.integers(1, e.size))]
    elif k == 3: e = np.concatenate([e, rng.integers(0,256,int(rng.integers(1,20))).astype(np.uint8)])
    else:
        for _ in range(int(rng.integers(2,8))):
            if e.size: e[int(rng.integers(e.size))] = int(rng.integers(256))
    count, stride = c["count"], c["stride"]
    if rng.random() < 0.2: count = max(0, count + int(rng.integers(-3,4)) * (3 if c["kind"]=="index" else 1))
    r_rc, r_out = M.ref_decode(c["kind"], count, stride, e, nosimd=True)
    o_rc, o_out = M.oracle_decode(c["kind"], count, stride, e)
    if r_rc != o_rc or (r_rc == 0 and not np.array_equal(r_out, o_out)):
        bad += 1
        if bad <= 10: print("MISMATCH", c["name"], c["kind"], "count", count, "stride", stride, "ref rc", r_rc, "oracle rc", o_rc, "outputs equal", bool(np.array_equal(r_out,o_out)) if r_rc==o_rc==0 else None)
    okboth += (r_rc == 0); errboth += (r_rc != 0)
print("mutants", N, "decoded by both", okboth, "refused by both", errboth, "mismatches", bad)
sys.exit(1 if bad else 0)

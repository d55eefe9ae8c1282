"""Meshlet partition + bounds: the CPU restatement against the reference's meshopt_buildMeshletsScan — frozen in
tests/golden/meshlet_scan.npz and, when oracle/_ref is present, live on fresh meshes — and the validity rules every partition
must meet (each triangle exactly once, <= 64 vertices / 124 triangles, bounds contain the vertices)."""
import numpy as np
import pytest

from tests import meshlet_lib as ML
from tests import meshopt_lib as M

G = np.load(ML.GOLDEN)
NAMES = [str(n) for n in G["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    pos, idx = G[name + "_pos"], G[name + "_idx"]
    m, mv, mt = ML.oracle_scan(idx, pos.shape[0])
    assert np.array_equal(m, G[name + "_m"]) and np.array_equal(mv, G[name + "_mv"]) and np.array_equal(mt, G[name + "_mt"])


@pytest.mark.parametrize("name", NAMES)
def test_partition_is_valid_and_bounds_contain(name):
    pos, idx = G[name + "_pos"], G[name + "_idx"]
    m, mv, mt = ML.oracle_scan(idx, pos.shape[0])
    tris = []
    for vo, to, vc, tc in m:
        assert 1 <= vc <= ML.MAXV and 1 <= tc <= ML.MAXT and to % 4 == 0
        local = mt[to:to + tc * 3].reshape(-1, 3)
        assert local.max() < vc
        tris.append(mv[vo:vo + vc][local])
    assert np.array_equal(np.concatenate(tris).reshape(-1), idx)          # every triangle exactly once, in order
    b = ML.oracle_bounds(m, mv, pos)
    for (vo, to, vc, tc), (ex, ey, ez, cx, cy, cz) in zip(m, b):
        p = pos[mv[vo:vo + vc]]
        c, e = np.array([cx, cy, cz]), np.array([ex, ey, ez])
        assert (np.abs(p - c) <= e * (1 + 1e-6) + 1e-7).all()


@pytest.mark.skipif(M.ref_lib() is None, reason="oracle/_ref not built (make ref needs /root/reference)")
@pytest.mark.parametrize("seed", range(1, 4))
def test_oracle_matches_reference_scan_builder_on_fresh_meshes(seed):
    for name, (pos, idx) in ML.meshes(seed).items():
        for maxv, maxt in ((64, 124), (32, 64), (3, 4), (64, 8)):
            a, b = ML.ref_scan(idx, pos.shape[0], maxv, maxt), ML.oracle_scan(idx, pos.shape[0], maxv, maxt)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), f"{name} {maxv}/{maxt}"

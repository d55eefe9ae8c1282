#!/usr/bin/env python
"""Freeze golden vectors for the device-side meshlet builder -> tests/golden/meshlet_scan.npz
Run in the build container only: the outputs come from meshopt_buildMeshletsScan of the REFERENCE's meshoptimizer built from
source (oracle/_ref/libmeshopt_ref.so, `make ref`) on the meshes of tests/meshlet_lib.py::meshes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import meshlet_lib as ML  # noqa: E402

out = {}
names = []
for name, (pos, idx) in ML.meshes().items():
    m, mv, mt = ML.ref_scan(idx, pos.shape[0])
    names.append(name)
    out[name + "_pos"], out[name + "_idx"], out[name + "_m"], out[name + "_mv"], out[name + "_mt"] = pos, idx, m, mv, mt
out["names"] = np.array(names)
np.savez_compressed(ML.GOLDEN, **out)
print(len(names), "meshes ->", ML.GOLDEN, os.path.getsize(ML.GOLDEN), "bytes")

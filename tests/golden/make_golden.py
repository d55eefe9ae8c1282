#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference ships no golden images or known-answer vectors for cull / raster / HiZ, and its shaders
cannot be run in this image, so these fixtures are NOT reference outputs: they freeze the oracle's answers (themselves
pinned where possible against oracle/_ref, see tests/test_oracle.py) so that neither the oracle nor the CUDA path can
drift silently.  Inputs are rebuilt from the seeded procedural scenes; the fixture stores a digest of the inputs too.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests import oracle_lib as O  # noqa: E402
from tests import scenes as S  # noqa: E402
from vk_gltf_viewer_b200.scene import Camera, Scene  # noqa: E402

CASES = {
    "icosphere16_160x120": (lambda: Scene.icosphere(16), (160, 120), [((0, 0, 3), (0, 0, 0)), ((0.6, 0.3, 2.8), (0, 0, 0))]),
    "atrium8_192x108": (lambda: Scene.atrium(8), (192, 108), [((0.5, 2.5, 27), (0, 4, 0)), ((1.5, 3.0, 20), (0, 4, 0))]),
    "soup_128x96": (lambda: S.random_soup(120, seed=5), (128, 96), [((0, 0, 3), (0, 0, 0)), ((0.5, 0.2, 2), (0, 0, -2))]),
    "ground_200x150": (lambda: S.ground_plane(10, 30.0), (200, 150), [((0, 0, 3), (0, -0.2, 0)), ((0, 2, 0), (0.3, -1, 0.2))]),
}


def input_digest(scene):
    h = hashlib.sha256()
    h.update(scene.draws().tobytes()); h.update(scene.transforms().tobytes()); h.update(scene.materials().tobytes())
    for i in range(scene.counts().primitives):
        p = scene.primitive(i)
        for k in ("vertex_indices", "triangles", "vertices", "meshlets"):
            h.update(p[k].tobytes())
    return h.hexdigest()


def run_case(name):
    make, (W, H), views = CASES[name]
    scene = make()
    cam = Camera(W, H)
    cam.look_at(*views[0])
    pc = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    out = {"input_digest": np.frombuffer(bytes.fromhex(input_digest(scene)), np.uint8)}
    for k, v in enumerate(views):
        if k:
            cam.look_at(*v)
        f = O.frame(pc, tg, two_pass=True)
        out[f"v{k}_camera"] = np.frombuffer(cam.raw(), np.uint8).copy()
        out[f"v{k}_visibleA"] = f["visibleA"]; out[f"v{k}_visibleB"] = f["visibleB"]
        out[f"v{k}_vis64"] = tg.vis64(); out[f"v{k}_ids_ref"] = tg.ids_ref.copy(); out[f"v{k}_tie"] = np.packbits(tg.tie)
        out[f"v{k}_pyramid"] = tg.pyramid.copy()
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(name))
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")

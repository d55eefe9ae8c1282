#!/usr/bin/env python
"""Regenerates tests/golden/llvmpipe.npz (python tests/golden/make_llvmpipe_golden.py): OUTPUTS OF LLVMPIPE — Mesa's software rasteriser, the
engine underneath lavapipe — for the fixed-function stages the reference's hot path relies on and ships no code for: triangle coverage + the
depth test, and the LINEAR sampler's texel footprint.  Unlike tests/golden/make_golden.py's files these are not the oracle's own answers:
they come from an independent, widely deployed implementation (Mesa 18.1.9 as found inside this image's Nsight Compute, driven without an X
server by tests/llvmpipe_lib.py), so they travel to machines without that library and hold both the CPU oracle (tests/test_golden.py::
test_oracle_matches_llvmpipe_*) and the CUDA path (::test_cuda_matches_llvmpipe_raster) to it.

Raster cases store the inputs that cannot be rebuilt bit for bit elsewhere (soup positions) or a digest of them (procedural scenes), and
llvmpipe's id image (packVisBuffer ids, 0xFFFFFFFF = nothing drawn) and depth image.  Sampler cases store coordinates and, per coordinate,
which texel classes (x % 3, y % 3) received a non-zero LINEAR weight.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests import llvmpipe_cases as K  # noqa: E402

PATH = os.path.join(HERE, "llvmpipe.npz")

# name -> (kind, parameters); "soup" cases use the identity camera (positions are clip space, w = 1)
RASTER_CASES = {
    "lattice64": ("lattice", dict(seed=5, size=64, ntri=300)),
    "lattice256": ("lattice", dict(seed=7, size=256, ntri=2000)),
    "float257x193": ("float", dict(seed=16, W=257, H=193, ntri=3000, scale=0.3)),
    "icosphere": ("scene", dict(name="icosphere")),
    "ground_clipped": ("scene", dict(name="ground_clipped")),
    "mirrored": ("scene", dict(name="mirrored")),
}
SAMPLER_SIZES = [(1, 1), (5, 3), (67, 120), (960, 540)]
HIZ_CHAINS = [(97, 33), (480, 270)]           # 270 -> 135 -> 67: odd sources, where the pyramid is not conservative


def build_raster_case(kind, p, stored_positions=None):
    """-> (scene, camera, W, H, positions or None); positions come from the fixture when given (rng streams are not part of the contract)"""
    if kind == "scene":
        make, (W, H) = K.SCENE_CASES[p["name"]]
        scene, cam = make(W, H)
        return scene, cam, W, H, None
    if kind == "lattice":
        W = H = p["size"]
        P = stored_positions if stored_positions is not None else K.lattice_positions(np.random.default_rng(p["seed"]), W, H, p["ntri"])
    else:
        W, H = p["W"], p["H"]
        P = stored_positions if stored_positions is not None else K.float_positions(np.random.default_rng(p["seed"]), p["ntri"], p["scale"], inside=True)
    return K.soup_scene(P), K.identity_camera(W, H), W, H, P


def triangles_digest(V, I):
    return np.frombuffer(hashlib.sha256(V.tobytes() + I.tobytes()).digest(), np.uint8)


def hiz_coords(dw, dh):
    x, y = np.meshgrid(np.arange(dw, dtype=np.float32), np.arange(dh, dtype=np.float32))
    return np.stack([(x + np.float32(0.5)) / np.float32(dw), (y + np.float32(0.5)) / np.float32(dh)], -1).reshape(-1, 2).astype(np.float32)


if __name__ == "__main__":
    from tests import llvmpipe_lib as LP
    lp = LP.instance()
    out = {"renderer": np.frombuffer((lp.renderer + " / " + lp.version).encode(), np.uint8)}
    for name, (kind, p) in RASTER_CASES.items():
        scene, cam, W, H, P = build_raster_case(kind, p)
        V, I = K.oracle_triangles(scene, scene.host_push_constants(cam))
        ids, depth = lp.raster(W, H, V, I)
        if P is not None:
            out[f"raster_{name}_positions"] = P
        out[f"raster_{name}_triangles_digest"] = triangles_digest(V, I)
        u = np.full(ids.shape, 0xFFFFFFFF, np.uint32)
        u[ids >= 0] = ids[ids >= 0].astype(np.uint32)
        out[f"raster_{name}_ids"] = u
        out[f"raster_{name}_depth"] = depth
    for (w, h) in SAMPLER_SIZES:
        uv = K.sampler_coords(np.random.default_rng(w * 131 + h), w, h, 1500)
        out[f"sampler_{w}x{h}_uv"] = uv
        out[f"sampler_{w}x{h}_classes"] = np.packbits(K.llvmpipe_footprint_classes(lp, w, h, uv))
    for (W, H) in HIZ_CHAINS:
        for (sw, sh), (dw, dh) in K.hiz_level_sizes(W, H):
            out[f"hiz_{W}x{H}_{sw}x{sh}_classes"] = np.packbits(K.llvmpipe_footprint_classes(lp, sw, sh, hiz_coords(dw, dh)))
    np.savez_compressed(PATH, **out)
    print(PATH, os.path.getsize(PATH), "bytes,", len(out), "arrays")

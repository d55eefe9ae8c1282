#!/usr/bin/env python
"""Context for the CPU baseline ("the CPU baseline is an untuned port ... the ratio would shrink against a tiled CPU rasteriser"):
how long does llvmpipe — Mesa's tiled, JIT-compiled, SIMD, multi-threaded rasteriser, what the reference's shaders would run on under lavapipe —
take for the RASTER stage of a BASELINE frame, beside the oracle's rasteriser on the same cores?

    python tools/llvmpipe_baseline.py [cfg1|cfg2|cfg3] [views]

Both rasterise exactly the triangles that survive the oracle's two-pass cull and the mesh shader's facing test for the view (llvmpipe gets them
as one client vertex array; its vertex fetch + pass-through vertex shader + clipper are inside its time, the oracle's mesh-shader arithmetic is
inside the oracle's).  Test infrastructure: prints one JSON line; nothing here is a bench.py number.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LP_NUM_THREADS", str(min(16, os.cpu_count() or 1)))

import numpy as np  # noqa: E402

from tests import llvmpipe_lib as LP  # noqa: E402
from tests import oracle_lib as O  # noqa: E402
from vk_gltf_viewer_b200.scene import Camera, Scene  # noqa: E402

CONFIGS = {
    "cfg1": (lambda: Scene.icosphere(57), (640, 480)),
    "cfg2": (lambda: Scene.atrium(128), (1920, 1080)),
    "cfg3": (lambda: Scene.lattice(10, 10, 10, 224, 0x5EED0003), (3840, 2160)),
}


def surviving_triangles(scene, pc, draw_ids):
    clip, cull = O.mesh_shader(pc, draw_ids)[:2]
    draws = scene.draws()
    prims, V = {}, []
    for k, d in enumerate(draw_ids):
        pi = int(draws[d]["primitiveIndex"])
        if pi not in prims:
            p = scene.primitive(pi)
            prims[pi] = (p["meshlets"], p["triangles"])
        ml = prims[pi][0][int(draws[d]["meshletIndex"])]
        tc, to = int(ml["triangleCount"]), int(ml["triangleOffset"])
        tri = prims[pi][1][to:to + 3 * tc].reshape(tc, 3)
        V.append(clip[k][tri[cull[k, :tc] == 0]].reshape(-1, 4))
    return np.concatenate(V) if V else np.zeros((0, 4), np.float32)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    views = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    make, (W, H) = CONFIGS[name]
    scene = make()
    lp = LP.instance()
    cam = Camera(W, H).look_at(*scene.default_view(0, 64))
    pc = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    O.frame(pc, tg, two_pass=True)                      # warm-up view: fills the pyramid
    rows = []
    for v in range(1, views + 1):
        cam.look_at(*scene.default_view(v, 64))
        stage = {}
        f = O.frame(pc, tg, two_pass=True, stage_s=stage)
        ids = np.concatenate([f["visibleA"], f["visibleB"]]).astype(np.uint32)
        V = surviving_triangles(scene, pc, ids)
        I = np.repeat(np.arange(V.shape[0] // 3, dtype=np.float32) % 1024, 3)
        lp.raster(W, H, V[:300], I[:300])               # JIT warm-up
        t = time.perf_counter()
        lids, _ = lp.raster(W, H, V, I)
        dt = time.perf_counter() - t
        diff = (lids >= 0) ^ (tg.ids_ref != 0xFFFFFFFF)
        ys, xs = np.nonzero(diff)
        # llvmpipe clips triangles that cross the viewport's border (and rounds the new vertices); everything else must agree
        border = int(((xs < 2) | (xs >= W - 2) | (ys < 2) | (ys >= H - 2)).sum())
        rows.append(dict(view=v, meshlets=int(ids.size), triangles=int(V.shape[0] // 3), llvmpipe_raster_ms=round(dt * 1e3, 1),
                         oracle_stage_ms={k: round(s * 1e3, 1) for k, s in stage.items()}, covered_pixels=int((lids >= 0).sum()), coverage_differs=int(diff.sum()), of_which_within_2px_of_the_border=border))
    print(json.dumps(dict(config=name, resolution=[W, H], cores=os.cpu_count(), lp_threads=int(os.environ["LP_NUM_THREADS"]), renderer=lp.renderer, views=rows)))


if __name__ == "__main__":
    main()

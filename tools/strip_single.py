"""strip-mode frame on ONE GPU (a world of one rank: the strip is the whole screen, nothing is pulled) — lets ncu look at
strip_merge_hiz_kernel's local phases.  usage: python tools/strip_single.py [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vk_gltf_viewer_b200 import api
from vk_gltf_viewer_b200.scene import Camera, Scene
W, H = 7680, 4320
scene = Scene.lattice(22, 22, 21, 224, 0x5EED0003)
views = [scene.default_view(i, 64) for i in range(8)]
cam = Camera(W, H).look_at(*views[0])
r = api.Renderer(W, H, device=0)
pc = r.upload_scene(scene, cam)
r.ipc_attach(0, [r.ipc_export()])
for k in range(5):
    cam.look_at(*views[k]); r.update_camera(pc, cam)
    st = r.frame(pc, api.FRAME_TWO_PASS | api.FRAME_MERGE_STRIPS | api.FRAME_TIMED | api.FRAME_STAGES)
    print(f"frame {k}: total {st.total_ms:.3f} mergeA {st.merge_a_ms:.3f} mergeB {st.merge_b_ms:.3f} hizA {st.hiz_a_ms:.3f} sent {st.strip_texels_sent}")
    st = r.frame(pc, api.FRAME_TWO_PASS | api.FRAME_TIMED | api.FRAME_STAGES)
    print(f"   plain: total {st.total_ms:.3f} hizA {st.hiz_a_ms:.3f} hizB {st.hiz_b_ms:.3f}")
r.ipc_detach(); r.close()

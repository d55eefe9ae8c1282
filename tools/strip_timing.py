import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vk_gltf_viewer_b200 import api, multigpu
from vk_gltf_viewer_b200.scene import Camera, Scene
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 7680, 4320
NF = int(os.environ.get('NF', '4'))
scene = Scene.lattice(22, 22, 21, 224, 0x5EED0003)
views = [scene.default_view(i, 64) for i in range(8)]
cam = Camera(W, H).look_at(*views[0])
r = api.Renderer(W, H, device=local)
pc = r.upload_scene(scene, cam)
r.set_shard_interleaved(rank, world, int(os.environ.get("VKV_SHARD_BLOCK_LOG2", "11")))
multigpu.attach_peers(r, dist)
lines = []
for k in range(NF):
    cam.look_at(*views[k]); r.update_camera(pc, cam)
    st = r.frame(pc, api.FRAME_TWO_PASS | api.FRAME_MERGE_STRIPS | api.FRAME_TIMED | api.FRAME_STAGES)
    lines.append(f"[rank {rank}] " +f"frame {k}: total {st.total_ms:.3f} cullA {st.cull_a_ms:.3f} rasterA {st.raster_a_ms:.3f} mergeA {st.merge_a_ms:.3f} cullB {st.cull_b_ms:.3f} rasterB {st.raster_b_ms:.3f} mergeB {st.merge_b_ms:.3f} pulled {st.strip_tiles_pulled} sent {st.strip_texels_sent}")
dist.barrier()
import time
time.sleep(0.05 * rank)
print("\n".join(lines[-2:]), file=sys.stderr)
dist.barrier(); r.ipc_detach(); r.close(); dist.destroy_process_group()

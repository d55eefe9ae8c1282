"""torchrun worker for tests/test_multigpu.py: meshlet-range-sharded two-pass frames on WORLD_SIZE GPUs must be
bit-identical to the single-list CPU oracle on every rank (visbuffer, pyramid) and the ranks' visible sets must
partition the oracle's."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests import oracle_lib as O  # noqa: E402
from vk_gltf_viewer_b200 import api, multigpu  # noqa: E402
from vk_gltf_viewer_b200.scene import Camera, Scene  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    W, H = (1000, 500) if mode.endswith("e2") else (1280, 720)   # 1000x500: two exact mips only, a ragged last tile row
    scene = Scene.lattice(4, 3, 4, 48)
    views = [scene.default_view(i, 24) for i in range(3)]
    cam = Camera(W, H).look_at(*views[0])
    r = api.Renderer(W, H, device=local)
    pc_dev = r.upload_scene(scene, cam)
    pc_host = scene.host_push_constants(cam)
    N = pc_host.meshletDrawCount
    if mode.endswith("interleaved") or mode in ("strips", "strips-e2"):   # blocks of 32 draws dealt round-robin (small scene: make the interleave real)
        r.set_shard_interleaved(rank, world, 5)
        mine = lambda ids: ids[multigpu.interleaved_owner(ids, world, 5) == rank]
    else:
        first, count = multigpu.shard_range(N, rank, world)
        r.set_shard(first, count)
        mine = lambda ids: ids[(ids >= first) & (ids < first + count)]
    multigpu.attach_peers(r, dist)
    tg = O.Targets(W, H)
    for k, (eye, center) in enumerate(views):
        if k:
            cam.look_at(eye, center)
            r.update_camera(pc_dev, cam)
        prev_pyramid = tg.pyramid.copy()
        out = O.frame(pc_host, tg, two_pass=True)
        if mode.startswith("strips"):
            # screen-strip ownership: the merged keys live in each owner's strip, the pyramid is complete on every rank
            st = r.frame(pc_dev, api.FRAME_TWO_PASS | api.FRAME_MERGE_STRIPS | api.FRAME_STATUS)
            rows = r.owned_rows(rank, world)
            own = r.read_visbuffer64()[rows]
            assert np.array_equal(own, tg.vis64()[rows]), f"rank {rank} view {k}: the rows this rank owns differ from the single-list oracle ({int((own != tg.vis64()[rows]).sum())} keys)"
            assert np.array_equal(r.read_pyramid().view(np.uint32), tg.pyramid.view(np.uint32)), f"rank {rank} view {k}: pyramid differs (strip mode)"
            # digests: the strip and the pyramid hash like the same data rendered by one GPU would (what bench.py's merge_parity compares)
            r1 = api.Renderer(W, H, device=local)
            pc1 = r1.upload_scene(scene, cam)
            r1._ck(r1.L.vkv_write_pyramid(r1.h, prev_pyramid.ctypes.data, prev_pyramid.size))
            r1.frame(pc1, api.FRAME_TWO_PASS)
            assert r1.hash(0, rank, world) == r.hash(0, rank, world) and r1.hash(1) == r.hash(1), f"rank {rank} view {k}: digests differ"
            assert r1.hash(0) != r.hash(0, rank, world)
            r1.close()
            r.gather_strips()   # whole image everywhere, for the comparison below
        elif mode.startswith("p2p"):
            st = r.frame(pc_dev, api.FRAME_TWO_PASS | api.FRAME_MERGE | api.FRAME_STATUS)
        else:  # library collective as the cross-check: stage by stage with ncclAllReduce(min) in place of vkv_merge
            r.clear()
            r.cull(pc_dev, 0, api.FRAME_STATUS); r.raster(pc_dev, 0); multigpu.nccl_min_merge(r, dist); r.hiz()
            r.cull(pc_dev, 1, api.FRAME_STATUS); r.raster(pc_dev, 1); multigpu.nccl_min_merge(r, dist); r.hiz()
            r.sync()
        vis = r.read_visbuffer64()
        assert np.array_equal(vis, tg.vis64()), f"rank {rank} view {k}: merged visbuffer differs from the single-list oracle ({int((vis != tg.vis64()).sum())} keys)"
        assert np.array_equal(r.read_pyramid().view(np.uint32), tg.pyramid.view(np.uint32)), f"rank {rank} view {k}: pyramid differs"
        assert np.array_equal(np.sort(r.read_visible(0)), mine(out["visibleA"])), f"rank {rank} view {k}: pass-A set"
        assert np.array_equal(np.sort(r.read_visible(1)), mine(out["visibleB"])), f"rank {rank} view {k}: pass-B set"
        nA = torch.tensor([r.read_visible(0).size, r.read_visible(1).size], device="cuda")
        dist.all_reduce(nA)
        assert nA.tolist() == [out["visibleA"].size, out["visibleB"].size]
    if mode == "p2p":
        # a resize detaches the peers (the visbuffer is reallocated); re-export + re-attach must start from a clean
        # barrier state (stale epochs in the flag slots would let every barrier pass at once) — the next merged frames must still
        # be bit-exact
        dist.barrier()
        r.resize(W, H)
        multigpu.attach_peers(r, dist)
        tg = O.Targets(W, H)   # vkv_resize restarts the pyramid from its cleared state
        cam = Camera(W, H).look_at(*views[0])
        r.update_camera(pc_dev, cam)
        pc_host = scene.host_push_constants(cam)
        for k, (eye, center) in enumerate(views[:2]):
            if k:
                cam.look_at(eye, center)
                r.update_camera(pc_dev, cam)
            O.frame(pc_host, tg, two_pass=True)
            r.frame(pc_dev, api.FRAME_TWO_PASS | api.FRAME_MERGE)
            assert np.array_equal(r.read_visbuffer64(), tg.vis64()), f"rank {rank}: merged visbuffer differs after resize + re-attach (view {k})"
            assert np.array_equal(r.read_pyramid().view(np.uint32), tg.pyramid.view(np.uint32))
    r.ipc_detach()
    r.close()
    dist.barrier()
    if rank == 0:
        print(f"mgpu ok: {world} ranks, mode {mode}, {N} draws, {len(views)} views bit-exact")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

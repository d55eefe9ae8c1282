"""Multi-GPU paths.

CPU (gloo, world_size 2): the host logic — range partition, view round-robin, IPC-handle exchange plumbing — and the
merge *semantics*: each rank rasterises its shard of the draw list with the oracle, an all_reduce(MIN) of the 64-bit keys
must equal the oracle's single-list image bit for bit (this is the property vkv_merge relies on).
GPU (needs >= 2 devices): tests/mgpu_worker.py under torchrun — the real NVLink peer-memory merge against the oracle, and
the NCCL all-reduce as a second implementation of the same merge.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from vk_gltf_viewer_b200 import multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_list():
    for n in (0, 1, 7, 102, 1059, 1490000, (1 << 25)):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                lo, cnt = multigpu.shard_range(n, r, world)
                assert lo == nxt and cnt >= 0
                nxt = lo + cnt
            assert nxt == n
            sizes = [multigpu.shard_range(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_range(10, 2, 2)


def test_interleaved_owner_covers_every_draw_once():
    ids = np.arange(100000, dtype=np.uint32)
    for world in (2, 3, 4, 8):
        own = multigpu.interleaved_owner(ids, world, 11)
        assert own.min() == 0 and own.max() == world - 1
        assert (own[:2048] == 0).all() and (own[2048:4096] == 1).all()
        counts = np.bincount(own.astype(np.int64), minlength=world)
        assert counts.sum() == ids.size and counts.max() - counts.min() <= 2048


def test_view_shard_round_robin():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in multigpu.view_shard(64, r, world))
        assert seen == list(range(64))
        assert all(len(multigpu.view_shard(64, r, world)) == 64 // world for r in range(world))


def test_merge_host_is_unsigned_min():
    a = np.array([0xFFFFFFFFFFFFFFFF, 0xC07FFFFF00000001, 5], np.uint64)
    b = np.array([0xC080000000000080, 0xC07FFFFF00000000, 0xFFFFFFFFFFFFFFFF], np.uint64)
    assert multigpu.merge_host([a, b]).tolist() == [0xC080000000000080, 0xC07FFFFF00000000, 5]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from tests import oracle_lib as O
    from vk_gltf_viewer_b200.scene import Camera, Scene
    try:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        # handle exchange plumbing (host bytes, rank order)
        mine = bytes([rank]) * 128
        got = multigpu.exchange_handles(mine, dist)
        assert got == [bytes([r]) * 128 for r in range(world)]
        # merge semantics on the oracle: shard -> raster -> all_reduce(MIN) == single-list image
        W, H = 320, 200
        scene = Scene.lattice(3, 2, 3, 24)
        cam = Camera(W, H).look_at(*scene.default_view(0, 8))
        pc = scene.host_push_constants(cam)
        N = pc.meshletDrawCount
        full = O.Targets(W, H)
        O.raster(pc, full, np.arange(N, dtype=np.uint32))
        first, count = multigpu.shard_range(N, rank, world)
        part = O.Targets(W, H)
        O.raster(pc, part, np.arange(first, first + count, dtype=np.uint32))
        keys = part.vis64().copy()
        assert (keys >> np.uint64(63)).all()  # every key (incl. clear) has the top bit set: signed min == unsigned min
        t = torch.from_numpy(keys.view(np.int64))
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        merged = t.numpy().view(np.uint64)
        assert np.array_equal(merged, full.vis64()), int((merged != full.vis64()).sum())
        # and the pyramid built from the merged depth equals the single-list pyramid
        depth = (~(merged >> np.uint64(32)).astype(np.uint32)).view(np.float32)
        pyr = np.zeros_like(full.pyramid)
        O.lib().orc_hiz(W, H, np.ascontiguousarray(depth).ctypes.data, pyr.ctypes.data, 1)
        O.hiz(full)
        assert np.array_equal(pyr.view(np.uint32), full.pyramid.view(np.uint32))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + repr(e) + "\n" + traceback.format_exc()))


def test_gloo_world2_shard_merge_semantics():
    import socket

    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(m == "ok" for _, m in res), res


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "nccl", "p2p-interleaved", "strips", "strips-contiguous", "strips-e2"])
def test_range_sharded_frames_match_single_list_oracle(mode):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "mgpu ok" in out.stdout


@pytest.mark.gpu
def test_two_contexts_on_two_devices_in_one_process():
    """independent views need no more than one context per GPU; a single process may hold several (vkv_create's
    cuda_device argument).  Per-device state (function attributes, occupancy) must not leak between them: both render the same
    two-pass frames bit-exactly, interleaved."""
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    from tests import oracle_lib as O
    from tests import scenes as S
    from vk_gltf_viewer_b200 import api
    from vk_gltf_viewer_b200.scene import Camera
    W, H = 640, 480
    scene = S.occluder_and_hidden()
    cam = Camera(W, H).look_at((0, 0, 8), (0, 0, 0))
    rs = [api.Renderer(W, H, device=d) for d in (0, 1)]
    pcs = [r.upload_scene(scene, cam) for r in rs]
    pc_host = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    for eye in ((0, 0, 8), (0.3, 0.1, 8), (0.6, 0.0, 7.5)):
        cam.look_at(eye, (0, 0, 0))
        out = O.frame(pc_host, tg, two_pass=True)
        for r, pc in zip(rs, pcs):
            r.update_camera(pc, cam)
            r.frame(pc, api.FRAME_TWO_PASS)
        for r in rs:
            assert np.array_equal(np.sort(r.read_visible(0)), out["visibleA"])
            assert np.array_equal(r.read_visbuffer64(), tg.vis64())
            assert np.array_equal(r.read_pyramid().view(np.uint32), tg.pyramid.view(np.uint32))
    for r in rs:
        r.close()

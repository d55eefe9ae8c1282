"""Frames in flight (vkv_update_staged / vkv_frame_submit / vkv_frame_wait) against blocking vkv_frame calls and the CPU oracle.

The reference keeps Application::frameOverlap = 3 frames in flight (application.hpp:146), each with its own camera buffer
(camera.cpp:86) and a fence waited for before the slot is reused (application.cpp:642-660); frame k still culls against frame
k-1's pyramid.  Submitting without waiting must therefore change nothing observable: same counters per frame, same bits at the end.
"""
import numpy as np
import pytest

from tests import oracle_lib as O
from vk_gltf_viewer_b200 import api
from vk_gltf_viewer_b200.scene import Camera, Scene

pytestmark = pytest.mark.gpu

FO = 3


def _views(scene, n):
    return [scene.default_view(i, 24) for i in range(n)]


@pytest.mark.parametrize("two_pass", [False, True])
def test_frames_in_flight_equal_blocking_frames_and_the_oracle(two_pass):
    import torch
    W, H = 640, 480
    scene = Scene.icosphere(40)
    views = _views(scene, 8)
    flags = api.FRAME_TWO_PASS if two_pass else api.FRAME_ONE_PASS
    cam = Camera(W, H)
    cam.look_at(*views[0])

    # oracle + blocking frames
    r = api.Renderer(W, H)
    pc = r.upload_scene(scene, cam)
    pc_host = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    want = []
    for k, v in enumerate(views):
        cam.look_at(*v)
        r.update_camera(pc, cam)
        out = O.frame(pc_host, tg, two_pass=two_pass)
        st = r.frame(pc, flags)
        assert st.visible_a == out["visibleA"].size
        want.append((st.visible_a, st.occluded_a, st.visible_b, st.tested_b, st.kernel_launches))
    vis_blocking, pyr_blocking = r.read_visbuffer64(), r.read_pyramid()
    assert np.array_equal(vis_blocking, tg.vis64())
    r.close()

    # the same sweep, three frames in flight, per-slot camera and transform buffers fed from pinned memory
    r = api.Renderer(W, H)
    cam.look_at(*views[0])
    pc = r.upload_scene(scene, cam)
    transforms = np.ascontiguousarray(scene.transforms())
    pin_tr = torch.from_numpy(transforms.reshape(-1).copy()).pin_memory()
    pin_cam = [torch.zeros(352, dtype=torch.uint8).pin_memory() for _ in range(FO)]
    slot_cam = [r.upload(np.zeros(352, np.uint8)) for _ in range(FO)]
    slot_tr = [r.upload(np.zeros(transforms.nbytes, np.uint8)) for _ in range(FO)]
    tickets, got = [0] * FO, []
    for k, v in enumerate(views):
        sl = k % FO
        if tickets[sl]:
            st = r.frame_wait(tickets[sl])
            got.append((st.visible_a, st.occluded_a, st.visible_b, st.tested_b, st.kernel_launches))
        cam.look_at(*v)
        pin_cam[sl].numpy()[:] = np.frombuffer(cam.raw(), np.uint8)
        r.update_staged(slot_cam[sl], pin_cam[sl].data_ptr(), 352)
        r.update_staged(slot_tr[sl], pin_tr.data_ptr(), transforms.nbytes)
        pc.cameraBuffer, pc.transformBuffer = slot_cam[sl], slot_tr[sl]
        tickets[sl] = r.frame_submit(pc, flags)
    order = [(len(views) - FO + i) % FO for i in range(FO)]  # oldest outstanding slot first
    for sl in order:
        st = r.frame_wait(tickets[sl])
        got.append((st.visible_a, st.occluded_a, st.visible_b, st.tested_b, st.kernel_launches))
    assert got == want
    assert np.array_equal(r.read_visbuffer64(), vis_blocking)
    assert np.array_equal(r.read_pyramid().view(np.uint32), pyr_blocking.view(np.uint32))
    r.close()


def test_frames_in_flight_limits_and_errors():
    W, H = 128, 96
    scene = Scene.icosphere(6)
    cam = Camera(W, H)
    cam.look_at(*scene.default_view(0, 24))
    r = api.Renderer(W, H)
    pc = r.upload_scene(scene, cam)
    with pytest.raises(api.VkvError) as e:   # timed frames are blocking frames
        r.frame_submit(pc, api.FRAME_TIMED)
    assert e.value.code == -2
    with pytest.raises(api.VkvError):        # nothing in flight yet
        r.frame_wait(7)
    t = [r.frame_submit(pc, api.FRAME_TWO_PASS) for _ in range(4)]
    assert len(set(t)) == 4 and 0 not in t
    with pytest.raises(api.VkvError) as e:   # the ring is four deep
        r.frame_submit(pc, api.FRAME_TWO_PASS)
    assert e.value.code == -5
    first = r.frame_wait(t[0])
    with pytest.raises(api.VkvError):        # a ticket is released by its wait
        r.frame_wait(t[0])
    t.append(r.frame_submit(pc, api.FRAME_TWO_PASS))
    rest = [r.frame_wait(x) for x in t[1:]]
    assert all(s.draws == first.draws and s.visible_a > 0 for s in rest)
    # a blocking frame after the queue has drained sees the same steady state
    assert r.frame(pc, api.FRAME_TWO_PASS).visible_a == rest[-1].visible_a
    r.close()

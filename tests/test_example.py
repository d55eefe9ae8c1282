"""examples/headless.cpp — the reference's frame loop (application.cpp:620-1010) against the C ABI, in the reference's language.

CPU: it builds against include/vkv.h + include/vkv_host.h, and without a CUDA device it fails loudly (exit code 3, no CPU fallback).
GPU: it renders the BASELINE cfg-1 scene with three frames in flight and writes the resolved image.
"""
import os
import subprocess

import numpy as np
import pytest

from .conftest import ROOT, has_gpu

EXE = os.path.join(ROOT, "examples", "headless")


def _build():
    subprocess.check_call(["make", "-C", ROOT, "examples"], stdout=subprocess.DEVNULL)
    assert os.path.exists(EXE)


def test_headless_example_builds_and_refuses_to_run_without_a_gpu():
    if not os.path.exists(os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv.so")):
        pytest.skip("libvkv.so not built (run `make` / __graft_entry__.build())")
    _build()
    if has_gpu():
        pytest.skip("a GPU is present: the run is covered by the gpu test")
    p = subprocess.run([EXE, "--scene", "icosphere", "--frames", "2"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stderr)
    assert "no CPU fallback" in p.stderr
    assert "677 meshlet draws" in p.stdout          # the host side (scene + meshlets) ran before the device was asked for


def test_headless_example_rejects_unknown_arguments():
    if not os.path.exists(os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv.so")):
        pytest.skip("libvkv.so not built")
    _build()
    p = subprocess.run([EXE, "--no-such-flag"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 2 and "unknown argument" in p.stderr
    p = subprocess.run([EXE, "--asset", "/nonexistent/file.glb"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 2 and "file.glb" in p.stderr


@pytest.mark.gpu
@pytest.mark.late   # ran green on the B200 before the decoder hook was added to the example; the present binary has not run on a device
def test_headless_example_renders(tmp_path):
    _build()
    out = tmp_path / "frame.ppm"
    p = subprocess.run([EXE, "--scene", "icosphere", "--size", "640x480", "--frames", "7", "--out", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert "7 frames" in p.stdout
    raw = out.read_bytes()
    head = b"P6\n640 480\n255\n"
    assert raw.startswith(head) and len(raw) == len(head) + 640 * 480 * 3
    img = np.frombuffer(raw[len(head):], np.uint8).reshape(480, 640, 3)
    # the sphere is in view: a background colour plus several material colours
    assert len(np.unique(img.reshape(-1, 3), axis=0)) >= 2


def test_headless_example_sends_compressed_views_to_the_device_decoder(tmp_path, meshopt_ref):
    """an EXT_meshopt_compression asset: the example installs libvkv's device decoder behind the host reader's hook, so without a GPU the load
    fails at the first compressed bufferView (and says so) instead of falling back to a CPU decoder"""
    if not os.path.exists(os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv.so")):
        pytest.skip("libvkv.so not built")
    if has_gpu():
        pytest.skip("a GPU is present")
    from tests import meshopt_lib as M
    from tests.test_gltf import _compressed_and_plain_assets
    _build()
    comp, _ = _compressed_and_plain_assets(M)
    path = tmp_path / "compressed.glb"
    path.write_bytes(comp)
    p = subprocess.run([EXE, "--asset", str(path), "--frames", "1"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "did not decode" in p.stderr, (p.returncode, p.stderr)

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "late: collected last (see pytest_collection_modifyitems)")
    # host library + oracle are plain g++ builds: make sure they exist for the CPU suite
    need = [os.path.join(ROOT, "vk_gltf_viewer_b200", "libvkv_host.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT, "vk_gltf_viewer_b200/libvkv_host.so", "oracle/liboracle.so"])


def pytest_collection_modifyitems(config, items):
    """tests marked `late` run after everything else: GPU tests added after the round's GPU budget was spent (never run on a device yet)
    must not stand in front of the measured-green ones under `pytest -x`"""
    late = [i for i in items if i.get_closest_marker("late")]
    if late:
        items[:] = [i for i in items if not i.get_closest_marker("late")] + late


def has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int()
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def ref_shim():
    """oracle/_ref/libref_shim.so — the reference's own C++-compilable pieces (built here by oracle/build_ref.sh)."""
    import ctypes
    p = os.path.join(ROOT, "oracle", "_ref", "libref_shim.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["sh", os.path.join(ROOT, "oracle", "build_ref.sh")])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return ctypes.CDLL(p)


@pytest.fixture(scope="session")
def meshopt_ref():
    import ctypes
    p = os.path.join(ROOT, "oracle", "_ref", "libmeshopt_ref.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["sh", os.path.join(ROOT, "oracle", "build_ref.sh")])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return ctypes.CDLL(p)

"""Loads the in-tree native libraries.  There is NO fallback: a missing library is a hard error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)


class NativeLibraryMissing(RuntimeError):
    pass


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} is not built. Run `make` (or `python -c 'import __graft_entry__ as g; g.build()'`) in {ROOT}. "
            "There is no CPU fallback for the CUDA path."
        )
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


_host = None
_vkv = None


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        _host = _load("libvkv_host.so")
    return _host


def vkv_lib() -> C.CDLL:
    global _vkv
    if _vkv is None:
        # VKV_LIBVKV: load another in-tree build of the SAME library (kernel tuning variants); still no fallback
        _vkv = _load(os.environ.get("VKV_LIBVKV", "libvkv.so"))
    return _vkv

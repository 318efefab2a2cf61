"""The C-ABI library loads and exports every symbol include/voxcore_gpu.h declares; without a
CUDA device the product path fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

from voxel_ma_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "voxcore_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(api.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    for name in _declared():
        assert hasattr(lib, name), f"{name} is declared in include/voxcore_gpu.h but not exported"
    assert lib.vc_abi_version() == 1


def _no_gpu():
    lib = api.load_library()
    h = ctypes.c_void_p()
    st = lib.vc_ctx_create(0, ctypes.byref(h))
    if st == 0:
        lib.vc_ctx_destroy(h)
    return st != 0


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_no_device_means_error_not_fallback():
    lib = api.load_library()
    h = ctypes.c_void_p()
    assert lib.vc_ctx_create(0, ctypes.byref(h)) == -2  # VC_ERR_CUDA
    assert not h.value
    with pytest.raises(api.VoxcoreError):
        api.Context(0)


def test_missing_library_raises():
    with pytest.raises(api.VoxcoreError):
        api.load_library("/nonexistent/libvoxcore_gpu.so")


def test_null_context_is_rejected():
    lib = api.load_library()
    assert lib.vc_classify_grid(None, None) == -1
    assert lib.vc_closest_grid(None, None, None) == -1
    assert lib.vc_last_error(None) == b"null context"


def test_product_package_never_touches_the_oracle():
    """the product path must not import, link or dlopen anything under oracle/"""
    pkg = os.path.join(ROOT, "voxel_ma_b200")
    for dp, _, files in os.walk(pkg):
        if os.sep + "lib" in dp or os.sep + "_build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".hpp")):
                for line in open(os.path.join(dp, f), errors="ignore"):
                    low = line.lower()
                    assert not ("oracle" in low and ("import" in low or "#include" in low or "cdll" in low)), (f, line)
                    assert "liboracle" not in low and "libvoxref" not in low, (f, line)
            if f.startswith("Makefile") or f.endswith((".mk", ".sh")):
                # build recipes too: the drop-in CLI compiles the reference's host objects itself (host/Makefile.dropin),
                # it takes nothing out of the checker tree
                for line in open(os.path.join(dp, f), errors="ignore"):
                    code = line.split("#", 1)[0]
                    assert "oracle/" not in code and "oracle\\" not in code, (f, line)

"""Build libvoxcore_gpu.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m voxel_ma_b200.build [--force]

The shared library lands in voxel_ma_b200/lib/ (git-ignored, but it travels with a gpurun
snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libvoxcore_gpu.so")
SOURCES = ["vc_api.cu", "vc_sites.cu", "vc_edt.cu", "vc_measures.cu", "vc_points.cu", "vc_mesh.cu", "vc_thin.cu", "vc_peer.cu", "vc_compact.cu", "vc_medial.cu", "vc_circum.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # float32 measures must match the reference bit for bit: no FMA contraction
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas", "-Xptxas", "-v",
]


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdr.append(os.path.join(os.path.dirname(HERE), "include", "voxcore_gpu.h"))
    return hdr


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src, objdir=None, defines=()):
    obj = os.path.join(objdir or OBJDIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    r = subprocess.run([NVCC, *FLAGS, *defines, "-c", path, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if log:
        with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


DROPIN_CLI = os.path.join(HERE, "host", "_build", "main_voroUtility_gpu")


def build_dropin_cli(reference_tree: str = "/root/reference"):
    """The reference CLI linked against host/dropin/*.cpp + libvoxcore_gpu.so, and the dense-core tool
    (host/Makefile.dropin: it compiles the reference's host objects from the reference tree itself, under
    host/_build).  Needs the reference tree: possible in the build container only; the binaries then travel
    with the snapshot.  Returns the CLI's path or None."""
    root = os.path.dirname(HERE)
    if not os.path.isdir(reference_tree):
        return DROPIN_CLI if os.path.exists(DROPIN_CLI) else None
    r = subprocess.run(["make", "-j8", "-f", os.path.join("voxel_ma_b200", "host", "Makefile.dropin"), f"REF={reference_tree}"],
                       cwd=root, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"drop-in CLI build failed:\n{r.stdout}\n{r.stderr}")
    return DROPIN_CLI


def build_variant(name: str, defines) -> str:
    """Development aid: the same sources with extra -D flags -> lib/variants/libvoxcore_gpu_<name>.so"""
    vdir = os.path.join(LIBDIR, "variants", name)
    os.makedirs(vdir, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, vdir, tuple(defines)), SOURCES))
    lib = os.path.join(LIBDIR, "variants", f"libvoxcore_gpu_{name}.so")
    r = subprocess.run([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", lib, *[o for o, _ in res], "-lcudart_static", "-ldl", "-lrt",
                        "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

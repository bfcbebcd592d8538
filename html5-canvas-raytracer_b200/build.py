"""In-tree build of csrc/librtb200.so (and the header-free N-API shim) with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the built .so
travels to the GPU box with the repo snapshot.  Incremental on mtimes.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "librtb200.so")
NAPI = os.path.join(CSRC, "rtb200.node")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-I", INCLUDE]
# per-TU extra flags: the float64 validation kernel must not contract a*b+c (JS rounds per operator)
SOURCES = {
    "abi.cu": [],
    "pixel_f64.cu": ["-fmad=false"],
    "pixel_f32.cu": [],
    "wavefront.cu": [],
    "peer.cu": [],
    "host_prep.cu": ["-Xcompiler", "-ffp-contract=off"],   # float64 host code, one rounding per JS operator
}
HEADERS = ["rtb_device.cuh", "rtb_internal.h", "pixel_kernel.cuh", "pixel_launch.inl",
           os.path.join(INCLUDE, "rtb200.h")]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        sys.stdout.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}")
    return r.stdout


def build_library(force: bool = False, verbose: bool = False, ptxas_v: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = max(_mtime(h if os.path.isabs(h) else os.path.join(CSRC, h)) for h in HEADERS)
    hdr_time = max(hdr_time, _mtime(os.path.abspath(__file__)))
    objs, relink = [], force or not os.path.exists(LIB)
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _mtime(o) < max(_mtime(s), hdr_time):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            if ptxas_v:
                cmd[1:1] = ["-Xptxas", "-v"]
            out = _run(cmd, verbose)
            if ptxas_v and not verbose:
                sys.stdout.write(out)
            relink = True
    if relink or any(_mtime(o) > _mtime(LIB) for o in objs):
        _run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-ldl"], verbose)   # peer.cu: dlopen/dlsym (glibc < 2.34)
    return LIB


def build_napi(force: bool = False, verbose: bool = False) -> str:
    """The thin Node-API addon (csrc/napi_shim.cc).  Header-free: napi_* symbols stay undefined and
    are bound by the `node` executable when it dlopen()s the .node file."""
    src = os.path.join(CSRC, "napi_shim.cc")
    if not os.path.exists(src):
        return ""
    if force or _mtime(NAPI) < max(_mtime(src), _mtime(os.path.join(INCLUDE, "rtb200.h"))):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        _run([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", INCLUDE, src, "-o", NAPI,
              "-L", CSRC, "-l:librtb200.so", "-Wl,-rpath,$ORIGIN"], verbose)
    return NAPI


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_v="--ptxas" in sys.argv))
    print(build_napi(force="--force" in sys.argv, verbose="-v" in sys.argv))

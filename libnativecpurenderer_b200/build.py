"""In-tree build of the product library (nvcc, sm_100a only) and of the test oracles.

    python -m libnativecpurenderer_b200.build            # product + oracles
    python -m libnativecpurenderer_b200.build --product  # product only

Output: libnativecpurenderer_b200/lib/libNativeCPURenderer.so — the file name the reference
binding loads from its working directory (reference src/libNativeCPURendererPybind.py:9).
The library is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libNativeCPURenderer.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["api.cu", "kernels.cu", "composite.cu", "host_misc.cpp", "batch.cpp"]
# -fmad=false: the reference's f64 expression trees must not be contracted into FMAs on the device;
# -ffp-contract=off keeps the host-side per-call math (state.h) uncontracted as well.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_variant(name: str, defines: list[str]) -> str:
    """Development helper: the product compiled with extra -D flags into lib/variants/<name>/ (for A/B timing)."""
    out_dir = os.path.join(LIBDIR, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.rsplit(".", 1)[0] + ".o")
        extra = os.environ.get("NCR_EXTRA_NVCC", "").split()   # development: extra compiler flags for an A/B build
        cmd = [NVCC, *NVCC_FLAGS, *extra, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        if src == "composite.cu":
            with open(os.path.join(out_dir, "ptxas_composite.log"), "w") as f:
                f.write(res.stderr)
        objs.append(obj)
    lib = os.path.join(out_dir, "libNativeCPURenderer.so")
    res = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lpthread"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return lib


def build_product(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ncr_b200.h"), __file__]
    if not force and _newer(LIB, deps):
        return LIB
    objs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        if src.endswith(".cu") and src != "api.cu":
            with open(os.path.join(LIBDIR, f"ptxas_{src[:-3]}.log"), "w") as f:
                f.write(res.stderr)
        objs.append(obj)
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return LIB


REPLAY_LIB = os.path.join(LIBDIR, "libncr_replay.so")


def build_replayer(force: bool = False) -> str:
    """Host-side trace replayer (csrc/ncr_replay.cpp): a plain C++ caller of the reference C ABI through dlopen, used by
    bench.py's end-to-end measurement and by tests to drive any of the libraries without per-call FFI cost."""
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "ncr_replay.cpp")
    if not force and _newer(REPLAY_LIB, [src, os.path.join(CSRC, "ncr_trace.h")]):
        return REPLAY_LIB
    cmd = [os.environ.get("CXX", "g++"), "-shared", "-fPIC", "-O2", "-std=c++17", "-Wall", "-o", REPLAY_LIB, src, "-ldl", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("replayer build failed")
    return REPLAY_LIB


def build_oracles() -> None:
    """CPU checkers under oracle/ (test infrastructure; building them is not using them)."""
    res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("oracle build failed")


def main(argv: list[str]) -> int:
    build_product(force="--force" in argv, verbose="-v" in argv)
    build_replayer(force="--force" in argv)
    if "--product" not in argv:
        build_oracles()
    print(LIB)
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))

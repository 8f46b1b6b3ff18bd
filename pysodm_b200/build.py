"""In-tree build of the CUDA library (`pysodm_b200/libsodm.so`) for sm_100a.

`python -m pysodm_b200.build` compiles every `csrc/*.cu` with nvcc (cross-compiles without a GPU) and
links one shared library whose exported symbols are exactly the C ABI of `include/sodm.h`.  The built
`.so` is git-ignored but travels with the working tree to the GPU box.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ_DIR = os.path.join(PKG, "csrc", "_obj")
LIB_PATH = os.path.join(PKG, "libsodm.so")
HOSTCHECK_SRC = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cu")
HOSTCHECK_LIB = os.path.join(ROOT, "tests", "hostcheck", "libsodm_hostcheck.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing",
          "-Werror", "cross-execution-space-call"]      # a host constexpr called from device code silently misbehaves


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "sodm.h"))
    return sorted(hs)


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _run(cmd):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + proc.stdout)
    if "warning" in proc.stdout and "-Xptxas" not in cmd:
        sys.stderr.write(proc.stdout)
    return proc.stdout


def _compile_one(src, flags, tag):
    os.makedirs(OBJ_DIR, exist_ok=True)
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + tag + ".o")
    stamp = obj + ".sha"
    want = _digest([src] + _headers(), " ".join(flags))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    _run([NVCC] + ARCH + COMMON + flags + ["-c", src, "-o", obj])
    with open(stamp, "w") as f:
        f.write(want)
    return obj


def build_library(verbose=False, force=False):
    """Compile csrc/*.cu and link pysodm_b200/libsodm.so.  Returns the library path."""
    srcs = _sources()
    want = _digest(srcs + _headers(), "lib")
    stamp = LIB_PATH + ".sha"
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB_PATH
    flags = ["-Xptxas", "-v"] if verbose else []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, flags, ""), srcs))
    _run([NVCC] + ARCH + ["-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"])
    with open(stamp, "w") as f:
        f.write(want)
    return LIB_PATH


def build_hostcheck():
    """Test-only CPU harness around the same __host__ __device__ code (tests/hostcheck)."""
    srcs = [HOSTCHECK_SRC, os.path.join(CSRC, "sodm_core.cu")]
    want = _digest(srcs + _headers(), "hostcheck")
    stamp = HOSTCHECK_LIB + ".sha"
    if os.path.exists(HOSTCHECK_LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return HOSTCHECK_LIB
    _run([NVCC] + ARCH + COMMON + ["-Xcompiler", "-fopenmp", "-Xcompiler", "-ffp-contract=off", "-shared", "-o",
                                   HOSTCHECK_LIB] + srcs + ["-cudart", "static", "-lgomp"])
    with open(stamp, "w") as f:
        f.write(want)
    return HOSTCHECK_LIB


if __name__ == "__main__":
    print(build_library(verbose="-v" in sys.argv, force="-f" in sys.argv))
    if "--hostcheck" in sys.argv:
        print(build_hostcheck())

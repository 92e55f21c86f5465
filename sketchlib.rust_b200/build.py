"""Builds lib/libsketchdist.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libsketchdist.so")
SOURCES = ["kernels.cu", "capi.cu", "peer.cu", "sortknn.cu"]
HEADERS = ["kernels.h", "layout.h", "internal.h", os.path.join("..", "..", "include", "sketchdist.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-cudart", "static",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsketchdist cannot be built")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compiles every source to an object (in parallel, only the stale ones) and links the shared library."""
    out = out or LIB
    if not force and not extra_flags and out == LIB and not stale():
        return LIB
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.dirname(out), exist_ok=True)
    # several ranks of one job may find the library stale at the same moment: one builds, the others wait and re-check
    lock = open(os.path.join(os.path.dirname(out), ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not extra_flags and out == LIB and not stale():
            return LIB
        return _build_locked(force, verbose, extra_flags, out)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(force, verbose, extra_flags, out):
    from concurrent.futures import ThreadPoolExecutor
    tag = "" if not extra_flags else "_" + "".join(c if c.isalnum() else "_" for c in " ".join(extra_flags))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = find_nvcc()
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + tag + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(os.path.join(CSRC, src))):
            subprocess.check_call([nvcc] + flags + ["-c", "-o", obj, src], cwd=CSRC)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", out] + objs)
    return out


BIN = os.path.join(HERE, "bin")
HOST = os.path.join(CSRC, "host")
HOST_FLAGS = ["-O2", "-std=c++17", "-Wall", "-Wextra", "-ffp-contract=off"]


HOSTLIB = os.path.join(HERE, "lib", "libsketchhost.so")


def build_hostlib(force=False):
    """Host-only shared library over the C++ readers (.skm / .ski / completeness files) for the Python layer."""
    srcs = ["host_capi.cpp", "multisketch.cpp", "inverted.cpp", "sketcher.cpp"]
    deps = [os.path.join(HOST, f) for f in srcs + ["multisketch.hpp", "inverted.hpp"]]
    newest = max(os.path.getmtime(d) for d in deps)
    if force or not os.path.exists(HOSTLIB) or os.path.getmtime(HOSTLIB) < newest:
        os.makedirs(os.path.dirname(HOSTLIB), exist_ok=True)
        subprocess.check_call(["g++"] + HOST_FLAGS + ["-shared", "-fPIC", "-o", HOSTLIB] + srcs + ["-lz", "-lpthread"], cwd=HOST)
    return HOSTLIB


def build_host(force=False):
    """Host-only C++: the drop-in `sketchlib dist` CLI (links libsketchdist.so), the skm_info test tool and
    libsketchhost.so."""
    os.makedirs(BIN, exist_ok=True)
    build_hostlib(force)
    cli, info = os.path.join(BIN, "sketchlib"), os.path.join(BIN, "skm_info")
    srcs = [os.path.join(HOST, f) for f in ("dist_main.cpp", "multisketch.cpp", "multisketch.hpp", "skm_info.cpp",
                                            "sketcher.cpp", "inverted.cpp", "inverted.hpp")]
    srcs.append(os.path.join(HERE, "..", "include", "sketchdist.h"))
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(cli) or os.path.getmtime(cli) < max(newest, os.path.getmtime(LIB)):
        subprocess.check_call(["g++"] + HOST_FLAGS + ["-o", cli, "dist_main.cpp", "multisketch.cpp", "sketcher.cpp", "inverted.cpp", "-lz",
                                                    "-L", os.path.dirname(LIB), "-lsketchdist",
                                                    "-Wl,-rpath,$ORIGIN/../lib", "-ldl", "-lpthread", "-lrt"], cwd=HOST)
    if force or not os.path.exists(info) or os.path.getmtime(info) < newest:
        subprocess.check_call(["g++"] + HOST_FLAGS + ["-o", info, "skm_info.cpp", "multisketch.cpp"], cwd=HOST)
    return cli, info


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))

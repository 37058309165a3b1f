"""Builds libmgpu.so (hand-written sm_100a kernels + C++ host + C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmgpu.so")

CUDA_SOURCES = ["cuda/kernels.cu"]
HOST_SOURCES = ["host/index_format.cpp", "host/engine.cpp", "host/sharded.cpp", "host/api.cpp", "host/query_parser.cpp", "host/index_check.cpp", "host/api_wire.cpp"]
# the index writer + synthetic corpus: a host-only library of its own (no CUDA), so that the CPU arm of bench.py and the
# golden-corpus builders never map libmgpu.so
WRITER_LIB = os.path.join(HERE, "libmgpu_writer.so")
WRITER_SOURCES = ["host/index_format.cpp", "host/index_writer.cpp", "host/writer_api.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # TF*IDF arithmetic must match the reference's scalar fp32
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function,-ffp-contract=off",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES]


def _writer_sources():
    return [os.path.join(CSRC, s) for s in WRITER_SOURCES]


def _deps():
    deps = list(_sources()) + [os.path.join(CSRC, "host/index_writer.cpp"), os.path.join(CSRC, "host/writer_api.cpp")]
    for root, _, files in os.walk(CSRC):
        deps += [os.path.join(root, f) for f in files if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "mgpu.h"))
    deps.append(os.path.join(HERE, "..", "include", "mgpu_writer.h"))
    return deps


STAMP = LIB + ".stamp"


def _digest():
    import hashlib
    h = hashlib.sha1()
    for d in sorted(_deps()):
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    """content based (mtimes do not survive the snapshot copy to the GPU box)"""
    if not os.path.exists(LIB) or not os.path.exists(WRITER_LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != _digest()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    import fcntl
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    with open(os.path.join(HERE, "_obj", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)       # several ranks may get here at once
        if force or needs_build():
            _build_locked(verbose)
    return LIB


def _build_locked(verbose):
    nvcc = _nvcc()
    objs = []
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu" if src.endswith(".cu") else "c++", "-c", src, "-o", obj]
        if not src.endswith(".cu"):
            # host files only need cuda_runtime.h; compile them as plain C++ through nvcc's host compiler
            cmd = [nvcc] + NVCC_FLAGS + ["-x", "c++", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-cudart", "static", "-Xlinker", "--no-undefined", "-o", LIB + ".tmp"] + objs + ["-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    os.replace(LIB + ".tmp", LIB)
    # the writer library: plain g++ through nvcc's host compiler, no CUDA runtime
    wobjs = []
    for src in _writer_sources():
        obj = os.path.join(objdir, "w_" + os.path.basename(src) + ".o")
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", "-ffp-contract=off", "-I/usr/local/cuda/include",
                            "-c", src, "-o", obj], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("g++ failed on %s" % src)
        wobjs.append(obj)
    r = subprocess.run(["g++", "-shared", "-Wl,--no-undefined", "-o", WRITER_LIB + ".tmp"] + wobjs + ["-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libmgpu_writer.so failed")
    os.replace(WRITER_LIB + ".tmp", WRITER_LIB)
    open(STAMP, "w").write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

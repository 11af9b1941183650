"""In-tree build of libtsp_b200.so (hand-written CUDA for sm_100a + the C ABI of include/tsp.h).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libtsp_b200.so")
SOURCES = [os.path.join(CSRC, "tsp_engine.cu")]
HEADERS = [os.path.join(CSRC, h) for h in ("tsp_device.cuh", "tsp_fast.cuh", "tsp_tc.cuh", "tsp_stream.cuh", "tsp_rootio.cuh")] + [
    os.path.join(os.path.dirname(_HERE), "include", "tsp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the engine for sm_100a; returns the path of the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    # one builder at a time (one process per GPU may import the package at once); the library appears atomically, so a
    # concurrent dlopen never sees a torn file
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB_PATH  # another process built it while we waited
        tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
        cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
        try:
            proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
        except FileNotFoundError:
            raise RuntimeError("nvcc not found (looked for $NVCC, /usr/local/cuda/bin/nvcc, nvcc on PATH): cannot build %s" % LIB_PATH)
        if proc.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
        os.replace(tmp, LIB_PATH)
        if verbose:
            print(proc.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

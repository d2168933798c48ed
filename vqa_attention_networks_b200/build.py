"""Build the C-ABI CUDA library (csrc/*.cu -> csrc/libvqa_b200.so) for sm_100a with nvcc.

The library is built in-tree so that it travels to the GPU box with the repository snapshot.
nvcc cross-compiles sm_100a without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import contextlib
import fcntl
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libvqa_b200.so")
SOURCES = ["gemm.cu", "kernels_misc.cu", "lstm.cu", "optim.cu"]
HEADERS = ["ptx.cuh", "gemm_sm100.cuh", "common.h", os.path.join("..", "..", "include", "vqa_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    return None


def up_to_date() -> bool:
    if not os.path.isfile(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return all(os.path.getmtime(d) <= t for d in deps if os.path.isfile(d))


@contextlib.contextmanager
def _build_lock():
    """Inter-process lock: under torchrun every rank may find the library stale at once (the .so is git-ignored); only
    one of them may run nvcc into csrc/build, the others wait and then find the library up to date."""
    fd = os.open(os.path.join(CSRC, ".build.lock"), os.O_CREAT | os.O_RDWR, 0o644)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        try:
            fcntl.flock(fd, fcntl.LOCK_UN)
        finally:
            os.close(fd)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link libvqa_b200.so.  Returns the library path."""
    if not force and up_to_date():
        return LIB_PATH
    with _build_lock():
        if not force and up_to_date():          # another process built it while this one waited for the lock
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libvqa_b200.so (set NVCC or install the CUDA toolkit)")
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    pid = os.getpid()

    # VQA_B200_DEBUG=1: instrumented kernels + the process-global debug hooks (include/vqa_b200.h, last section)
    debug = ["-DVQA_B200_DEBUG"] if os.environ.get("VQA_B200_DEBUG", "0") == "1" else []
    if os.environ.get("VQA_B200_LSTM_STRICT_BARRIER", "0") == "1":     # release / acquire on the recurrence's cluster barrier
        debug.append("-DVQA_B200_LSTM_STRICT_BARRIER")

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, "%s.%d.o" % (os.path.splitext(src)[0], pid))
        cmd = [nvcc, *NVCC_FLAGS, *debug, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = "%s.%d.tmp" % (LIB_PATH, pid)
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    for o in objs:
        with contextlib.suppress(OSError):
            os.remove(o)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

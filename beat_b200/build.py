"""In-tree build of libbeatgpu.so (nvcc, sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "beatgpu.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("beatgpu.cu", "sweep.cuh", "stack.cuh", "aux.cuh", "gemm.cuh", "tma.cuh", "probe.cuh", "geom.cuh", "geom_host.inc", "trace_io.inc")] + [
    os.path.join(ROOT, "include", "beatgpu.h")]
OUT = os.path.join(HERE, "libbeatgpu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-diag-suppress", "550", "-Xcompiler", "-pthread"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libbeatgpu.so")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def source_hash():
    """sha256 over the sources libbeatgpu.so is built from (first 16 hex digits): compiled into the library
    (beatgpu_source_hash) so that a committed ncu traffic record can be matched against the running build."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(DEPS):
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    return h.hexdigest()[:16]


def kernel_hash(files=("stack.cuh",)):
    """sha256 (16 hex digits) over the kernel sources a committed ncu traffic record belongs to (bench.py matches it,
    together with the L2 blocking the capture ran with, before reporting `roofline.traffic`)."""
    import hashlib
    h = hashlib.sha256()
    for f in files:
        h.update(open(os.path.join(HERE, "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def build(force=False, verbose=False, extra_flags=()):
    if not force and up_to_date():
        return OUT
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-DBEATGPU_SRC_HASH=\"%s\"" % source_hash(),
                                                            "-I", os.path.join(ROOT, "include"), "-o", OUT, SRC]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))

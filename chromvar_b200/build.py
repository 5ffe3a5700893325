"""In-tree build of libchromvar_b200.so (nvcc, sm_100a only).

    python -m chromvar_b200.build [--force] [--verbose]

The shared library is written to chromvar_b200/lib/libchromvar_b200.so; it is git-ignored
(history stays source-only) but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libchromvar_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["cv_counts.cu", "cv_deviations.cu", "cv_dense.cu", "cv_dense_fp4.cu", "cv_dense_build.cu", "cv_dense_gather.cu", "cv_background.cu", "cv_knn.cu", "cv_compat.cu", "cv_multi.cu", "cv_hostnarrow.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libchromvar_b200 cannot be built")


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link the shared library.
    Returns the library path.  Recompiles only objects whose inputs changed."""
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "chromvar_b200.h"))
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    relink = force or not os.path.exists(LIB)
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        stamp = obj + ".sha"
        dig = _digest([path] + headers, " ".join(NVCC_FLAGS))
        old = open(stamp).read() if os.path.exists(stamp) else ""
        if force or dig != old or not os.path.exists(obj):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(r.stdout + r.stderr)
            with open(stamp, "w") as f:
                f.write(dig)
            relink = True
        objs.append(obj)
    if relink:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Build rpkg/src/cv_rglue.c (the .Call glue) against the stand-in R runtime in rpkg/rstub/.

    python rpkg/build_stub.py

Produces rpkg/rstub/_build/chromVARb200.so = glue + rstub.c, linked against
chromvar_b200/lib/libchromvar_b200.so.  Test infrastructure: R itself is not installed in this
image; tests/test_gpu_rglue.py loads the result and drives every registered routine.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT_DIR = os.path.join(HERE, "rstub", "_build")
OUT = os.path.join(OUT_DIR, "chromVARb200.so")
SOURCES = [os.path.join(HERE, "src", "cv_rglue.c"), os.path.join(HERE, "rstub", "rstub.c")]
HEADERS = [os.path.join(HERE, "rstub", "include", n) for n in ("R.h", "Rinternals.h", os.path.join("R_ext", "Rdynload.h"))]
HEADERS.append(os.path.join(ROOT, "include", "chromvar_b200.h"))
LIBDIR = os.path.join(ROOT, "chromvar_b200", "lib")


def build(force=False):
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler: the .Call glue cannot be built")
    lib = os.path.join(LIBDIR, "libchromvar_b200.so")
    if not os.path.exists(lib):
        raise RuntimeError("libchromvar_b200.so is not built (python -m chromvar_b200.build)")
    deps = SOURCES + HEADERS + [lib]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [cc, "-shared", "-fPIC", "-O2", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Wno-cast-function-type",
           "-I", os.path.join(HERE, "rstub", "include"), "-I", os.path.join(ROOT, "include")] + SOURCES + \
          ["-L", LIBDIR, "-lchromvar_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../../../chromvar_b200/lib",
           "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the .Call glue failed")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

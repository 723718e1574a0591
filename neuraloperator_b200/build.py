"""In-tree build of libspectral_conv_b200.so (sm_100a only) with nvcc.

    python -m neuraloperator_b200.build [--force]

The shared object lands next to this file so that it travels with the source tree; it is never
installed into site-packages and there is no JIT cache.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libspectral_conv_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
SOURCES = ["sc_api.cu", "sc_generic.cu", "sc_fast.cu", "sc_collective.cu", "sc_layer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libspectral_conv_b200.so")


def _inputs():
    files = [os.path.join(CSRC, s) for s in SOURCES]
    files += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    files.append(os.path.join(ROOT, "include", "spectral_conv_b200.h"))
    return files


def needs_build():
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in _inputs())


def build_library(force=False, verbose=False):
    """Compiles every CUDA source of the package for sm_100a into one shared object. Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    nvcc = _nvcc()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", os.path.join(CSRC, src), "-o", obj]
        cmd += os.environ.get("SC_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DSC_TRACE_QUAD for the debug timeline build
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs]   # static cudart; driver entry points via cudaGetDriverEntryPoint
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libzebrapose_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is a plain
C-ABI shared object loaded with ctypes).  Every source is compiled to its own object (in parallel) and linked; the
exact-replay solver (zp_cvsolve.cu) is compiled with -fmad=false because its arithmetic must not be contracted."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzebrapose_b200.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["zp_api.cu", "zp_decode.cu", "zp_ransac.cu", "zp_cvsolve.cu", "zp_finsplit.cu", "zp_eval.cu", "zp_head.cu"]
EXTRA_FLAGS = {"zp_cvsolve.cu": ["-fmad=false"]}
HEADERS = ["zp_common.cuh", "zp_epnp.cuh", "zp_cvepnp.cuh", "zp_proj.cuh",
           os.path.join("..", "..", "include", "zebrapose_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def _newest_header():
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS + []) if HEADERS else 0


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    stamp = max(os.path.getmtime(path), _newest_header(), os.path.getmtime(os.path.abspath(__file__)))
    if os.path.exists(obj) and os.path.getmtime(obj) > stamp and not verbose:
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            if f.endswith(".o"):
                os.remove(os.path.join(OBJDIR, f))
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    cmd = [_nvcc(), "-shared", "-o", LIB] + [o for o, _ in res]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libzebrapose_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

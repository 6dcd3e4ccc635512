"""Build libcfnet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m coarse_fine_networks_b200.build [--force] [--verbose]
"""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libcfnet_b200.so")
OBJ_DIR = os.path.join(PKG, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include")]


def _digest(path, extra=()):
    h = hashlib.sha1()
    h.update(" ".join(extra).encode())
    for f in [path] + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "cfnet_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose, obj_dir=OBJ_DIR, extra=()):
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src, extra)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, True


VARIANTS = {"cvt": ("-DCFNET_AB", "-DCFNET_TF32_CVT"), "pdlwait": ("-DCFNET_AB", "-DCFNET_PDL_NOTRIGGER"),
            "p2timing": ("-DCFNET_AB", "-DCFNET_P2_TIMING")}      # per-role cycle counters of pw_tc2_kernel (tools/bench_pw.py)       # named experiment builds: libcfnet_b200_<name>.so


def build(force=False, verbose=False, ab=False, variant=None):
    """ab=True builds libcfnet_b200_ab.so with -DCFNET_AB: the same library with its experiment switches (environment
    variables) compiled in -- select it with CFNET_LIB=<path> for same-box A/B runs; the shipped library reads no environment."""
    obj_dir = OBJ_DIR + ("_ab" if ab else "")
    lib = LIB.replace(".so", "_ab.so") if ab else LIB
    extra = ("-DCFNET_AB",) if ab else ()
    if variant:
        obj_dir, lib, extra = OBJ_DIR + "_" + variant, LIB.replace(".so", "_" + variant + ".so"), VARIANTS[variant]
    os.makedirs(obj_dir, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if force:
        for f in glob.glob(os.path.join(obj_dir, "*.sha1")):
            os.remove(f)
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose, obj_dir, extra), srcs))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(lib):
        cmd = [NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ab="--ab" in sys.argv,
                variant=sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None))

"""Build libmpiflow_sm100.so (and the drop-in external/forward_warping/libwarping.so) in-tree.

    python mpi-flow_b200/build.py [--force]

Plain nvcc, sm_100a only, no torch/pybind headers: the product is a C-ABI shared library.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpiflow_sm100.so")
WARPLIB = os.path.join(HERE, "external", "forward_warping", "libwarping.so")
SOURCES = ["api.cu", "homography.cu", "render_fused.cu", "render_pk16.cu", "render_pk8.cu", "composite.cu", "splat.cu",
           "raster.cu", "assemble.cu", "inpaint.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static", "--threads", "0"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines, verbose=False):
    """development builds of the same library under another name (A/B runs: MPIFLOW_LIB=<path> tools/ab.py ...)"""
    nvcc = os.environ.get("NVCC", "nvcc")
    out = os.path.join(HERE, f"libmpiflow_sm100_{name}.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    extra = ["-Xptxas", "-v"] if verbose else []
    subprocess.check_call([nvcc] + NVCC_FLAGS + extra + [f"-D{d}" for d in defines] + ["-o", out] + srcs)
    return out


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mpiflow.h")]
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    extra = ["-Xptxas", "-v"] if verbose else []
    if force or _stale(LIB, deps):
        subprocess.check_call([nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + srcs)
    if force or _stale(WARPLIB, deps):
        os.makedirs(os.path.dirname(WARPLIB), exist_ok=True)
        # same symbol, same relative path as the reference loads (moving_obj.py:12-13): the same binary
        shutil.copyfile(LIB, WARPLIB)
        os.chmod(WARPLIB, 0o755)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:        # python build.py --variant scalar MPIFLOW_SCALAR_FP MPIFLOW_AB
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if not a.startswith("-")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

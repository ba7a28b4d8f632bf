"""In-tree build of the CUDA C-ABI library (sm_100a only) with plain nvcc.

    python -m sp_coupler_b200.build [--force] [--verbose]

Produces sp_coupler_b200/lib/libspcpl_b200.so (git-ignored; it travels to the GPU box with the
gpurun snapshot). nvcc cross-compiles without a GPU. No torch headers are needed: the boundary is
a C ABI (include/spcpl_b200.h) that the host side binds with ctypes.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libspcpl_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--compress-mode=size", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# translation unit -> extra flags. The profile kernels are compiled without FMA contraction so
# that float64 results are bit-identical to numpy's ("Numerics").
SOURCES = {
    "spc_abi.cu": [],
    "slab_reduce.cu": [],
    "profiles.cu": ["--fmad=false"],
    "les_state.cu": ["--fmad=false"],
    "nudge.cu": ["--fmad=false"],
}


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


TUNE_LIB = os.path.join(LIBDIR, "libspcpl_b200_tune.so")


def build(force=False, verbose=False, tune=False):
    """Production library (tune=False): only the kernel shapes the dispatch uses, no mutable global state.
    tune=True builds libspcpl_b200_tune.so with -DSPC_TUNING: the ring-shape sweep variants and the per-handle
    tuning setters (spc_tune_k1, spc_tune_profiles) that tools/*_probe.py drive; never loaded by the package."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(OBJDIR, "tune") if tune else OBJDIR
    os.makedirs(objdir, exist_ok=True)
    lib = TUNE_LIB if tune else LIB
    headers = [os.path.join(ROOT, "include", "spcpl_b200.h"), os.path.join(CSRC, "spc_common.cuh"), __file__]
    objs, rebuilt = [], False
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + ARCH + COMMON + extra + (["-DSPC_TUNING"] if tune else []) + \
                  (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            rebuilt = True
    if rebuilt or not os.path.exists(lib):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", lib] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return lib


TORCH_LIB = os.path.join(LIBDIR, "spcpl_b200_torch.so")


def build_torch_ext(force=False, verbose=False):
    """The thin PyTorch C++ extension (csrc/torch_ext.cpp -> torch.ops.spcpl_b200.*): plain g++ against
    torch's headers, linked to libspcpl_b200.so next to it (rpath $ORIGIN). In-tree, no JIT cache."""
    build(force=force, verbose=verbose)
    src = os.path.join(CSRC, "torch_ext.cpp")
    if not (force or _stale(TORCH_LIB, [src, LIB, os.path.join(ROOT, "include", "spcpl_b200.h")])):
        return TORCH_LIB
    import torch
    import torch.utils.cpp_extension as ce
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    for inc in ce.include_paths("cuda"):
        cmd += ["-isystem", inc]
    cmd += ["-I", os.path.join(ROOT, "include"), src, "-o", TORCH_LIB]
    for lib in ce.library_paths("cuda"):
        cmd += ["-L" + lib, "-Wl,-rpath," + lib]
    cmd += ["-L" + LIBDIR, "-Wl,-rpath,$ORIGIN", "-lspcpl_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
            "-ltorch"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    if "--tune" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, tune=True))
    if "--torch" in sys.argv:
        print(build_torch_ext(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

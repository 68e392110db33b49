"""Build libwbc_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so travels with the repo."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libwbc_b200.so")
SOURCES = ["wbc_b200.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "wbc_b200.h")]
# -maxrregcount: the kernels carry __launch_bounds__, which overrides it for the entry functions (ptxas still reports 168 / 255
# registers for them); what it caps is the register budget of the separately compiled device functions the solver calls (symv,
# chol_build30, tri_solve30, the multiplier update ...).  Measured on B200, interleaved A/B of variant builds on one box
# (profiles/r02_al_maxrregcount_ab.txt): any cap from 112 to 160 takes the 4 096-instance solver kernel from 3.18 to 3.01 ms
# (+5.5 % solves/s; 96 a little less), the 65 536-instance stage-task kernel is unchanged within +-0.5 %; results bit-identical.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-maxrregcount=128",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-Wno-deprecated-gpu-targets"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libwbc_b200.so cannot be built (there is no CPU fallback)")


def have_nvcc():
    return any(c and os.path.exists(c) for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources.  Returns the .so path."""
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("WBC_NVCC_EXTRA", "").split()      # experiments only (e.g. -DWBC_SOLVE_T=64)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libwbc_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

"""In-tree build of the CUDA library (nvcc, sm_100a only).  `python -m eepacc_mpc_casadi_matlab_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeepacc.so")
SOURCES = ["k1_eval.cu", "k2_build.cu", "k3_qp.cu", "k3r_qp.cu", "k4_post.cu", "k5_scenario.cu", "k6_ocp.cu", "k7_nlp.cu", "eepacc_capi.cu"]
HEADERS = ["eepacc_dev.cuh", "eepacc_kernels.h", "eepacc_rows.cuh", "k2_build.cuh", "k4_post.cuh", "qp_rules.cuh", "k6_ocp.cuh",
           os.path.join("..", "..", "include", "eepacc.h")]
# headers only some sources include (a change there must not recompile the 12-minute QP kernels)
EXTRA_HEADERS = {"k7_nlp.cu": ["k7_nlp.cuh", os.path.join("..", "..", "include", "eepacc_nlp.h")]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, timers=None):
    """Compile the sources concurrently (one nvcc per file) and link.  timers=True (or EEP_K3_TIMERS=1)
    builds the developer variant libeepacc_timers.so with clock64() phase timers in the QP kernels;
    api.load_library() picks it up through EEPACC_LIB."""
    nvcc = os.environ.get("NVCC", "nvcc")
    if timers is None:
        timers = bool(os.environ.get("EEP_K3_TIMERS"))
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objdir = os.path.join(HERE, "build_timers" if timers else "build")
    lib = LIB.replace("libeepacc.so", "libeepacc_timers.so") if timers else LIB
    os.makedirs(objdir, exist_ok=True)
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        extra = [os.path.join(CSRC, h) for h in EXTRA_HEADERS.get(s, [])]
        if force or _stale(obj, [src] + hdrs + extra):
            cmd = [nvcc] + NVCC_FLAGS + (["-DEEP_K3_TIMERS"] if timers else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            jobs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, job in jobs:
        if job.wait() != 0:
            raise subprocess.CalledProcessError(job.returncode, cmd)
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return lib


def build_variant(name, k3r_flags, verbose=True):
    """Developer aid: libeepacc_<name>.so = the standard objects with k3r_qp.cu recompiled under extra
    flags (e.g. -DK3R_NT=64 -DK3R_MINB=4).  Select it at run time with EEPACC_LIB=<path>."""
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "build")            # the standard objects must exist (build_library)
    vobj = os.path.join(objdir, f"k3r_qp_{name}.o")
    cmd = [nvcc] + NVCC_FLAGS + list(k3r_flags) + ["-Xptxas", "-v", "-c", os.path.join(CSRC, "k3r_qp.cu"), "-o", vobj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES if s != "k3r_qp.cu"] + [vobj]
    lib = LIB.replace("libeepacc.so", f"libeepacc_{name}.so")
    subprocess.check_call([nvcc, "-shared", "-o", lib] + objs + ["-lcudart"])
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build_library(force="--force" in sys.argv, verbose=True, timers=("--timers" in sys.argv) or None))

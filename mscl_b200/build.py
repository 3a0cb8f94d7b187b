"""Build libmscl_b200.so (sm_100a only) in-tree with nvcc.

    python -m mscl_b200.build          # rebuild if sources are newer than the library

The library is the C-ABI drop-in boundary declared in include/mscl_b200.h; it links
the CUDA runtime statically and resolves the one driver symbol it needs
(cuTensorMapEncodeTiled) at run time, so it loads on a machine without a GPU driver
(the symbol-export test runs there).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.environ.get("MSCL_LIB_OUT") or os.path.join(LIBDIR, "libmscl_b200.so")      # MSCL_LIB_OUT: debug builds kept aside
SOURCES = ["abi.cu", "enqueue.cu", "ema.cu", "fra.cu", "lmcl.cu", "infonce.cu", "infonce_tc.cu", "infonce_fused.cu", "resample.cu", "augment.cu", "optim.cu", "retrieval.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "mscl_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o") if not os.environ.get("MSCL_LIB_OUT") else "dbg_" + src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if os.environ.get("MSCL_TIMELINE"):      # debug build: per-CTA phase timestamps in the tcgen05 kernel
            cmd.insert(1, "-DMSCL_TC_TIMELINE")
        for d in os.environ.get("MSCL_DEFS", "").split():      # experiment switches, e.g. MSCL_DEFS="-DMSCL_EXP_NOLOAD2"
            cmd.insert(1, d)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build libvdetr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python v-detr_b200/build.py [--force] [-v]

The library is a plain CUDA-runtime C-ABI shared object (no torch, no pybind): see include/vdetr_b200.h.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# developer knob: VDETR_LIB_SUFFIX=_x builds lib_x/libvdetr_b200.so (an experimental variant next to the product library;
# load it with VDETR_B200_LIB=<path>, see _C.py)
LIBDIR = os.path.join(HERE, "lib" + os.environ.get("VDETR_LIB_SUFFIX", ""))
LIB = os.path.join(LIBDIR, "libvdetr_b200.so")
SOURCES = ["api.cu", "pointnet2.cu", "rpe_simt.cu", "rpe_xattn_fwd.cu", "rpe_xattn_bwd.cu", "rpe_dtables.cu", "rpe_dtables_umma.cu", "layernorm.cu", "batchnorm.cu", "boxdecode.cu", "matcher.cu", "optim.cu", "peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# developer knob: VDETR_EXTRA_NVCC="-DVDETR_DT_THREADS=768 -DVDETR_DT_QB=15" (tuning experiments)
NVCC_FLAGS += os.environ.get("VDETR_EXTRA_NVCC", "").split()


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and no up-to-date prebuilt libvdetr_b200.so in " + LIBDIR)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = [], None
    for src, p in procs:
        out = p.communicate()[0]
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0 and failed is None:
            failed = src
    open(os.path.join(LIBDIR, "ptxas.log"), "w").write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError(f"nvcc failed on {failed}")
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs)
    open(stamp, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

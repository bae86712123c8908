"""Build libivv_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot. No torch types cross the ABI;
the library links only against cudart (the driver's cuTensorMapEncodeTiled is resolved at run time).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libivv_b200.so")
SOURCES = ["api.cu", "gemm_tc.cu", "attention_tc.cu", "norm.cu", "temporal_attn.cu", "elementwise.cu", "warp.cu",
           "sampler.cu", "raft.cu"]
EXTRA_FLAGS = os.environ.get("IVV_NVCC_EXTRA", "").split()  # tuning builds only (e.g. -DIVV_TUNING)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stamp():
    """Hash of exactly what the library is built from: SOURCES, the headers beside them and the C-ABI header."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, n) for n in SOURCES] + \
        sorted(os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.endswith((".cuh", ".h")))
    files.append(os.path.join(HERE, "..", "include", "ivv.h"))
    for path in files:
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + EXTRA_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu into objects (parallel) and link the shared library. Returns the library path."""
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    stamp_file = os.path.join(objdir, "build_stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + EXTRA_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[ivv build] {src} FAILED\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[ivv build] {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed building libivv_b200.so")
    tmp = LIB + f".tmp{os.getpid()}"  # link beside the target, then rename: a process that mapped the old file keeps it
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

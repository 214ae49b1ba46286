"""Build ``libdcpt_sm100.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dcpt_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdcpt_sm100.so")
SOURCES = ["gemm_sm100.cu", "gemm_simt.cu", "layernorm.cu", "dwconv.cu", "misc.cu", "conv3x3_img.cu", "dchead.cu", "restormer.cu", "nafnet.cu", "optim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--threads", "4"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dcpt_ops.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True, trace=False, fp16=False):
    """trace=True builds the debug variant libdcpt_sm100_trace.so (-DDCPT_TRACE: per-CTA GEMM event timeline, tools/gemm_trace.py);
    fp16=True the parity variant libdcpt_sm100_fp16.so (-DDCPT_OPERAND_FP16: IEEE-half tensor-core operands, DCPT_OPERAND=fp16)."""
    out = LIB.replace(".so", "_trace.so") if trace else (LIB.replace(".so", "_fp16.so") if fp16 else LIB)
    if not trace and not force and not needs_build(out):
        return out
    defs = (["-DDCPT_TRACE"] if trace else []) + (["-DDCPT_OPERAND_FP16"] if fp16 else [])
    cmd = [_nvcc()] + NVCC_FLAGS + defs + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    if verbose:
        print("[dcpt_b200] " + " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, trace="--trace" in sys.argv, fp16="--fp16" in sys.argv))

"""Build the C-ABI shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m chunkformer_b200.build            # -> chunkformer_b200/csrc/libchunkformer_b200.so
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libchunkformer_b200.so")
LIB_ABLATION = os.path.join(CSRC, "libchunkformer_b200_ablation.so")   # tools only: -DCF_ABLATION (timing / phase switches)
SOURCES = ["api.cu", "plan.cpp"]
HEADERS = ["common.cuh", "transducer.cuh", "gemm.cuh", "gemm_host.cuh", "gemm_ln.cuh", "ffn_fused.cuh", "norm_conv.cuh", "frontend.cuh", "attention_simt.cuh",
           "attention_tc.cuh", "fbank.cuh", "misc_kernels.cuh", "plan.h", os.path.join("..", "..", "include", "chunkformer_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, ablation: bool = False) -> str:
    """Product library by default.  ablation=True builds the tools flavour next to it: the same sources with -DCF_ABLATION,
    which compiles in the phase-ablation switches (CF_GEMM_DEBUG, CF_RNNT_DEBUG environment variables) that the product
    library does not contain; select it with CHUNKFORMER_B200_LIB=<path> when running tools/."""
    out = LIB_ABLATION if ablation else LIB
    if not ablation and not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "--shared", "-cudart", "static",
           "-o", out] + (["-DCF_ABLATION"] if ablation else []) + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ablation="--ablation" in sys.argv))

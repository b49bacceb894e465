"""Build the REFERENCE's own CUDA extensions for sm_100a into baseline/_ref/ (git-ignored), from the sources where they lie
under /root/reference - nothing is copied.  They are the "kernel to beat" timed beside ours by scripts/ref_cuda_bench.py on
the GPU box; they are never linked into or imported by the product.

Arch flags follow requirements/Mamba/mamba/setup.py:108-114,137-160 with compute_100a added (the shipped list stops at sm_90).
Only the fp32 / bf16 real-A kernels are compiled (baseline/ref_stubs.cpp stubs the other instantiations).
nvcc cross-compiles without a GPU: run this in the build container, the .so files travel with the gpurun snapshot.
"""
import os
import sys

os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("MAX_JOBS", "4")
from torch.utils.cpp_extension import load  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MMU_REFERENCE_ROOT", "/root/reference")
SS = os.path.join(REF, "requirements/Mamba/mamba/csrc/selective_scan")
CC = os.path.join(REF, "requirements/Mamba/causal-conv1d/csrc")
NVCC = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_BFLOAT16_OPERATORS__",
        "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
        "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a"]


def build(which=("selective_scan_cuda", "causal_conv1d_cuda")):
    if not os.path.isdir(SS):
        raise SystemExit(f"reference sources not found under {REF}")
    out = {}
    if "selective_scan_cuda" in which:
        d = os.path.join(HERE, "_ref", "selective_scan_cuda")
        os.makedirs(d, exist_ok=True)
        out["selective_scan_cuda"] = load(
            name="selective_scan_cuda", build_directory=d, verbose=True, is_python_module=True,
            sources=[os.path.join(SS, f) for f in ("selective_scan.cpp", "selective_scan_fwd_fp32.cu", "selective_scan_fwd_bf16.cu",
                                                   "selective_scan_bwd_fp32_real.cu", "selective_scan_bwd_bf16_real.cu")]
            + [os.path.join(HERE, "ref_stubs.cpp")],
            extra_include_paths=[SS], extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=NVCC)
    if "causal_conv1d_cuda" in which:
        d = os.path.join(HERE, "_ref", "causal_conv1d_cuda")
        os.makedirs(d, exist_ok=True)
        out["causal_conv1d_cuda"] = load(
            name="causal_conv1d_cuda", build_directory=d, verbose=True, is_python_module=True,
            sources=[os.path.join(CC, f) for f in ("causal_conv1d.cpp", "causal_conv1d_fwd.cu", "causal_conv1d_bwd.cu", "causal_conv1d_update.cu")],
            extra_include_paths=[CC], extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=NVCC)
    return out


if __name__ == "__main__":
    build(tuple(sys.argv[1:]) or ("selective_scan_cuda", "causal_conv1d_cuda"))
    print("built reference CUDA extensions under baseline/_ref/")

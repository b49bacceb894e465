// Build helper for baseline/build_ref_cuda.py (NOT product code, never linked into libmmunet_b200.so).
// The reference's selective_scan.cpp dispatches over {fp32, fp16, bf16} x {real, complex}; only the fp32 / bf16 real kernels are
// timed on B200, so the instantiations of the translation units that are not compiled are provided here as stubs that raise.
#include <ATen/ATen.h>
#include <c10/util/complex.h>
#include <cuda_runtime.h>

#include "selective_scan.h"   // the reference's own header, found through the include path (requirements/Mamba/mamba/csrc/selective_scan)

using complex_t = c10::complex<float>;
template <typename input_t, typename weight_t> void selective_scan_fwd_cuda(SSMParamsBase &params, cudaStream_t stream);
template <typename input_t, typename weight_t> void selective_scan_bwd_cuda(SSMParamsBwd &params, cudaStream_t stream);

#define REF_STUB(FN, PARAMS, IT, WT) \
    template <> void FN<IT, WT>(PARAMS &, cudaStream_t) { TORCH_CHECK(false, #FN "<" #IT ", " #WT "> was not built for the B200 timing arm"); }
REF_STUB(selective_scan_fwd_cuda, SSMParamsBase, at::Half, float)
REF_STUB(selective_scan_fwd_cuda, SSMParamsBase, at::Half, complex_t)
REF_STUB(selective_scan_bwd_cuda, SSMParamsBwd, at::Half, float)
REF_STUB(selective_scan_bwd_cuda, SSMParamsBwd, at::Half, complex_t)
REF_STUB(selective_scan_bwd_cuda, SSMParamsBwd, float, complex_t)
REF_STUB(selective_scan_bwd_cuda, SSMParamsBwd, at::BFloat16, complex_t)

// Host side of the tensor-map (TMA) B/C tile variant of the v3 scan kernels (MMU_TMA_TILE=1, scan3_fwd.cuh / scan3_bwd.cuh).
#pragma once
#include "scan3.cuh"
#if MMU_TMA_TILE
#include <cuda.h>
#include <cudaTypedefs.h>
namespace mmu {
// tensor maps of B / C for the TMA tile experiment: (L, dstate, batch) fp32, box {256 tokens, 16 states, 1}
inline int encode_bc_map(void *out, const void *base, int L, int N, int B, int64_t ns, int64_t bs) {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
    }();
    if (fn == nullptr) return set_error(MMU_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[3] = {(cuuint64_t)L, (cuuint64_t)N, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)ns * 4, (cuuint64_t)bs * 4};
    const cuuint32_t box[3] = {256, 16, 1}, es[3] = {1, 1, 1};
    const CUresult rc = fn(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? MMU_OK : set_error(MMU_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed: %d", (int)rc);
}
}  // namespace mmu
#endif

// Scan-order gather / scatter (bit-exact index maps) for sm_100a.
//
// The reference expresses these as torch view/permute/flip/stack copies:
//   two-row column-interleaved "morph" order   src/UM_Net/MMUNet.py:68-93 (flatten), :95-121 (inverse)
//   flip                                       requirements/mamba_simple.py:230
//   nslices interleave                         requirements/mamba_simple.py:245-247 (gather), :263 (scatter)
// Here idx(l) is evaluated in closed form per element; the kernel is a pure HBM-bound permutation:
//   gather : dst[r][l]      = src[r][idx(l)]
//   scatter: dst[r][idx(l)] = src[r][l]
#include <algorithm>

#include "common.cuh"

namespace mmu {

using OrderMap = OrdMap;   // common.cuh

template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256) scan_order_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t rows,
                                                         int64_t src_rs, int64_t dst_rs, OrderMap m) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= m.L) return;
    const int k = m(l);
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        if (SCATTER) dst[r * dst_rs + k] = src[r * src_rs + l];
        else dst[r * dst_rs + l] = src[r * src_rs + k];
    }
}

// NSLICES through shared memory: per row the map is the transpose of an (ns x Ls) matrix (gather: dst[j*ns+s] = src[s*Ls+j];
// scatter: the inverse).  A CTA moves TJ = 64 consecutive j of every slice: global reads and writes are both contiguous
// runs (TJ elements per slice on the sliced side, TJ*ns elements on the interleaved side) instead of one element per sector.
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256) nslices_tiled_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t rows, int64_t src_rs,
                                                            int64_t dst_rs, int ns, int Ls) {
    constexpr int TJ = 64;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);                  // [ns][TJ + 1]
    const int j0 = blockIdx.x * TJ, nj = min(TJ, Ls - j0), tid = threadIdx.x;
    // element e of the interleaved run <-> (j, slice): e = j*ns + sl.  With 256 % ns == 0 a thread keeps its slice and advances
    // j by 256/ns per step (no division in the loops); otherwise one 32-bit division per element.
    const bool pow2 = (256 % ns) == 0;
    const int sl_f = tid % ns, j_f = tid / ns, dj = pow2 ? 256 / ns : 0;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const T *s = src + r * src_rs;
        T *d = dst + r * dst_rs;
        if (!SCATTER) {
            for (int e = tid; e < ns * TJ; e += 256) {          // sliced side: TJ-element runs
                const int sl = e / TJ, j = e - sl * TJ;
                if (j < nj) tile[sl * (TJ + 1) + j] = s[(int64_t)sl * Ls + j0 + j];
            }
            __syncthreads();
            // (the division-free form of this loop measured 60 % SLOWER here - 747 vs 467 us on 537 MB - so the gather keeps the division)
            for (int e = tid; e < nj * ns; e += 256) {          // interleaved side: one contiguous run of nj*ns elements
                const int j = e / ns, sl = e - j * ns;
                d[(int64_t)j0 * ns + e] = tile[sl * (TJ + 1) + j];
            }
        } else {
            if (pow2) {
                for (int e = tid, j = j_f; j < nj; e += 256, j += dj) tile[sl_f * (TJ + 1) + j] = s[(int64_t)j0 * ns + e];
            } else {
                for (int e = tid; e < nj * ns; e += 256) {
                    const int j = e / ns, sl = e - j * ns;
                    tile[sl * (TJ + 1) + j] = s[(int64_t)j0 * ns + e];
                }
            }
            __syncthreads();
            for (int e = tid; e < ns * TJ; e += 256) {
                const int sl = e / TJ, j = e - sl * TJ;
                if (j < nj) d[(int64_t)sl * Ls + j0 + j] = tile[sl * (TJ + 1) + j];
            }
        }
        __syncthreads();
    }
}

__global__ void scan_order_index_kernel(int64_t *idx, OrderMap m) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < m.L) idx[l] = m(l);
}

namespace {
int make_map(OrderMap &m, int order, int H, int W, int ns) {
    if (H <= 0 || W <= 0) return set_error(MMU_ERR_INVALID, "scan_order: empty map %dx%d", H, W);
    if ((int64_t)H * W > INT32_MAX) return set_error(MMU_ERR_UNSUPPORTED, "scan_order: H*W exceeds int32");
    if (order < MMU_ORDER_ROWMAJOR || order > MMU_ORDER_TWOROW) return set_error(MMU_ERR_INVALID, "scan_order: order %d", order);
    m.kind = order, m.W = W, m.L = H * W, m.ns = 1, m.Ls = m.L, m.even_tokens = 2 * (H / 2) * W, m.ns_shift = -1;
    if (order == MMU_ORDER_NSLICES) {
        if (ns <= 0 || m.L % ns != 0) return set_error(MMU_ERR_INVALID, "scan_order: L=%d not divisible by nslices=%d", m.L, ns);
        m.ns = ns, m.Ls = m.L / ns;
    }
    return MMU_OK;
}

template <bool SCATTER>
int run_order(const void *src, void *dst, int dtype, int64_t rows, int64_t src_rs, int64_t dst_rs, int order, int H, int W,
              int ns, void *stream) {
    OrderMap m;
    if (int rc = make_map(m, order, H, W, ns)) return rc;
    if (rows <= 0) return set_error(MMU_ERR_INVALID, "scan_order: rows=%lld", (long long)rows);
    if (!src || !dst) return set_error(MMU_ERR_INVALID, "scan_order: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (order == MMU_ORDER_NSLICES && m.ns >= 2 && m.ns <= 256 && m.Ls >= 16 && (dtype == MMU_F32 || dtype == MMU_BF16 || dtype == MMU_F16)) {
        dim3 tg((m.Ls + 63) / 64, (unsigned)std::min<int64_t>(rows, 65535));
        const size_t es = dtype == MMU_F32 ? 4 : 2, smem = (size_t)m.ns * 65 * es;
        if (dtype == MMU_F32) {
            auto k = nslices_tiled_kernel<float, SCATTER>;
            if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<tg, 256, smem, st>>>((const float *)src, (float *)dst, rows, src_rs, dst_rs, m.ns, m.Ls);
        } else {
            auto k = nslices_tiled_kernel<uint16_t, SCATTER>;
            k<<<tg, 256, smem, st>>>((const uint16_t *)src, (uint16_t *)dst, rows, src_rs, dst_rs, m.ns, m.Ls);
        }
        count_launch();
        return check_launch("scan_order (nslices, tiled)");
    }
    dim3 grid((m.L + 255) / 256, (unsigned)std::min<int64_t>(rows, 65535));
    if (dtype == MMU_F32)
        scan_order_kernel<float, SCATTER><<<grid, 256, 0, st>>>((const float *)src, (float *)dst, rows, src_rs, dst_rs, m);
    else if (dtype == MMU_BF16 || dtype == MMU_F16)
        scan_order_kernel<uint16_t, SCATTER><<<grid, 256, 0, st>>>((const uint16_t *)src, (uint16_t *)dst, rows, src_rs, dst_rs, m);
    else
        return set_error(MMU_ERR_UNSUPPORTED, "scan_order: dtype %d", dtype);
    count_launch();
    return check_launch("scan_order");
}
}  // namespace
}  // namespace mmu

extern "C" int mmu_scan_order_gather(const void *src, void *dst, int32_t dtype, int64_t rows, int64_t src_rs, int64_t dst_rs,
                                     int32_t order, int32_t H, int32_t W, int32_t nslices, void *stream) {
    return mmu::run_order<false>(src, dst, dtype, rows, src_rs, dst_rs, order, H, W, nslices, stream);
}
extern "C" int mmu_scan_order_scatter(const void *src, void *dst, int32_t dtype, int64_t rows, int64_t src_rs, int64_t dst_rs,
                                      int32_t order, int32_t H, int32_t W, int32_t nslices, void *stream) {
    return mmu::run_order<true>(src, dst, dtype, rows, src_rs, dst_rs, order, H, W, nslices, stream);
}
extern "C" int mmu_scan_order_index(int64_t *idx_dev, int32_t order, int32_t H, int32_t W, int32_t nslices, void *stream) {
    using namespace mmu;
    OrderMap m;
    if (int rc = make_map(m, order, H, W, nslices)) return rc;
    if (!idx_dev) return set_error(MMU_ERR_INVALID, "scan_order_index: null pointer");
    scan_order_index_kernel<<<(m.L + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(idx_dev, m);
    count_launch();
    return check_launch("scan_order_index");
}

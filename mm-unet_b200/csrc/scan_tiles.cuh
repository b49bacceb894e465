// Cooperative global <-> shared tile movers for the selective-scan kernels.
//
// A tile is ROWS x TL tokens of one (batch) slice of a (batch, rows, seqlen) tensor whose sequence stride is 1.
// In shared memory every tile is fp32, row pitch TL, 16-byte chunks XOR-swizzled (common.cuh: swz_chunk).
// `reverse` maps logical token t to memory position L-1-t (fused flip, requirements/mamba_simple.py:230).
// Tokens >= L and rows >= nrows_valid load as `fill` and are never stored.
#pragma once
#include "common.cuh"

namespace mmu {

template <typename IN_T> struct alignas(4 * sizeof(IN_T)) Quad { IN_T v[4]; };

template <typename IN_T, int T, int TL, int NT, typename F>
__device__ __forceinline__ void load_tile(float *__restrict__ dst, const IN_T *__restrict__ base, int64_t row_stride,
                                          int row0, int rows_tile, int nrows_valid, int t0, int L, bool reverse,
                                          bool vec_ok, int tid, float fill, F f) {
    constexpr int CPR = TL / 4;   // chunks per row
    for (int idx = tid; idx < rows_tile * CPR; idx += NT) {
        const int r = idx / CPR, c = idx - r * CPR;
        const int row = row0 + r;
        const int t = t0 + 4 * c;   // first logical token of the chunk
        float v[4] = {fill, fill, fill, fill};
        if (row < nrows_valid && t < L) {
            const IN_T *rp = base + (int64_t)row * row_stride;
            if (vec_ok && t + 3 < L) {
                if (!reverse) {
                    const Quad<IN_T> q = *reinterpret_cast<const Quad<IN_T> *>(rp + t);
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = Elem<IN_T>::to_f(q.v[k]);
                } else {
                    const Quad<IN_T> q = *reinterpret_cast<const Quad<IN_T> *>(rp + (L - 4 - t));
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = Elem<IN_T>::to_f(q.v[3 - k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (t + k < L) v[k] = Elem<IN_T>::to_f(rp[reverse ? (L - 1 - t - k) : (t + k)]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (t + k < L) v[k] = f(r, v[k]);
                else v[k] = fill;
        }
        *reinterpret_cast<float4 *>(dst + r * TL + 4 * swz_chunk<T>(c)) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

template <typename OUT_T, int T, int TL, int NT>
__device__ __forceinline__ void store_tile(const float *__restrict__ src, OUT_T *__restrict__ base, int64_t row_stride,
                                           int row0, int rows_tile, int nrows_valid, int t0, int L, bool reverse,
                                           bool vec_ok, int tid) {
    constexpr int CPR = TL / 4;
    for (int idx = tid; idx < rows_tile * CPR; idx += NT) {
        const int r = idx / CPR, c = idx - r * CPR;
        const int row = row0 + r;
        const int t = t0 + 4 * c;
        if (row >= nrows_valid || t >= L) continue;
        const float4 q = *reinterpret_cast<const float4 *>(src + r * TL + 4 * swz_chunk<T>(c));
        const float v[4] = {q.x, q.y, q.z, q.w};
        OUT_T *rp = base + (int64_t)row * row_stride;
        if (vec_ok && t + 3 < L) {
            Quad<OUT_T> o;
#pragma unroll
            for (int k = 0; k < 4; ++k) o.v[reverse ? 3 - k : k] = Elem<OUT_T>::from_f(v[k]);
            *reinterpret_cast<Quad<OUT_T> *>(rp + (reverse ? (L - 4 - t) : t)) = o;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (t + k < L) rp[reverse ? (L - 1 - t - k) : (t + k)] = Elem<OUT_T>::from_f(v[k]);
        }
    }
}

// host-side: can a (base, batch stride, row stride) tensor use 4-element vector access?
template <typename IN_T> inline bool quad_ok(const void *p, int64_t bs, int64_t ds, int L, bool reverse) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return (a % (4 * sizeof(IN_T)) == 0) && (bs % 4 == 0) && (ds % 4 == 0) && (!reverse || L % 4 == 0);
}

}  // namespace mmu

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

// ---- fast path: register prefetch (issue the loads for chunk c+1 before computing chunk c) ------------------------------
// Preconditions (checked by the caller): 4-element vector access is legal for the tensor and the chunk is entirely inside
// the sequence.  `total` = rows_tile * TL/4 quads are spread over the NT threads, K = ceil(total / NT) per thread.
template <typename IN_T, int TL, int NT, int K>
__device__ __forceinline__ void tile_prefetch(Quad<IN_T> (&q)[K], const IN_T *__restrict__ base, int64_t row_stride, int row0,
                                              int nrows_valid, int total, int t0, int L, bool reverse, int tid) {
    constexpr int CPR = TL / 4;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int idx = tid + k * NT;
        const int r = idx / CPR, c = idx - r * CPR;
        const int row = row0 + r;
        Quad<IN_T> v;
#pragma unroll
        for (int e = 0; e < 4; ++e) v.v[e] = Elem<IN_T>::from_f(0.f);
        if (idx < total && row < nrows_valid) {
            const int t = t0 + 4 * c;
            v = *reinterpret_cast<const Quad<IN_T> *>(base + (int64_t)row * row_stride + (reverse ? L - 4 - t : t));
        }
        q[k] = v;
    }
}

template <typename IN_T, int T, int TL, int NT, int K, typename F>
__device__ __forceinline__ void tile_commit(float *__restrict__ dst, const Quad<IN_T> (&q)[K], int total, bool reverse,
                                            int tid, F f) {
    constexpr int CPR = TL / 4;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int idx = tid + k * NT;
        if (idx < total) {
            const int r = idx / CPR, c = idx - r * CPR;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = f(r, Elem<IN_T>::to_f(q[k].v[reverse ? 3 - e : e]));
            *reinterpret_cast<float4 *>(dst + r * TL + 4 * swz_chunk<T>(c)) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ---- B / C tiles: pair-interleaved layout ------------------------------------------------------------------------------
// Shared layout [NP][TL][2] fp32: the two dstate rows of a pair sit next to each other per token, so one LDS.128 yields
// two ready-made (B_2p, B_2p+1) float2 operands for FFMA2 (no register shuffling).  A 16-byte chunk holds 2 tokens x 2
// states; a lane reads T/2 = 4 consecutive chunks, so the 4-chunks-per-lane swizzle (swz_chunk<16>) applies.
__device__ __forceinline__ int bc_off(int chunk2) { return 4 * swz_chunk<16>(chunk2); }

template <typename IN_T, int TL, int NT, int K2>
__device__ __forceinline__ void bc_prefetch(Quad<IN_T> (&q)[2 * K2], const IN_T *__restrict__ base, int64_t n_stride, int N,
                                            int NP, int t0, int L, bool reverse, int tid) {
    constexpr int CPR = TL / 4;
#pragma unroll
    for (int k = 0; k < K2; ++k) {
        const int idx = tid + k * NT;
        const int pr = idx / CPR, c = idx - pr * CPR;
        const int t = t0 + 4 * c;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            Quad<IN_T> v;
#pragma unroll
            for (int e = 0; e < 4; ++e) v.v[e] = Elem<IN_T>::from_f(0.f);
            const int n = 2 * pr + h;
            if (idx < NP * CPR && n < N)
                v = *reinterpret_cast<const Quad<IN_T> *>(base + (int64_t)n * n_stride + (reverse ? L - 4 - t : t));
            q[2 * k + h] = v;
        }
    }
}

template <typename IN_T, int TL, int NT, int K2>
__device__ __forceinline__ void bc_commit(float *__restrict__ dst, const Quad<IN_T> (&q)[2 * K2], int NP, bool reverse, int tid) {
    constexpr int CPR = TL / 4;
#pragma unroll
    for (int k = 0; k < K2; ++k) {
        const int idx = tid + k * NT;
        if (idx < NP * CPR) {
            const int pr = idx / CPR, c = idx - pr * CPR;
            float a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[e] = Elem<IN_T>::to_f(q[2 * k].v[reverse ? 3 - e : e]);
                b[e] = Elem<IN_T>::to_f(q[2 * k + 1].v[reverse ? 3 - e : e]);
            }
            float *row = dst + pr * 2 * TL;
            *reinterpret_cast<float4 *>(row + bc_off(2 * c)) = make_float4(a[0], b[0], a[1], b[1]);
            *reinterpret_cast<float4 *>(row + bc_off(2 * c + 1)) = make_float4(a[2], b[2], a[3], b[3]);
        }
    }
}

// generic (unaligned / ragged chunk) loader into the same layout
template <typename IN_T, int TL, int NT>
__device__ __forceinline__ void bc_load_generic(float *__restrict__ dst, const IN_T *__restrict__ base, int64_t n_stride, int N,
                                                int Ne, int t0, int L, bool reverse, int tid) {
    for (int i = tid; i < Ne * TL; i += NT) {
        const int n = i / TL, tok = i - n * TL, t = t0 + tok;
        const float v = (n < N && t < L) ? Elem<IN_T>::to_f(base[(int64_t)n * n_stride + (reverse ? L - 1 - t : t)]) : 0.f;
        dst[(n >> 1) * 2 * TL + bc_off(tok >> 1) + 2 * (tok & 1) + (n & 1)] = v;
    }
}

// store 4 consecutive logical tokens t..t+3 of one row (vector when legal, scalar with bounds otherwise)
template <typename OUT_T>
__device__ __forceinline__ void store_quad(OUT_T *__restrict__ rp, int t, int L, bool reverse, bool vec_ok, const float v[4]) {
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

// host-side: can a (base, batch stride, row stride) tensor use 4-element vector access?
template <typename IN_T> inline bool quad_ok(const void *p, int64_t bs, int64_t ds, int L, bool reverse) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return (a % (4 * sizeof(IN_T)) == 0) && (bs % 4 == 0) && (ds % 4 == 0) && (!reverse || L % 4 == 0);
}

}  // namespace mmu

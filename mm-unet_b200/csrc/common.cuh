// Shared device helpers for the sm_100a Mamba-block kernels (selective scan, causal conv1d, scan orders).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmunet_b200.h"

namespace mmu {

constexpr float kLog2e = 1.4426950408889634f;

// ---- host-side error plumbing (capi.cu) -------------------------------------------------------
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);
int knob(const char *name, int dflt);   // MMU_* tuning knob (environment, read once: capi.cu)

// ---- element conversion -----------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kVec = 4;   // elements per 16 bytes
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
    static constexpr int kVec = 8;
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

// 16-byte vector of elements
template <typename T> struct alignas(16) Vec16 { T v[Elem<T>::kVec]; };

// ---- packed fp32 (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ----------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }

// MUFU.EX2
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 ex2(float2 x) { return make_float2(ex2(x.x), ex2(x.y)); }

// softplus with the reference's threshold (selective_scan_fwd_kernel.cuh:153-156; F.softplus default)
// log1p(e^x) without the libm log1pf call: for small e = e^x the alternating series (error < e^5/5), otherwise
// log(1+e) through MUFU.LG2.  Relative error stays < 1e-5 over the whole range (delta feeds exp(delta*A), so it is the
// relative error that matters).
__device__ __forceinline__ float softplus_f(float x) {
    const float e = __expf(fminf(x, 20.f));
    const float series = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.f);
    const float lg = __logf(1.f + e);
    const float r = e < 1.5e-2f ? series : lg;
    return x <= 20.f ? r : x;
}
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ float2 shfl_up2(float2 v, int delta, int width) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, delta, width), __shfl_up_sync(0xffffffffu, v.y, delta, width));
}
__device__ __forceinline__ float2 shfl_down2(float2 v, int delta, int width) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, delta, width),
                       __shfl_down_sync(0xffffffffu, v.y, delta, width));
}

// ---- shared-memory tile addressing ---------------------------------------------------------------
// A tile row holds TL fp32 tokens.  Thread j of a row-group reads T consecutive tokens with LDS.128; the
// 16-byte chunk index is XOR-swizzled so that the 8 lanes of a quarter-warp hit 8 distinct bank groups.
template <int T> __device__ __forceinline__ int swz_chunk(int c) {
    static_assert(T == 4 || T == 8 || T == 16, "T");
    return c ^ ((c >> 3) & (T / 4 - 1));
}
template <int T> __device__ __forceinline__ int swz_tok(int tok) { return (swz_chunk<T>(tok >> 2) << 2) | (tok & 3); }

}  // namespace mmu

// Shared pieces of the v3 ("register-resident token lanes") selective-scan kernels for dstate <= 16.
//
// Decomposition: a lane owns 8 consecutive tokens of TWO channel rows and walks the states; everything
// per (row, token) - delta, delta*u, y, the S1/S2 sums of the backward - lives in that lane's registers and never crosses
// lanes.  The two rows ride in the two halves of packed fp32 instructions (FFMA2/FMUL2): h2 = (h_rowA, h_rowB), so B and C
// enter as scalar-broadcast operands straight from their natural (state, token) layout - no repacking, the tiles are
// cp.async copies of global memory.  LPR lanes (16) cover a chunk of 8*LPR tokens of a row pair, a warp covers 32/LPR row
// pairs; the chunk-level recurrence is closed by one LPR-lane shuffle scan of (prod a, h) per (row pair, state).
//
// Budget per SM, from profiles/r1_micro_summary.md: MUFU 2 clk per warp-instr, FMA pipe 0.5 clk per packed op,
// shared memory 128 B/clk of delivered bytes, SHFL 1 clk.
#pragma once
#include "common.cuh"
#ifndef MMU_TMA_TILE
#define MMU_TMA_TILE 0          // 1: fp32 B/C tiles by cp.async.bulk.tensor (scan3_fwd.cuh)
#endif

namespace mmu {

constexpr int kS3T = 8;   // tokens per lane

// ---- 8 consecutive elements -> 8 floats in logical-token order.  e[] is in memory order; REV: logical token i = element 7-i
// (the lane walks memory backwards: fused flip, mamba_simple.py:230).
template <bool REV> __device__ __forceinline__ void order8(const float (&e)[8], float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = e[REV ? 7 - i : i];
}

template <typename IN_T> struct Raw8;      // 8 elements as they sit in memory / shared memory
template <> struct Raw8<float> {
    static constexpr int kQuads = 2;       // 16-byte pieces
    static __device__ __forceinline__ void unpack(const uint4 (&q)[2], float (&e)[8]) {
        e[0] = __uint_as_float(q[0].x), e[1] = __uint_as_float(q[0].y), e[2] = __uint_as_float(q[0].z), e[3] = __uint_as_float(q[0].w);
        e[4] = __uint_as_float(q[1].x), e[5] = __uint_as_float(q[1].y), e[6] = __uint_as_float(q[1].z), e[7] = __uint_as_float(q[1].w);
    }
    static __device__ __forceinline__ void pack(const float (&e)[8], uint4 (&q)[2]) {
        q[0] = make_uint4(__float_as_uint(e[0]), __float_as_uint(e[1]), __float_as_uint(e[2]), __float_as_uint(e[3]));
        q[1] = make_uint4(__float_as_uint(e[4]), __float_as_uint(e[5]), __float_as_uint(e[6]), __float_as_uint(e[7]));
    }
};
template <> struct Raw8<__nv_bfloat16> {
    static constexpr int kQuads = 1;
    static __device__ __forceinline__ void unpack(const uint4 (&q)[1], float (&e)[8]) {
        const unsigned w[4] = {q[0].x, q[0].y, q[0].z, q[0].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            e[2 * k] = __uint_as_float(w[k] << 16);
            e[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void pack(const float (&e)[8], uint4 (&q)[1]) {
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(e[2 * k], e[2 * k + 1]);
            w[k] = *reinterpret_cast<const unsigned *>(&h);
        }
        q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// store 8 logical tokens to global (p = lowest address of the 8 elements)
template <typename IN_T, bool REV> __device__ __forceinline__ void store8(IN_T *p, const float (&v)[8]) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[REV ? 7 - i : i] = v[i];
    uint4 q[Raw8<IN_T>::kQuads];
    Raw8<IN_T>::pack(e, q);
#pragma unroll
    for (int k = 0; k < Raw8<IN_T>::kQuads; ++k) reinterpret_cast<uint4 *>(p)[k] = q[k];
}
template <typename IN_T, bool REV> __device__ __forceinline__ void load8_global(const IN_T *p, float (&v)[8]) {
    uint4 q[Raw8<IN_T>::kQuads];
#pragma unroll
    for (int k = 0; k < Raw8<IN_T>::kQuads; ++k) q[k] = reinterpret_cast<const uint4 *>(p)[k];
    float e[8];
    Raw8<IN_T>::unpack(q, e);
    order8<REV>(e, v);
}

// ---- cp.async (LDGSTS) ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }   // all but the newest group

// ---- TMA 1-D bulk copies (cp.async.bulk, UBLKCP) with mbarrier completion: the B/C tile --------------------------------------
// One warp issues the whole tile - a handful of bulk copies instead of ~1 100 16-byte LDGSTS spread over the CTA - and every
// thread waits on the mbarrier's phase.  (The per-thread input slots stay on cp.async: they are thread-private 16-byte pieces.)
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE_%=;\n\t"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n\t}" ::"r"(mbar),
        "r"(parity)
        : "memory");
}

// ---- fused scan order: my 8 consecutive logical tokens t0 .. t0+7 of a tensor that lives in NATURAL token order ----------------
// (the gate z, out, dout, dz; OrdMap in common.cuh).  The slot is the lane's usual landing slot (two 16-byte quads for fp32, one
// for 2-byte types), filled in a memory-friendly element order:
//   TWOROW, row-pair part:  [row 2r: cols c..c+3 | row 2r+1: cols c..c+3]  - two vector copies; token i = element (i&1)*4 + (i>>1)
//   TWOROW, odd tail row :  8 contiguous elements                            - token i = element i
//   NSLICES (fp32 only)  :  element i = token i, eight 4-byte copies at stride L/ns (2-byte types would need 2-byte copies: the
//                           host does not fuse them, mmu_scan_order_fusable)
template <typename IN_T> __device__ __forceinline__ void ord_issue8(const OrdMap &o, int t0, const IN_T *row, unsigned q0, unsigned q1) {
    constexpr bool kF32 = sizeof(IN_T) == 4;
    if (o.kind == MMU_ORDER_TWOROW) {
        if (t0 < o.even_tokens) {
            const int pair = t0 / (2 * o.W), rem = t0 - pair * 2 * o.W, base = 2 * pair * o.W + (rem >> 1);
            if constexpr (kF32) cp_async16(q0, row + base), cp_async16(q1, row + base + o.W);
            else cp_async8(q0, row + base), cp_async8(q0 + 8, row + base + o.W);
        } else {
            if constexpr (kF32) cp_async16(q0, row + t0), cp_async16(q1, row + t0 + 4);
            else cp_async16(q0, row + t0);
        }
    } else if constexpr (kF32) {        // NSLICES
        const int jj = o.ns_shift >= 0 ? t0 >> o.ns_shift : t0 / o.ns, s0 = t0 - jj * o.ns;
        const IN_T *src = row + (int64_t)s0 * o.Ls + jj;
#pragma unroll
        for (int i = 0; i < 8; ++i) cp_async4((i < 4 ? q0 : q1) + (i & 3) * 4, src + (int64_t)i * o.Ls);
    }
}
// e[] = the slot's 8 elements (memory-friendly order) -> v[] in logical token order
__device__ __forceinline__ void ord_to_tokens(const OrdMap &o, int t0, const float (&e)[8], float (&v)[8]) {
    const bool inter = o.kind == MMU_ORDER_TWOROW && t0 < o.even_tokens;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = inter ? e[(i & 1) * 4 + (i >> 1)] : e[i];
}
template <typename IN_T> __device__ __forceinline__ void ord_store8(const OrdMap &o, int t0, IN_T *row, const float (&v)[8]) {
    constexpr bool kF32 = sizeof(IN_T) == 4;
    if (o.kind == MMU_ORDER_TWOROW) {
        float e[8];
        const bool inter = t0 < o.even_tokens;
#pragma unroll
        for (int i = 0; i < 8; ++i) e[inter ? (i & 1) * 4 + (i >> 1) : i] = v[i];
        uint4 q[Raw8<IN_T>::kQuads];
        Raw8<IN_T>::pack(e, q);
        if (inter) {
            const int pair = t0 / (2 * o.W), rem = t0 - pair * 2 * o.W, base = 2 * pair * o.W + (rem >> 1);
            if constexpr (kF32) {
                *reinterpret_cast<uint4 *>(row + base) = q[0];
                *reinterpret_cast<uint4 *>(row + base + o.W) = q[1];
            } else {
                *reinterpret_cast<uint2 *>(row + base) = make_uint2(q[0].x, q[0].y);
                *reinterpret_cast<uint2 *>(row + base + o.W) = make_uint2(q[0].z, q[0].w);
            }
        } else {
#pragma unroll
            for (int k = 0; k < Raw8<IN_T>::kQuads; ++k) reinterpret_cast<uint4 *>(row + t0)[k] = q[k];
        }
    } else {                             // NSLICES: element-wise
        const int jj = o.ns_shift >= 0 ? t0 >> o.ns_shift : t0 / o.ns, s0 = t0 - jj * o.ns;
        IN_T *dst = row + (int64_t)s0 * o.Ls + jj;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[(int64_t)i * o.Ls] = Elem<IN_T>::from_f(v[i]);
    }
}

__device__ __forceinline__ float lg2_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus (F.softplus, threshold 20: selective_scan_fwd_kernel.cuh:153-156) and its derivative sigmoid(x), branch free:
//   e = exp(-|x|);  log1p(e) = lg2(1+e)*ln2, or the alternating series for small e (keeps the RELATIVE error ~1e-7 where
//   delta is tiny);  softplus = max(x,0) + log1p(e)   (for x > 20 this is x to fp32 precision, as the reference returns).
__device__ __forceinline__ float softplus3(float x, float &e_out) {
    const float e = ex2(-fabsf(x) * kLog2e);
    const float series = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.f);
    const float lg = lg2_fast(1.f + e) * 0.69314718056f;
    e_out = e;
    return fmaxf(x, 0.f) + (e < 1.5e-2f ? series : lg);
}
__device__ __forceinline__ float softplus3(float x) {
    float e;
    return softplus3(x, e);
}
// sigmoid(x) from e = exp(-|x|)
__device__ __forceinline__ float sigmoid_from_e(float x, float e) {
    const float r = rcp_fast(1.f + e);
    return x >= 0.f ? r : e * r;
}
__device__ __forceinline__ float sigmoid3(float x) { return rcp_fast(1.f + ex2(-x * kLog2e)); }

// ---- B / C tile: natural layout ---------------------------------------------------------------------------------------------
// fp32 rows of CH tokens in MEMORY order, one row per state, B rows then C rows.  A 16-byte pad after every 128 bytes makes
// the lanes' 32-byte reads (8 tokens) conflict free: quad q sits at byte 16*q + 16*(q/8).
template <int LPR> struct BcTile {
    static constexpr int CH = LPR * kS3T;
    static constexpr int kRowBytes = CH * 4 + (CH / 32) * 16;
    static constexpr int kBytes = 2 * 16 * kRowBytes;                     // B[16] | C[16]
    static __device__ __forceinline__ int quad_off(int q) { return 16 * q + 16 * (q >> 3); }
};

// Dense variant for the tensor-map (TMA 2-D tile) experiment (MMU_TMA_TILE): rows of CH fp32 tokens back to back, which is what a
// cp.async.bulk.tensor box without swizzle writes; the lanes' 32-byte reads then take 2-way bank conflicts (a 128B-swizzled box
// would put the eight 32-token boxes of a row on the same banks: 8-way).
template <int LPR> struct BcTileDense {
    static constexpr int CH = LPR * kS3T;
    static constexpr int kRowBytes = CH * 4;
    static constexpr int kBytes = 2 * 16 * kRowBytes;
    static __device__ __forceinline__ int quad_off(int q) { return 16 * q; }
};
// one 3-D box {CH tokens, 16 states, 1 batch element} of a (L, dstate, batch) tensor map -> shared memory, completion on the mbarrier
__device__ __forceinline__ void tma_tile_3d(unsigned dst, const void *tmap, int x, int y, int z, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(mbar)
                 : "memory");
}

// Tile of logical tokens [t0, t0 + CH): lane `lane` of the issuing warp owns row `lane` (rows 0..15 = B states, 16..31 = C states).
// fp32: the row is copied in 128-byte pieces into the padded layout (quad_off); 2-byte types: one copy of the raw row (widened
// later).  Pieces outside [0, L) are skipped (those lanes run with delta = 0).  The mbarrier must have been initialised with a
// count of 32: every lane arrives with its own byte count.
template <typename IN_T, int LPR, bool REV, bool WITH_C>
__device__ __forceinline__ void tile_bulk_issue(unsigned tile_or_raw, unsigned mbar, const IN_T *__restrict__ B_b, const IN_T *__restrict__ C_b,
                                                int64_t B_ns, int64_t C_ns, int N, int t0, int L, int lane) {
    using Tl = BcTile<LPR>;
    constexpr int CH = Tl::CH;
    constexpr bool kF32 = sizeof(IN_T) == 4;
    constexpr int PT = kF32 ? 32 : CH;                       // tokens per piece
    constexpr int ROWB = kF32 ? Tl::kRowBytes : CH * 2;
    const int m0 = REV ? L - t0 - CH : t0;                   // memory index of tile position 0 (may be < 0)
    const int st = lane & 15;
    const bool isC = lane >= 16, live = st < min(N, 16) && (WITH_C || !isC);
    const IN_T *src = (isC ? C_b + (int64_t)st * C_ns : B_b + (int64_t)st * B_ns);
    const unsigned dst = tile_or_raw + lane * ROWB;
    unsigned bytes = 0;
    int lo[CH / PT], n[CH / PT];
#pragma unroll
    for (int pc = 0; pc < CH / PT; ++pc) {
        const int a = max(m0 + pc * PT, 0), b = min(m0 + (pc + 1) * PT, L);
        lo[pc] = a, n[pc] = live && b > a ? b - a : 0;
        bytes += (unsigned)n[pc] * (unsigned)sizeof(IN_T);
    }
    fence_proxy_async();                                     // the tile was read through the generic proxy until the barrier before this call
    mbar_arrive_expect_tx(mbar, bytes);
#pragma unroll
    for (int pc = 0; pc < CH / PT; ++pc) {
        if (n[pc] > 0) {
            const int off = lo[pc] - m0;                     // token offset inside the tile row (multiple of 8)
            const unsigned d = dst + (kF32 ? (unsigned)Tl::quad_off(off >> 2) : (unsigned)off * 2u);
            bulk_g2s(d, src + lo[pc], (unsigned)n[pc] * (unsigned)sizeof(IN_T), mbar);
        }
    }
}

// Issue the cp.async copies of the tile holding logical tokens [t0, t0 + CH).  Pieces outside the sequence are skipped: the
// lanes that would read them run with delta = 0, so the stale (finite: the tile is zero-initialised) contents do not matter.
// ROWB = bytes between rows in shared memory, EPP = elements per 16-byte piece, PPR = pieces per row; `dst_off(q)` maps a
// piece to its byte offset inside a row.
template <typename IN_T, int CH, int NT, int ROWB, bool REV, bool WITH_C, typename OFF>
__device__ __forceinline__ void rows_async(unsigned dst, const IN_T *__restrict__ B_b, const IN_T *__restrict__ C_b, int64_t B_ns,
                                           int64_t C_ns, int N, int t0, int L, int tid, OFF dst_off) {
    constexpr int EPP = 16 / (int)sizeof(IN_T), PPR = CH / EPP;
    constexpr int TPR = NT >= PPR ? PPR : NT, RPP = NT / TPR, QIT = PPR / TPR;
    const int m0 = REV ? L - t0 - CH : t0;                                // memory index of tile position 0 (may be < 0)
    const int q0 = tid % TPR, r0 = tid / TPR;
    if (r0 >= RPP) return;                                                // NT not a multiple of the row width: idle tail threads
#pragma unroll
    for (int qi = 0; qi < QIT; ++qi) {
        const int q = q0 + qi * TPR, m = m0 + EPP * q;
        if (m >= 0 && m < L) {
            const unsigned d = dst + r0 * ROWB + dst_off(q);
            const IN_T *sb = B_b + (int64_t)r0 * B_ns + m, *sc = C_b + (int64_t)r0 * C_ns + m;
#pragma unroll
            for (int k = 0; k < (16 + RPP - 1) / RPP; ++k) {
                if (r0 + k * RPP < min(N, 16)) {
                    cp_async16(d + k * RPP * ROWB, sb + (int64_t)k * RPP * B_ns);
                    if (WITH_C) cp_async16(d + (16 + k * RPP) * ROWB, sc + (int64_t)k * RPP * C_ns);
                }
            }
        }
    }
}

template <int LPR, int NT, bool REV, bool WITH_C>
__device__ __forceinline__ void tile_async_f32(unsigned tile, const float *__restrict__ B_b, const float *__restrict__ C_b, int64_t B_ns,
                                               int64_t C_ns, int N, int t0, int L, int tid) {
    using Tl = BcTile<LPR>;
    rows_async<float, Tl::CH, NT, Tl::kRowBytes, REV, WITH_C>(tile, B_b, C_b, B_ns, C_ns, N, t0, L, tid,
                                                               [](int q) { return Tl::quad_off(q); });
}

// bf16 inputs: cp.async the raw rows (CH*2 bytes each, B rows then C rows) into a staging buffer ...
template <int LPR, int NT, bool REV, bool WITH_C>
__device__ __forceinline__ void raw_async_bf16(unsigned raw, const __nv_bfloat16 *__restrict__ B_b, const __nv_bfloat16 *__restrict__ C_b,
                                               int64_t B_ns, int64_t C_ns, int N, int t0, int L, int tid) {
    constexpr int CH = LPR * kS3T;
    rows_async<__nv_bfloat16, CH, NT, CH * 2, REV, WITH_C>(raw, B_b, C_b, B_ns, C_ns, N, t0, L, tid, [](int q) { return 16 * q; });
}
// ... and widen them into the fp32 tile the lanes read.
template <int LPR, int NT, bool WITH_C>
__device__ __forceinline__ void widen_bf16_tile(unsigned char *tile, const unsigned char *raw, int tid) {
    using Tl = BcTile<LPR>;
    constexpr int CH = Tl::CH, PPR = CH / 8;
    for (int idx = tid; idx < (WITH_C ? 2 : 1) * 16 * PPR; idx += NT) {
        const int q = idx % PPR, rn = idx / PPR;
        const uint4 w = *reinterpret_cast<const uint4 *>(raw + (rn * PPR + q) * 16);
        const uint4 v[1] = {w};
        float e[8];
        Raw8<__nv_bfloat16>::unpack(v, e);
        unsigned char *row = tile + rn * Tl::kRowBytes;
        *reinterpret_cast<float4 *>(row + Tl::quad_off(2 * q)) = make_float4(e[0], e[1], e[2], e[3]);
        *reinterpret_cast<float4 *>(row + Tl::quad_off(2 * q + 1)) = make_float4(e[4], e[5], e[6], e[7]);
    }
}

}  // namespace mmu

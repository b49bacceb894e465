// v4 selective-scan kernels for WIDE problems (dim >= 64, dstate <= 16, seqlen % 8 == 0, 16-byte aligned rows).
// Math: SURVEY.md Appendix A; replaces selective_scan_fwd_kernel / selective_scan_bwd_kernel
// (selective_scan_fwd_kernel.cuh:67-303, selective_scan_bwd_kernel.cuh:75-489).
//
// Decomposition ("rows in lanes"): a lane owns TWO channel rows (packed in the halves of FFMA2/FMUL2) and a warp owns 64 rows of
// one batch element over one L-segment.  Consequences:
//   * B_n[t] and C_n[t] are the same for every lane of the warp: they are warp-uniform shared-memory broadcasts (LDS.128 of 4
//     states) that enter the packed instructions as scalar-broadcast operands;
//   * every recurrence (h forward, e = a*dh backward) is carried IN the lane: no shuffle scan, no barrier, no predication;
//   * the forward keeps all 16 states of its rows in registers (16 independent chains per lane: enough ILP that two warps per
//     scheduler keep the MUFU pipe busy), walks the tokens of its segment in order and costs 1 MUFU + 2 packed FMA issue
//     slots per (row, token, state);
//   * the sequence is cut into segments for parallelism (64 rows per warp leave only batch*dim/64 warps otherwise): an aggregate
//     pass computes each segment's end state from zero, a tiny kernel chains them, the main pass starts from the true state
//     (the existing chain kernels of scan3 are reused: same [row][segment][16] layout);
//   * the backward needs h_t and dh_t at the same token: its lanes hold 8 tokens of their two rows in registers (like scan3)
//     and walk the states, but take the state entering their 8 tokens from x, which the v4 forward saves after EVERY 8th
//     token (x stride 8 instead of 64: +N/8 floats per (row, token) of HBM traffic, paid because these kernels are MUFU /
//     FMA-pipe bound, not HBM bound - DESIGN.md 4.0), so there is no forward scan either; the reverse carry e is in registers.
//     dB / dC (sums over rows = over lanes) are reduced by a 16-shuffle recursive-halving transpose and one red.global per
//     (warp, 8 tokens, state): dim/64 atomics per element (the reference does dim, selective_scan_bwd_kernel.cuh:306-315).
#pragma once
#include "scan3.cuh"

namespace mmu {

constexpr int kS4W = 4;                 // warps per CTA; every warp is an independent work item
constexpr int kS4Rows = 64;             // rows per warp
constexpr int kS4MaxSeg = 512;          // segments per sequence (workspace bound)

// The lanes of a warp read 32-byte (fp32) / 16-byte (bf16) pieces of 64 different rows per stage; asking L2 to fetch 256 B
// around each piece turns the DRAM side into 256-byte runs per row (the next 7 / 15 stages then hit L2).
#ifndef MMU_V4_L2PF
#define MMU_V4_L2PF 256
#endif
#define MMU_STR2(x) #x
#define MMU_STR(x) MMU_STR2(x)
__device__ __forceinline__ void cp_async16_pf(unsigned dst, const void *src) {
#if MMU_V4_L2PF
    asm volatile("cp.async.cg.shared.global.L2::" MMU_STR(MMU_V4_L2PF) "B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4_pf(unsigned dst, const void *src) {
#if MMU_V4_L2PF
    asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ uint4 ldg16_pf(const void *src) {
    uint4 v;
#if MMU_V4_L2PF
    asm volatile("ld.global.nc.L2::" MMU_STR(MMU_V4_L2PF) "B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src));
#else
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src));
#endif
    return v;
}

struct Fwd4Args {
    const void *u, *delta, *z, *Bm, *Cm;
    const float *A, *Dv, *dbias;
    void *out, *ysave;
    float *x, *last_state;
    float *seg_hend, *seg_dsum;         // aggregate pass outputs: [b*D + row][nseg][16], [b*D + row][nseg]
    const float *hin;                   // main pass: state entering each segment (same layout), NULL when nseg == 1
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, y_bs, y_ds, B_bs, B_ns, C_bs, C_ns;
    int B, D, L, N;
    int nseg, sps, nstage;              // sps = 8-token stages per segment, nstage = L / 8
    int nrg;                            // row groups of 64
    int xs8, nx;                        // x holds the state after every xs8-th stage (1 or 8); nx = rows of x per (b, d)
    int softplus;
    int nitems;                         // warps with work
};

// shared memory of one warp: element slots [2 parity][NTEN tensors][2 rows][NQ quads][32 lanes] x 16 B, then the B/C tile
// [8 tokens][32] fp32 (B states 0..15 | C states 0..15 of a token are 128 contiguous bytes)
template <typename IN_T, int NTEN> struct S4Fwd {
    static constexpr int NQ = Raw8<IN_T>::kQuads;
    static constexpr int kSlotBytes = 2 * NTEN * 2 * NQ * 32 * 16;
    static constexpr int kTileBytes = 8 * 32 * 4;
    static constexpr int kWarpBytes = kSlotBytes + kTileBytes;
};

template <typename IN_T, bool REV, bool AGG>
__global__ void __launch_bounds__(32 * kS4W, 3) scan4_fwd_kernel(const __grid_constant__ Fwd4Args p) {
    constexpr int T = 8, NTEN = AGG ? 2 : 3;
    using Sm = S4Fwd<IN_T, NTEN>;
    constexpr int NQ = Sm::NQ, EPQ = 16 / (int)sizeof(IN_T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kS4W + warp;
    if (item >= p.nitems) return;
    const int rg = item % p.nrg, it1 = item / p.nrg, b = it1 % p.B, seg = it1 / p.B;
    const int D = p.D, L = p.L, N = p.N;
    const bool has_z = !AGG && p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *s_slot = smem_raw + (size_t)warp * Sm::kWarpBytes;
    float *s_tile = reinterpret_cast<float *>(s_slot + Sm::kSlotBytes);
    const unsigned slot_u32 = smem_u32(s_slot) + lane * 16;
    const unsigned char *slot_t = s_slot + lane * 16;

    // ---- my two rows -------------------------------------------------------------------------------------------------------------
    int row[2];
    bool row_ok[2];
    float2 A2[16], h[16];
    float bias[2], Dsk[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = rg * kS4Rows + lane + 32 * r;
        row_ok[r] = rr < D;
        row[r] = min(rr, D - 1);
        bias[r] = p.dbias != nullptr ? p.dbias[row[r]] : 0.f;
        Dsk[r] = (!AGG && p.Dv != nullptr) ? p.Dv[row[r]] : 0.f;
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        float a[2], h0[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            a[r] = n < N ? p.A[(int64_t)row[r] * N + n] * kLog2e : 0.f;
            h0[r] = 0.f;
            if (!AGG && p.hin != nullptr && seg > 0 && n < N) h0[r] = p.hin[(((int64_t)b * D + row[r]) * p.nseg + seg) * 16 + n];
        }
        A2[n] = make_float2(a[0], a[1]);
        h[n] = make_float2(h0[0], h0[1]);
    }
    const int s_begin = seg * p.sps, s_end = min(p.nstage, s_begin + p.sps);
    // memory index of logical stage s (8 consecutive tokens, walked backwards when REV)
    auto moff = [&](int s) { return REV ? L - T * (s + 1) : T * s; };
    const IN_T *src[NTEN][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        src[0][r] = reinterpret_cast<const IN_T *>(p.u) + (int64_t)b * p.u_bs + (int64_t)row[r] * p.u_ds;
        src[1][r] = reinterpret_cast<const IN_T *>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)row[r] * p.dl_ds;
        if constexpr (!AGG) src[2][r] = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs + (int64_t)row[r] * p.z_ds : nullptr;
    }
    // B/C tile: lane l stages row l (B states 0..15, C states 0..15) through registers
    const int tn = lane & 15;
    const bool t_live = tn < N && (!AGG || lane < 16);
    const IN_T *t_src = lane < 16 ? reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)tn * p.B_ns
                                  : reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)tn * p.C_ns;
    uint4 treg[NQ];
    auto tile_ldg = [&](int s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            treg[q] = t_live ? ldg16_pf(reinterpret_cast<const uint4 *>(t_src + moff(s)) + q) : make_uint4(0u, 0u, 0u, 0u);
    };
    auto tile_sts = [&]() {
        float e[8];
        Raw8<IN_T>::unpack(treg, e);
#pragma unroll
        for (int k = 0; k < 8; ++k) s_tile[(REV ? 7 - k : k) * 32 + lane] = e[k];
    };
    auto issue_elems = [&](int s, int par) {
        const int mo = moff(s);
#pragma unroll
        for (int t = 0; t < NTEN; ++t) {
            if (t == 2 && !has_z) continue;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    cp_async16_pf(slot_u32 + ((((par * NTEN + t) * 2 + r) * NQ + q) * 32) * 16, src[t][r] + mo + q * EPQ);
        }
    };
    auto load_slot = [&](int par, int t, int r, float (&v)[T]) {
        uint4 q[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(slot_t + ((((par * NTEN + t) * 2 + r) * NQ + k) * 32) * 16);
        float e[8];
        Raw8<IN_T>::unpack(q, e);
        order8<REV>(e, v);
    };

    issue_elems(s_begin, 0);
    cp_async_commit();
    tile_ldg(s_begin);
    tile_sts();
    float dsum[2] = {0.f, 0.f};

    for (int s = s_begin; s < s_end; ++s) {
        const int par = (s - s_begin) & 1;
        cp_async_wait_all();            // my slots of stage s have landed (they are lane private)
        __syncwarp();                   // tile of stage s is visible
        if (s + 1 < s_end) {
            issue_elems(s + 1, par ^ 1);
            tile_ldg(s + 1);
        }
        cp_async_commit();

        float uu[2][T], dd[2][T];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            load_slot(par, 0, r, uu[r]);
            load_slot(par, 1, r, dd[r]);
#pragma unroll
            for (int i = 0; i < T; ++i) {
                const float xx = dd[r][i] + bias[r];
                dd[r][i] = sp ? softplus3(xx) : xx;
                if (AGG) dsum[r] += dd[r][i];
            }
        }
        float2 ya[T];
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const float2 dl = make_float2(dd[0][i], dd[1][i]);
            const float2 dlu = make_float2(dd[0][i] * uu[0][i], dd[1][i] * uu[1][i]);
            float2 y0 = make_float2(Dsk[0] * uu[0][i], Dsk[1] * uu[1][i]), y1 = make_float2(0.f, 0.f);
            const float4 *tb = reinterpret_cast<const float4 *>(s_tile + i * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 b4 = tb[g];
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
                float cv[4] = {0.f, 0.f, 0.f, 0.f};
                if (!AGG) {
                    const float4 c4 = tb[4 + g];
                    cv[0] = c4.x, cv[1] = c4.y, cv[2] = c4.z, cv[3] = c4.w;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int n = 4 * g + k;
                    const float2 a = ex2(fmul2(dl, A2[n]));
                    h[n] = ffma2(a, h[n], fmul2(dlu, splat(bv[k])));
                    if (!AGG) {
                        if (k & 1) y1 = ffma2(h[n], splat(cv[k]), y1);
                        else y0 = ffma2(h[n], splat(cv[k]), y0);
                    }
                }
            }
            if (!AGG) ya[i] = fadd2(y0, y1);
        }
        __syncwarp();                   // every lane is done with the tile of stage s
        if (s + 1 < s_end) tile_sts();

        if constexpr (!AGG) {
            const int mo = moff(s);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float yv[T];
#pragma unroll
                for (int i = 0; i < T; ++i) yv[i] = r ? ya[i].y : ya[i].x;
                if (row_ok[r]) {
                    if (p.ysave != nullptr)
                        store8<IN_T, REV>(reinterpret_cast<IN_T *>(p.ysave) + (int64_t)b * p.y_bs + (int64_t)row[r] * p.y_ds + mo, yv);
                    if (has_z) {
                        float zz[T];
                        load_slot(par, 2, r, zz);
#pragma unroll
                        for (int i = 0; i < T; ++i) yv[i] *= zz[i] * sigmoid3(zz[i]);
                    }
                    store8<IN_T, REV>(reinterpret_cast<IN_T *>(p.out) + (int64_t)b * p.o_bs + (int64_t)row[r] * p.o_ds + mo, yv);
                }
            }
            // state after stage s -> x[b][row][k][n]
            if (p.x != nullptr && (s + 1) % p.xs8 == 0) {
                const int k = (s + 1) / p.xs8 - 1;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (!row_ok[r]) continue;
                    float *xp = p.x + (((int64_t)b * D + row[r]) * p.nx + k) * N;
                    if (N == 16) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            reinterpret_cast<float4 *>(xp)[g] = r ? make_float4(h[4 * g].y, h[4 * g + 1].y, h[4 * g + 2].y, h[4 * g + 3].y)
                                                                  : make_float4(h[4 * g].x, h[4 * g + 1].x, h[4 * g + 2].x, h[4 * g + 3].x);
                    } else {
#pragma unroll
                        for (int n = 0; n < 16; ++n)
                            if (n < N) xp[n] = r ? h[n].y : h[n].x;
                    }
                }
            }
        }
    }

    if constexpr (AGG) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!row_ok[r]) continue;
            const int64_t o = ((int64_t)b * D + row[r]) * p.nseg + seg;
            float4 *hp = reinterpret_cast<float4 *>(p.seg_hend + o * 16);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                hp[g] = r ? make_float4(h[4 * g].y, h[4 * g + 1].y, h[4 * g + 2].y, h[4 * g + 3].y)
                          : make_float4(h[4 * g].x, h[4 * g + 1].x, h[4 * g + 2].x, h[4 * g + 3].x);
            p.seg_dsum[o] = dsum[r];
        }
    } else if (p.last_state != nullptr && seg == p.nseg - 1) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!row_ok[r]) continue;
#pragma unroll
            for (int n = 0; n < 16; ++n)
                if (n < N) p.last_state[((int64_t)b * D + row[r]) * N + n] = r ? h[n].y : h[n].x;
        }
    }
}

}  // namespace mmu

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

// ---- row-block staging ---------------------------------------------------------------------------------------------------------
// A lane that fetched its own 16 bytes of its own row would make every warp load touch 32 different cache lines for 16 useful
// bytes each: the first version of these kernels was bound by that (24 M line requests per forward at config 2, bf16 no faster
// than fp32).  Instead the warp moves a BLOCK of its 64 rows x (P pieces of 16 bytes) cooperatively: P consecutive lanes take the
// P consecutive pieces of one row, so a warp instruction touches 32/P lines with P*16 contiguous bytes each (64 B for fp32:
// 16 tokens = two 8-token stages).  Piece (row, p) lives at 16-byte unit  ((row >> 5) * P + p) * 32 + (((row & 31) + p * (8 / P)) & 31):
// the owner lane (row & 31) reads piece p with a conflict-free LDS.128, and the rotation by p * (8 / P) keeps the writers of
// one instruction (32/P rows x P pieces) on distinct bank groups too.
template <int P> __device__ __forceinline__ int rb_unit(int row, int p) { return ((row >> 5) * P + p) * 32 + (((row & 31) + p * (8 / P)) & 31); }

// global -> shared (cp.async).  g0 = address of piece 0 of the warp's row 0; rows valid: row < nrows; pieces valid: p < npieces.
template <int P> __device__ __forceinline__ void rb_load_async(unsigned s_u32, const char *g0, int64_t row_bytes, int nrows, int npieces, int lane) {
#pragma unroll 1        // (unrolled, the compiler keeps every piece's address in registers across the whole kernel)
    for (int it = 0; it < 2 * P; ++it) {
        const int id = it * 32 + lane, row = id / P, p = id % P;
        if (row < nrows && p < npieces) cp_async16_pf(s_u32 + rb_unit<P>(row, p) * 16, g0 + (int64_t)row * row_bytes + p * 16);
    }
}
// shared -> global
template <int P> __device__ __forceinline__ void rb_store(const unsigned char *s, char *g0, int64_t row_bytes, int nrows, int npieces, int lane) {
#pragma unroll 2
    for (int it = 0; it < 2 * P; ++it) {
        const int id = it * 32 + lane, row = id / P, p = id % P;
        if (row < nrows && p < npieces)
            *reinterpret_cast<uint4 *>(g0 + (int64_t)row * row_bytes + p * 16) = *reinterpret_cast<const uint4 *>(s + rb_unit<P>(row, p) * 16);
    }
}

constexpr int kS4SB = 2;                // 8-token stages per block

// shared memory of one warp: blocks [2 parity][NTEN tensors][64 rows x P pieces] x 16 B, then the B/C tile [8 tokens][32] fp32
// (B states 0..15 | C states 0..15 of a token are 128 contiguous bytes)
template <typename IN_T, int NTEN> struct S4Fwd {
    static constexpr int NQ = Raw8<IN_T>::kQuads;          // pieces per row per stage
    static constexpr int P = kS4SB * NQ;                   // pieces per row per block
    static constexpr int kBlkBytes = 64 * P * 16;          // one tensor, one block
    static constexpr int kSlotBytes = 2 * NTEN * kBlkBytes;
    static constexpr int kTileBytes = 8 * 32 * 4;
    static constexpr int kWarpBytes = kSlotBytes + kTileBytes;
};

template <typename IN_T, bool REV, bool AGG>
__global__ void __launch_bounds__(32 * kS4W, (AGG || sizeof(IN_T) == 2) ? 3 : 2) scan4_fwd_kernel(const __grid_constant__ Fwd4Args p) {
    constexpr int T = 8, NTEN = AGG ? 2 : 3, SB = kS4SB;
    using Sm = S4Fwd<IN_T, NTEN>;
    constexpr int NQ = Sm::NQ, P = Sm::P, ES = (int)sizeof(IN_T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kS4W + warp;
    if (item >= p.nitems) return;
    const int rg = item % p.nrg, it1 = item / p.nrg, b = it1 % p.B, seg = it1 / p.B;
    const int D = p.D, L = p.L, N = p.N;
    const bool has_z = !AGG && p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_slot = smem_raw + (size_t)warp * Sm::kWarpBytes;
    float *s_tile = reinterpret_cast<float *>(s_slot + Sm::kSlotBytes);
    const unsigned slot_u32 = smem_u32(s_slot);

    // ---- my two rows -------------------------------------------------------------------------------------------------------------
    const int row0 = rg * kS4Rows, nrows = min(kS4Rows, D - row0);          // the warp's rows
    int row[2];
    bool row_ok[2];
    float2 A2[16], h[16];
    float bias[2], Dsk[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = row0 + lane + 32 * r;
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
    const int nblk = (s_end - s_begin + SB - 1) / SB;
    // memory index of the first element of logical stages [s, s + nst) (walked backwards when REV)
    auto moff = [&](int s, int nst) { return REV ? L - T * (s + nst) : T * s; };
    // row 0 of the warp, per tensor (bytes)
    const char *g_in[NTEN];
    int64_t rs_in[NTEN];
    g_in[0] = reinterpret_cast<const char *>(p.u) + ((int64_t)b * p.u_bs + (int64_t)row0 * p.u_ds) * ES, rs_in[0] = p.u_ds * ES;
    g_in[1] = reinterpret_cast<const char *>(p.delta) + ((int64_t)b * p.dl_bs + (int64_t)row0 * p.dl_ds) * ES, rs_in[1] = p.dl_ds * ES;
    if constexpr (!AGG) g_in[2] = reinterpret_cast<const char *>(p.z) + ((int64_t)b * p.z_bs + (int64_t)row0 * p.z_ds) * ES, rs_in[2] = p.z_ds * ES;
    // B/C tile: lane l stages row l (B states 0..15, C states 0..15) through registers
    const int tn = lane & 15;
    const bool t_live = tn < N && (!AGG || lane < 16);
    const IN_T *t_src = lane < 16 ? reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)tn * p.B_ns
                                  : reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)tn * p.C_ns;
    uint4 treg[NQ];
    auto tile_ldg = [&](int s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            treg[q] = t_live ? ldg16_pf(reinterpret_cast<const uint4 *>(t_src + moff(s, 1)) + q) : make_uint4(0u, 0u, 0u, 0u);
    };
    auto tile_sts = [&]() {
        float e[8];
        Raw8<IN_T>::unpack(treg, e);
#pragma unroll
        for (int k = 0; k < 8; ++k) s_tile[(REV ? 7 - k : k) * 32 + lane] = e[k];
    };
    auto issue_block = [&](int blk, int par) {
        const int s0 = s_begin + blk * SB, nst = min(SB, s_end - s0);
        const int64_t mo = (int64_t)moff(s0, nst) * ES;
#pragma unroll
        for (int t = 0; t < NTEN; ++t) {
            if (t == 2 && !has_z) continue;
            rb_load_async<P>(slot_u32 + (par * NTEN + t) * Sm::kBlkBytes, g_in[t] + mo, rs_in[t], nrows, nst * NQ, lane);
        }
    };
    // my 8 tokens of stage i (memory stage ms inside the block) of tensor t, row r
    auto load_stage = [&](int par, int t, int r, int ms, float (&v)[T]) {
        uint4 q[NQ];
        const unsigned char *base = s_slot + (par * NTEN + t) * Sm::kBlkBytes;
#pragma unroll
        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(base + rb_unit<P>(lane + 32 * r, ms * NQ + k) * 16);
        float e[8];
        Raw8<IN_T>::unpack(q, e);
        order8<REV>(e, v);
    };
    auto store_stage = [&](int par, int t, int r, int ms, const float (&v)[T]) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[REV ? 7 - i : i] = v[i];
        uint4 q[NQ];
        Raw8<IN_T>::pack(e, q);
        unsigned char *base = s_slot + (par * NTEN + t) * Sm::kBlkBytes;
#pragma unroll
        for (int k = 0; k < NQ; ++k) *reinterpret_cast<uint4 *>(base + rb_unit<P>(lane + 32 * r, ms * NQ + k) * 16) = q[k];
    };

    issue_block(0, 0);
    cp_async_commit();
    tile_ldg(s_begin);
    tile_sts();
    float dsum[2] = {0.f, 0.f};

    for (int blk = 0; blk < nblk; ++blk) {
        const int par = blk & 1;
        const int s0 = s_begin + blk * SB, nst = min(SB, s_end - s0);
        cp_async_wait_all();            // my pieces of this block have landed ...
        __syncwarp();                   // ... and everybody else's (the rows are fetched cooperatively)
        if (blk + 1 < nblk) issue_block(blk + 1, par ^ 1);
        cp_async_commit();

#pragma unroll 1
        for (int i = 0; i < nst; ++i) {
            const int s = s0 + i, ms = REV ? nst - 1 - i : i;
            __syncwarp();               // tile of stage s is visible
            if (s + 1 < s_end) tile_ldg(s + 1);

            float uu[2][T], dd[2][T];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                load_stage(par, 0, r, ms, uu[r]);
                load_stage(par, 1, r, ms, dd[r]);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const float xx = dd[r][k] + bias[r];
                    dd[r][k] = sp ? softplus3(xx) : xx;
                    if (AGG) dsum[r] += dd[r][k];
                }
            }
            float2 ya[T];
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const float2 dl = make_float2(dd[0][k], dd[1][k]);
                const float2 dlu = make_float2(dd[0][k] * uu[0][k], dd[1][k] * uu[1][k]);
                float2 y0 = make_float2(Dsk[0] * uu[0][k], Dsk[1] * uu[1][k]), y1 = make_float2(0.f, 0.f);
                const float4 *tb = reinterpret_cast<const float4 *>(s_tile + k * 32);
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
                    for (int j = 0; j < 4; ++j) {
                        const int n = 4 * g + j;
                        const float2 a = ex2(fmul2(dl, A2[n]));
                        h[n] = ffma2(a, h[n], fmul2(dlu, splat(bv[j])));
                        if (!AGG) {
                            if (j & 1) y1 = ffma2(h[n], splat(cv[j]), y1);
                            else y0 = ffma2(h[n], splat(cv[j]), y0);
                        }
                    }
                }
                if (!AGG) ya[k] = fadd2(y0, y1);
            }
            __syncwarp();               // every lane is done with the tile of stage s
            if (s + 1 < s_end) tile_sts();

            if constexpr (!AGG) {
                // y goes into the (consumed) u piece of this stage, out into the delta piece: the block is stored cooperatively below
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float yv[T];
#pragma unroll
                    for (int k = 0; k < T; ++k) yv[k] = r ? ya[k].y : ya[k].x;
                    if (p.ysave != nullptr) store_stage(par, 0, r, ms, yv);
                    if (has_z) {
                        float zz[T];
                        load_stage(par, 2, r, ms, zz);
#pragma unroll
                        for (int k = 0; k < T; ++k) yv[k] *= zz[k] * sigmoid3(zz[k]);
                    }
                    store_stage(par, 1, r, ms, yv);
                }
                // state after stage s -> x[b][row][k][n]
                if (p.x != nullptr && (s + 1) % p.xs8 == 0) {
                    const int kx = (s + 1) / p.xs8 - 1;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (!row_ok[r]) continue;
                        float *xp = p.x + (((int64_t)b * D + row[r]) * p.nx + kx) * N;
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
        if constexpr (!AGG) {
            __syncwarp();               // the block's y / out pieces are complete
            const int64_t mo = (int64_t)moff(s0, nst) * ES;
            if (p.ysave != nullptr)
                rb_store<P>(s_slot + (par * NTEN + 0) * Sm::kBlkBytes, reinterpret_cast<char *>(p.ysave) + ((int64_t)b * p.y_bs + (int64_t)row0 * p.y_ds) * ES + mo,
                            p.y_ds * ES, nrows, nst * NQ, lane);
            rb_store<P>(s_slot + (par * NTEN + 1) * Sm::kBlkBytes, reinterpret_cast<char *>(p.out) + ((int64_t)b * p.o_bs + (int64_t)row0 * p.o_ds) * ES + mo,
                        p.o_ds * ES, nrows, nst * NQ, lane);
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

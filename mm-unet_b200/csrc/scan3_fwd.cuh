// Selective scan forward, v3 kernel (dstate <= 16, seqlen % 8 == 0, 16-byte aligned rows).  See scan3.cuh for the decomposition.
// Math: SURVEY.md Appendix A; replaces selective_scan_fwd_kernel (selective_scan_fwd_kernel.cuh:67-303).
//
// Per chunk (CH = 8*LPR tokens) and state n, lane j (tokens 8j..8j+7 of the row pair) computes, from a zero state,
//     H_i = a_i H_{i-1} + delta_i u_i B_i,   Pc_i = prod_{k<=i} a_k,   y_i += C_i H_i,   cp_i = C_i Pc_i          (packed: 2 rows)
// then the lanes' (Pc_7, H_7) are combined by an inclusive shuffle scan (lane 0 first absorbs the state entering the chunk),
// which gives every lane the true state hs before its first token, and y_i += cp_i * hs closes the recurrence without a
// second serial pass.  One MUFU.EX2 per (row, token, state).
#pragma once
#include <type_traits>

#include "scan3.cuh"

// MMU_BULK_TILE=1 loads the B/C tile with TMA 1-D bulk copies (cp.async.bulk + mbarrier, scan3.cuh:tile_bulk_issue) issued by
// one warp instead of per-thread cp.async.  Measured on B200 at config 2 (forward, us): fp32 265 vs 175, bf16 156 vs 150; RCG
// L = 65 536 bf16 1 966 vs 1 799 - the padded (bank-conflict-free) row layout forces 128-byte pieces (256 bulk copies per fp32 tile),
// too small for the bulk path, so the default stays cp.async.  A tensor-map copy (cp.async.bulk.tensor, 128B swizzle) is the
// version that can win: one instruction per tile and no padding.
// (Also measured and NOT kept in this kernel, unlike the backward: two cp.async groups per chunk and pointers advanced before
// use - together 173 -> 180 us here.)
#ifndef MMU_BULK_TILE
#define MMU_BULK_TILE 0
#endif
// MMU_TMA_TILE=1 (fp32 only, experiment): the B/C tile of a chunk is TWO cp.async.bulk.tensor copies (tensor maps built by the host
// per call, UTMALDG in the SASS) into a dense tile instead of ~1 100 per-thread LDGSTS into the padded one.
// experiment knobs (csrc/build.sh MMU_VARIANT builds): registers per thread the launch bounds aim at, state-loop unroll
#ifndef MMU_FWD3_REGS
#define MMU_FWD3_REGS 168
#endif
#ifndef MMU_FWD3_UNROLL
#define MMU_FWD3_UNROLL 2
#endif
#define MMU_PRAGMA_(x) _Pragma(#x)
#define MMU_UNROLL(n) MMU_PRAGMA_(unroll n)

namespace mmu {

struct Fwd3Args {
    const void *u, *delta, *z, *Bm, *Cm;
    const float *A, *Dv, *dbias;
    void *out, *ysave;             // out = y*silu(z) (or y when z == NULL); ysave = pre-gate y (optional)
    float *x, *last_state;
    float *seg_hend, *seg_dsum;    // AGG pass outputs
    const float *hin;              // state entering each segment (main pass, nseg > 1)
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, y_bs, y_ds, B_bs, B_ns, C_bs, C_ns;
    int B, D, L, N;
    int nseg, cps, nchunks, nx;    // cps = chunks per segment
    int softplus;
    OrdMap ord;                    // ORD kernels: the gate z and the output live at ord(l) (natural token order), everything else at l
#if MMU_TMA_TILE
    alignas(64) unsigned char tmB[128], tmC[128];   // CUtensorMap of B and C: (L, dstate, batch) fp32, box {256, 16, 1}
#endif
};

template <typename IN_T, int LPR, int W> struct Fwd3Cfg {
    static constexpr bool kF32 = sizeof(IN_T) == 4;
    static constexpr int CH = LPR * kS3T, RPW = 32 / LPR, RW = 2 * RPW, R = W * RW, NT = 32 * W, NRP = R / 2;
    static constexpr int NCK = CH / MMU_STATE_STRIDE;                     // saved states per chunk (one per 8 lanes)
    static constexpr int kRawBytes = kF32 ? 0 : 2 * 16 * CH * 2;
    static constexpr int NQ = Raw8<IN_T>::kQuads;
    static constexpr int kElemBytes = 3 * 2 * NQ * NT * 16;               // u | delta | z : [tensor][row][quad][thread] x 16 B
    static constexpr int kTabBytes = (1 + NCK) * NRP * 16 * (int)sizeof(float2);   // A*log2e | states after every 64 tokens
    static constexpr size_t smem_bytes = (size_t)BcTile<LPR>::kBytes + kRawBytes + kElemBytes + kTabBytes;
};

// ORD: fused scan order (NSLICES / TWOROW) - z is gathered and out scattered through p.ord (ord_issue8 / ord_store8 in scan3.cuh);
// never together with REV or AGG.
template <typename IN_T, int LPR, int W, bool REV, bool AGG, bool ORD = false>
__global__ void __launch_bounds__(32 * W, AGG ? 1 : (65536 / (32 * W * MMU_FWD3_REGS))) scan3_fwd_kernel(const __grid_constant__ Fwd3Args p) {
    static_assert(!ORD || (!REV && !AGG), "ordered gate / output: forward direction, main pass only");
    using Cfg = Fwd3Cfg<IN_T, LPR, W>;
    constexpr bool kF32 = Cfg::kF32;
    constexpr bool kTmaTile = MMU_TMA_TILE != 0 && kF32 && LPR == 32;   // B/C tile by two tensor-map copies into a dense tile
    using Tl = typename std::conditional<kTmaTile, BcTileDense<LPR>, BcTile<LPR>>::type;
    constexpr int CH = Cfg::CH, RPW = Cfg::RPW, R = Cfg::R, NT = Cfg::NT, NRP = Cfg::NRP, T = kS3T, NQ = Cfg::NQ, NCK = Cfg::NCK;
    constexpr int EPQ = 16 / (int)sizeof(IN_T);          // elements per 16-byte piece
    constexpr bool kBulkTile = MMU_BULK_TILE != 0 && !kTmaTile;   // B/C tile by cp.async.bulk (TMA) instead of per-thread cp.async
    constexpr int NSTEP = LPR == 32 ? 5 : (LPR == 16 ? 4 : 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rp = lane / LPR, j = lane % LPR;
    const int b = blockIdx.y, row0 = blockIdx.x * R, seg = blockIdx.z;
    const int D = p.D, L = p.L, N = p.N;
    const int rpg = warp * RPW + rp;                    // row pair inside the CTA
    const bool has_z = p.z != nullptr, sp = p.softplus != 0;
    const int NS = min(16, (N + 1) & ~1);               // states walked (two at a time)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_tile = smem_raw;                                                  // [Tl::kBytes] (the allocation is sized for the padded tile)
    unsigned char *s_rawbc = s_tile + BcTile<LPR>::kBytes;                             // bf16 only
    unsigned char *s_elem = s_rawbc + Cfg::kRawBytes;                                  // [3][2][NQ][NT] x 16 B
    float2 *s_A = reinterpret_cast<float2 *>(s_elem + Cfg::kElemBytes);                // [NRP][16]  A*log2e of (row A, row B)
    float2 *s_ck = s_A + NRP * 16;                                                     // [NRP][NCK][16]  state after every 64th token;
                                                                                       // slot NCK-1 = state entering the next chunk
    // ---- one-time initialisation -----------------------------------------------------------------------------------------
    for (int i = tid; i < (int)((BcTile<LPR>::kBytes + Cfg::kRawBytes + Cfg::kElemBytes) / 16); i += NT)
        reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < NRP * 16; i += NT) {
        const int g = i >> 4, n = i & 15;
        float a[2], h[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = row0 + 2 * g + r;
            const bool ok = row < D && n < N;
            a[r] = ok ? p.A[(int64_t)row * N + n] * kLog2e : 0.f;
            h[r] = 0.f;
            if (!AGG && ok && p.hin != nullptr && seg > 0) h[r] = p.hin[(((int64_t)b * D + row) * p.nseg + seg) * 16 + n];
        }
        s_A[i] = make_float2(a[0], a[1]);
        s_ck[(g * NCK + NCK - 1) * 16 + n] = make_float2(h[0], h[1]);
    }

    // ---- my two rows --------------------------------------------------------------------------------------------------------
    const int rowA = row0 + 2 * rpg;
    const int c_begin = seg * p.cps, c_end = min(p.nchunks, c_begin + p.cps);
    // memory index of my 8 tokens in chunk c_begin (advances by +-CH per chunk)
    int tl = c_begin * CH + T * j;
    const int mo0 = REV ? L - T - tl : tl;
    bool row_ok[2];
    const IN_T *u_p[2], *d_p[2], *z_p[2];               // point at my 8 tokens of the chunk being PREFETCHED
    IN_T *o_p[2], *y_p[2];                               // point at my 8 tokens of the chunk being COMPUTED
    float bias[2], Dsk[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        row_ok[r] = rowA + r < D;
        const int row = min(rowA + r, D - 1);
        u_p[r] = reinterpret_cast<const IN_T *>(p.u) + (int64_t)b * p.u_bs + (int64_t)row * p.u_ds + mo0;
        d_p[r] = reinterpret_cast<const IN_T *>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)row * p.dl_ds + mo0;
        z_p[r] = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs + (int64_t)row * p.z_ds + mo0 : nullptr;
        o_p[r] = AGG ? nullptr : reinterpret_cast<IN_T *>(p.out) + (int64_t)b * p.o_bs + (int64_t)row * p.o_ds + mo0;
        y_p[r] = (AGG || p.ysave == nullptr) ? nullptr
                                             : reinterpret_cast<IN_T *>(p.ysave) + (int64_t)b * p.y_bs + (int64_t)row * p.y_ds + mo0;
        bias[r] = p.dbias != nullptr ? p.dbias[row] : 0.f;
        Dsk[r] = p.Dv != nullptr ? p.Dv[row] : 0.f;
    }
    constexpr int STEP = REV ? -CH : CH;
    [[maybe_unused]] const IN_T *z_row[2];
    [[maybe_unused]] IN_T *o_row[2];
    if constexpr (ORD) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = min(rowA + r, D - 1);
            z_row[r] = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs + (int64_t)row * p.z_ds : nullptr;
            o_row[r] = reinterpret_cast<IN_T *>(p.out) + (int64_t)b * p.o_bs + (int64_t)row * p.o_ds;
        }
    }
    [[maybe_unused]] int tz = tl;        // ORD: first logical token of my 8 in the chunk whose z is being prefetched
    const IN_T *B_b = reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs;
    const IN_T *C_b = reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs;

    bool ge[NSTEP];
#pragma unroll
    for (int s = 0; s < NSTEP; ++s) ge[s] = j >= (1 << s);
    float dsum[2] = {0.f, 0.f};
    // my 8 tokens inside a tile row (memory order): two adjacent quads
    const int qa = REV ? 2 * (LPR - 1 - j) : 2 * j;
    const unsigned char *tile = s_tile + Tl::quad_off(qa);
    const unsigned s_tile_u32 = smem_u32(s_tile), s_raw_u32 = smem_u32(s_rawbc);
    const unsigned s_elem_u32 = smem_u32(s_elem) + tid * 16;
    const unsigned char *s_elem_t = s_elem + tid * 16;

    __shared__ __align__(8) unsigned long long s_mbar;       // used by the bulk-copy variant only
    const unsigned mbar = smem_u32(&s_mbar);
    [[maybe_unused]] unsigned tile_phase = 0;
    if constexpr (kBulkTile || kTmaTile) {
        if (tid == 0) {
            mbar_init(mbar, kTmaTile ? 1 : 32);
            fence_mbar_init();
        }
    }
    auto issue_tile = [&](int c) {
#if MMU_TMA_TILE
        if constexpr (kTmaTile) {       // one thread: expect the tile's bytes, then one box per tensor (out-of-range tokens / states read as 0)
            if (tid == 0) {
                const int m0 = REV ? L - (c + 1) * CH : c * CH;
                fence_proxy_async();
                mbar_arrive_expect_tx(mbar, (AGG ? 1u : 2u) * 16u * CH * 4u);
                tma_tile_3d(s_tile_u32, p.tmB, m0, 0, b, mbar);
                if (!AGG) tma_tile_3d(s_tile_u32 + 16 * Tl::kRowBytes, p.tmC, m0, 0, b, mbar);
            }
            return;
        }
#endif
        if constexpr (kBulkTile) {      // TMA bulk copies issued by warp 0, completion on the mbarrier
            if (warp == 0)
                tile_bulk_issue<IN_T, LPR, REV, !AGG>(kF32 ? s_tile_u32 : s_raw_u32, mbar, B_b, C_b, p.B_ns, p.C_ns, N, c * CH, L, lane);
            return;
        }
        if constexpr (kF32) {
            tile_async_f32<LPR, NT, REV, !AGG>(s_tile_u32, reinterpret_cast<const float *>(B_b), reinterpret_cast<const float *>(C_b),
                                               p.B_ns, p.C_ns, N, c * CH, L, tid);
        } else {
            raw_async_bf16<LPR, NT, REV, !AGG>(s_raw_u32, reinterpret_cast<const __nv_bfloat16 *>(B_b),
                                               reinterpret_cast<const __nv_bfloat16 *>(C_b), p.B_ns, p.C_ns, N, c * CH, L, tid);
        }
    };
    auto issue_ud = [&](bool in_seq) {      // u and delta of the chunk the prefetch pointers stand on, then advance them
        if (in_seq) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    cp_async16(s_elem_u32 + ((0 * 2 + r) * NQ + q) * NT * 16, u_p[r] + q * EPQ);
                    cp_async16(s_elem_u32 + ((1 * 2 + r) * NQ + q) * NT * 16, d_p[r] + q * EPQ);
                }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) u_p[r] += STEP, d_p[r] += STEP;
    };
    auto issue_z = [&](bool in_seq) {
        if constexpr (ORD) {
            if (has_z && in_seq) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    ord_issue8<IN_T>(p.ord, tz, z_row[r], s_elem_u32 + ((2 * 2 + r) * NQ) * NT * 16, s_elem_u32 + ((2 * 2 + r) * NQ + NQ - 1) * NT * 16);
            }
            tz += CH;
            return;
        }
        if (!AGG && has_z) {
            if (in_seq) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int q = 0; q < NQ; ++q) cp_async16(s_elem_u32 + ((2 * 2 + r) * NQ + q) * NT * 16, z_p[r] + q * EPQ);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) z_p[r] += STEP;
        }
    };
    auto load_elem = [&](int which, int r, float (&v)[T]) {      // my 8 tokens of tensor `which`, row r, from the staging area
        uint4 q[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(s_elem_t + ((which * 2 + r) * NQ + k) * NT * 16);
        float e[8];
        Raw8<IN_T>::unpack(q, e);
        order8<REV>(e, v);
    };

    __syncthreads();                    // zero fill and tables visible before the first copies land
    issue_tile(c_begin);
    issue_ud(tl < L);
    issue_z(tl < L);
    cp_async_commit();

    for (int c = c_begin; c < c_end; ++c, tl += CH) {
        const bool ok = tl < L;
        cp_async_wait_all();
        if constexpr (kBulkTile || kTmaTile) mbar_wait(mbar, tile_phase++ & 1u);
        __syncthreads();                // chunk c has landed
        if constexpr (!kF32) {
            widen_bf16_tile<LPR, NT, !AGG>(s_tile, s_rawbc, tid);
            __syncthreads();
        }
        // ---- per (row, token) registers, the two rows packed: .x = row A, .y = row B ------------------------------------
        float2 dl[T], dlu[T], ya[T];
        {
            float uu[2][T], dd[2][T];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                load_elem(0, r, uu[r]);
                load_elem(1, r, dd[r]);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float xx = dd[r][i] + bias[r];
                    const float v = sp ? softplus3(xx) : xx;
                    dd[r][i] = ok ? v : 0.f;
                    if (AGG) dsum[r] += dd[r][i];
                }
            }
#pragma unroll
            for (int i = 0; i < T; ++i) {
                dl[i] = make_float2(dd[0][i], dd[1][i]);
                dlu[i] = make_float2(dd[0][i] * uu[0][i], dd[1][i] * uu[1][i]);
                ya[i] = make_float2(Dsk[0] * uu[0][i], Dsk[1] * uu[1][i]);
            }
        }
        // u / delta of the next chunk (the staging slots are private to this thread and were just consumed)
        if (c + 1 < c_end) issue_ud(tl + CH < L);

MMU_UNROLL(MMU_FWD3_UNROLL)
        for (int n0 = 0; n0 < NS; n0 += 2) {
            const float4 A4 = *reinterpret_cast<const float4 *>(s_A + rpg * 16 + n0);
            const float4 car4 = *reinterpret_cast<const float4 *>(s_ck + (rpg * NCK + NCK - 1) * 16 + n0);
            float2 cp[2][T], H[2], P[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const unsigned char *rowB = tile + (n0 + s) * Tl::kRowBytes;
                float Bn[T], Cn[T];
                {
                    const float4 b0 = *reinterpret_cast<const float4 *>(rowB), b1 = *reinterpret_cast<const float4 *>(rowB + 16);
                    const float eb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    order8<REV>(eb, Bn);
                }
                if (!AGG) {
                    const unsigned char *rowC = rowB + 16 * Tl::kRowBytes;
                    const float4 c0 = *reinterpret_cast<const float4 *>(rowC), c1 = *reinterpret_cast<const float4 *>(rowC + 16);
                    const float ec[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    order8<REV>(ec, Cn);
                }
                const float2 A2 = s ? make_float2(A4.z, A4.w) : make_float2(A4.x, A4.y);
                float2 h = make_float2(0.f, 0.f), pc;
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float2 a = ex2(fmul2(dl[i], A2));
                    h = ffma2(a, h, fmul2(dlu[i], splat(Bn[i])));
                    pc = i == 0 ? a : fmul2(pc, a);
                    if (!AGG) {
                        ya[i] = ffma2(h, splat(Cn[i]), ya[i]);
                        cp[s][i] = fmul2(pc, splat(Cn[i]));
                    }
                }
                // lane 0 absorbs the state entering the chunk
                const float2 hc = s ? make_float2(car4.z, car4.w) : make_float2(car4.x, car4.y);
                const float2 h0 = ffma2(pc, hc, h);
                H[s] = j == 0 ? h0 : h;
                P[s] = pc;
            }
            // inclusive scan over the LPR lanes of my row pair
#pragma unroll
            for (int st = 0; st < NSTEP; ++st) {
                float2 Hn[2], Pn[2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    Hn[s] = shfl_up2(H[s], 1 << st, LPR);
                    if (st + 1 < NSTEP) Pn[s] = shfl_up2(P[s], 1 << st, LPR);
                }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (ge[st]) {
                        H[s] = ffma2(P[s], Hn[s], H[s]);
                        if (st + 1 < NSTEP) P[s] = fmul2(P[s], Pn[s]);
                    }
                }
            }
            if (!AGG) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    float2 hs = shfl_up2(H[s], 1, LPR);
                    if (j == 0) hs = s ? make_float2(car4.z, car4.w) : make_float2(car4.x, car4.y);
#pragma unroll
                    for (int i = 0; i < T; ++i) ya[i] = ffma2(cp[s][i], hs, ya[i]);
                }
            }
            // lanes 7, 15, ... hold the state after every 64th token; the last one is the state entering the next chunk
            if ((j & 7) == 7) *reinterpret_cast<float4 *>(s_ck + (rpg * NCK + (j >> 3)) * 16 + n0) = make_float4(H[0].x, H[0].y, H[1].x, H[1].y);
        }

        if (!AGG) {
            __syncthreads();            // everybody is done with the B/C tile: fetch the next one under the epilogue
            if (c + 1 < c_end) issue_tile(c + 1);
            // ---- epilogue: gate and store ---------------------------------------------------------------------------------------
            float zz[2][T];
            if constexpr (ORD) {
                if (has_z) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        uint4 q[NQ];
#pragma unroll
                        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(s_elem_t + ((2 * 2 + r) * NQ + k) * NT * 16);
                        float e[8];
                        Raw8<IN_T>::unpack(q, e);
                        ord_to_tokens(p.ord, tl, e, zz[r]);
                    }
                }
            } else if (has_z) {
                load_elem(2, 0, zz[0]);
                load_elem(2, 1, zz[1]);
            }
            if (c + 1 < c_end) issue_z(tl + CH < L);
            cp_async_commit();
            if (ok) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (row_ok[r]) {
                        float yv[T];
#pragma unroll
                        for (int i = 0; i < T; ++i) yv[i] = r ? ya[i].y : ya[i].x;
                        if (y_p[r] != nullptr) store8<IN_T, REV>(y_p[r], yv);
                        if (has_z) {
#pragma unroll
                            for (int i = 0; i < T; ++i) yv[i] *= zz[r][i] * sigmoid3(zz[r][i]);
                        }
                        if constexpr (ORD) {
                            ord_store8<IN_T>(p.ord, tl, o_row[r], yv);
                        } else {
                            store8<IN_T, REV>(o_p[r], yv);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                o_p[r] += STEP;
                if (y_p[r] != nullptr) y_p[r] += STEP;
            }
            // ---- saved states x[b][row][k][n] = h after token 64(k+1)-1 (the __syncthreads above ordered the s_ck writes) -------
            if (p.x != nullptr) {
                const int n = lane & 15, r = lane >> 4;
#pragma unroll
                for (int g = 0; g < RPW; ++g) {
                    const int row = row0 + 2 * (warp * RPW + g) + r;
                    float *xp = p.x + (((int64_t)b * D + row) * p.nx + c * NCK) * N + n;
#pragma unroll
                    for (int ck = 0; ck < NCK; ++ck) {
                        const float2 v = s_ck[((warp * RPW + g) * NCK + ck) * 16 + n];
                        if (row < D && n < N && c * NCK + ck < p.nx) xp[ck * N] = r ? v.y : v.x;
                    }
                }
            }
        } else {
            __syncthreads();
            if (c + 1 < c_end) issue_tile(c + 1);
            cp_async_commit();
        }
    }

    // ---- segment / sequence end ---------------------------------------------------------------------------------------------
    __syncthreads();
    if (AGG) {
        for (int i = tid; i < NRP * 16; i += NT) {
            const int g = i >> 4, n = i & 15;
            const float2 h = s_ck[(g * NCK + NCK - 1) * 16 + n];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + 2 * g + r;
                if (row < D) p.seg_hend[(((int64_t)b * D + row) * p.nseg + seg) * 16 + n] = r ? h.y : h.x;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float s = dsum[r];
#pragma unroll
            for (int k = 1; k < LPR; k <<= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
            if (j == 0 && row_ok[r]) p.seg_dsum[((int64_t)b * D + rowA + r) * p.nseg + seg] = s;
        }
    } else if (p.last_state != nullptr && seg == p.nseg - 1) {
        for (int i = tid; i < NRP * 16; i += NT) {
            const int g = i >> 4, n = i & 15;
            const float2 h = s_ck[(g * NCK + NCK - 1) * 16 + n];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + 2 * g + r;
                if (row < D && n < N) p.last_state[((int64_t)b * D + row) * N + n] = r ? h.y : h.x;
            }
        }
    }
}

}  // namespace mmu

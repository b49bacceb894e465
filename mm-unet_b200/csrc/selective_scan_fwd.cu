// Selective scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (requirements/Mamba/mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303)
// with a different decomposition (see DESIGN.md "scan forward"):
//
//   * a warp owns 4 channel rows x 8 token-lanes; a lane owns T consecutive tokens of RD rows, so the cross-lane
//     prefix combination is a 3-step shuffle scan (8 lanes) instead of a 128-thread CUB BlockScan;
//   * the dstate axis is split over NGW warps of the CTA and walked two states at a time with packed
//     FFMA2/FMUL2 (Blackwell issues two fp32 lanes per slot); partial y's meet in shared memory;
//   * B/C chunks are staged once per CTA and shared by all its rows (the reference re-reads them per channel);
//   * the recurrence is applied in "correction form": y_t += C_t*(hloc_t + Pcum_t*h_start), which needs one
//     MUFU.EX2 per (token,state) and no second serial pass;
//   * the sequence can be split over CTAs (grid.z segments): an aggregate pass produces per-segment
//     (sum(delta), h_end), a tiny kernel chains them, the main pass starts each segment from its true state.
//     This is what fills the GPU for MM-UNet's D=6 / L=65536 scans (hard part 1).
//
// The state h is fp32 in registers; chunk-to-chunk carries live in shared memory (dstate is a runtime value).
#include <algorithm>
#include <cstdlib>

#include "scan3_fwd.cuh"
#include "tma_map.cuh"
#include "scan4.cuh"
#include "scan5_fwd.cuh"
#include "scan_tiles.cuh"

namespace mmu {

struct FwdArgs {
    const void *u, *delta, *z, *Bm, *Cm;
    const float *A, *Dv, *dbias;
    void *out, *ysave;
    float *x, *last_state;
    float *seg_hend, *seg_dsum;
    const float *hin;
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, y_bs, y_ds, B_bs, B_ns, C_bs, C_ns;
    int B, D, L, N, Ne;
    int nseg, cps, nchunks, nx;
    int softplus, reverse;
    unsigned vec_mask;   // bit0 u, 1 delta, 2 z, 3 out, 4 B, 5 C, 6 ysave
};

template <int RD, int RQ, int NGW> struct FwdCfg {
    static constexpr int T = 8, LS = 8, RG = 4;
    static constexpr int WROWS = RG * RD, R = RQ * WROWS, TL = LS * T, NT = 32 * RQ * NGW, CPT = T / 4;
    static_assert(TL == MMU_STATE_STRIDE, "one saved state per chunk");
    static size_t smem_bytes(int Ne) {
        return sizeof(float) * ((size_t)3 * R * TL + 2 * (size_t)Ne * TL + (size_t)NGW * R * TL + 2 * (size_t)R * Ne +
                                2 * R);
    }
};

template <typename IN_T, int RD, int RQ, int NGW, bool AGG>
__global__ void __launch_bounds__(32 * RQ * NGW, 512 / (32 * RQ * NGW)) scan_fwd_kernel(const __grid_constant__ FwdArgs p) {
    using Cfg = FwdCfg<RD, RQ, NGW>;
    constexpr int T = Cfg::T, R = Cfg::R, TL = Cfg::TL, NT = Cfg::NT, CPT = Cfg::CPT, WROWS = Cfg::WROWS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wq = warp / NGW, g = warp % NGW, rg = lane >> 3, j = lane & 7;
    const int b = blockIdx.y, row0 = blockIdx.x * R, seg = blockIdx.z;
    const int N = p.N, Ne = p.Ne, NP = Ne >> 1, D = p.D, L = p.L;
    const bool has_z = p.z != nullptr;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_u = reinterpret_cast<float *>(smem_raw);   // [R][TL]
    float *s_dl = s_u + R * TL;                          // [R][TL]  softplus(delta + bias)
    float *s_z = s_dl + R * TL;                          // [R][TL]
    float *s_B = s_z + R * TL;                           // [NP][TL][2]  pair-interleaved
    float *s_C = s_B + Ne * TL;                          // [NP][TL][2]
    float *s_yp = s_C + Ne * TL;                         // [NGW][R][TL] partial y per dstate group
    float *s_A2 = s_yp + NGW * R * TL;                   // [R][Ne]  A * log2(e)
    float *s_carry = s_A2 + R * Ne;                      // [R][Ne]  state entering the next chunk
    float *s_bias = s_carry + R * Ne;                    // [R]
    float *s_D = s_bias + R;                             // [R]

    for (int i = tid; i < R * Ne; i += NT) {
        const int r = i / Ne, n = i - r * Ne, row = row0 + r;
        const bool ok = row < D && n < N;
        s_A2[i] = ok ? p.A[(int64_t)row * N + n] * kLog2e : 0.f;
        float h0 = 0.f;
        if (!AGG && ok && p.hin != nullptr && seg > 0) h0 = p.hin[(((int64_t)b * D + row) * p.nseg + seg) * Ne + n];
        s_carry[i] = h0;
    }
    for (int r = tid; r < R; r += NT) {
        const int row = row0 + r;
        s_bias[r] = (row < D && p.dbias != nullptr) ? p.dbias[row] : 0.f;
        s_D[r] = (row < D && p.Dv != nullptr) ? p.Dv[row] : 0.f;
    }

    const IN_T *u_b = reinterpret_cast<const IN_T *>(p.u) + (int64_t)b * p.u_bs;
    const IN_T *dl_b = reinterpret_cast<const IN_T *>(p.delta) + (int64_t)b * p.dl_bs;
    const IN_T *z_b = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs : nullptr;
    const IN_T *B_b = reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs;
    const IN_T *C_b = reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs;
    IN_T *o_b = AGG ? nullptr : reinterpret_cast<IN_T *>(p.out) + (int64_t)b * p.o_bs;
    IN_T *y_b = (AGG || p.ysave == nullptr) ? nullptr : reinterpret_cast<IN_T *>(p.ysave) + (int64_t)b * p.y_bs;
    const bool rev = p.reverse != 0, sp = p.softplus != 0;

    const int npw = (NP + NGW - 1) / NGW;
    const int pair0 = g * npw, pair1 = min(NP, pair0 + npw);

    // per-thread constant shared-memory offsets (hoisted out of every loop)
    int off_t[RD][CPT];            // my T tokens inside a row tile
#pragma unroll
    for (int r = 0; r < RD; ++r)
#pragma unroll
        for (int cc = 0; cc < CPT; ++cc) off_t[r][cc] = (wq * WROWS + rg * RD + r) * TL + 4 * swz_chunk<T>(j * CPT + cc);
    int off_bc[T / 2];             // my T tokens inside a pair-row of a B / C tile
#pragma unroll
    for (int cc = 0; cc < T / 2; ++cc) off_bc[cc] = bc_off(j * (T / 2) + cc);
    const int lr0 = wq * WROWS + rg * RD;

    float dsum[RD];
#pragma unroll
    for (int r = 0; r < RD; ++r) dsum[r] = 0.f;

    // ---- chunk loop with register prefetch: the global loads of chunk c+1 are in flight while chunk c is scanned ----
    constexpr int QT = R * (TL / 4);                      // quads of a row tile
    constexpr int KR = (QT + NT - 1) / NT;
    constexpr int K2 = (8 * (TL / 4) + NT - 1) / NT;      // pair-quads per thread of a B / C tile (fast path: dstate <= 16)
    const bool fast_all = (p.vec_mask & 0x80u) != 0;      // every tensor vector-accessible and dstate <= 16 (host-checked)
    Quad<IN_T> q_u[KR], q_dl[KR], q_z[KR], q_B[2 * K2], q_C[2 * K2];
    auto ident = [](int, float v) { return v; };
    auto dl_xf = [&](int r, float v) {
        const float xx = v + s_bias[r];
        return sp ? softplus_f(xx) : xx;
    };
    auto prefetch = [&](int c) {
        const int t0 = c * TL;
        tile_prefetch<IN_T, TL, NT, KR>(q_u, u_b, p.u_ds, row0, D, QT, t0, L, rev, tid);
        tile_prefetch<IN_T, TL, NT, KR>(q_dl, dl_b, p.dl_ds, row0, D, QT, t0, L, rev, tid);
        bc_prefetch<IN_T, TL, NT, K2>(q_B, B_b, p.B_ns, N, NP, t0, L, rev, tid);
        if (!AGG) {
            if (has_z) tile_prefetch<IN_T, TL, NT, KR>(q_z, z_b, p.z_ds, row0, D, QT, t0, L, rev, tid);
            bc_prefetch<IN_T, TL, NT, K2>(q_C, C_b, p.C_ns, N, NP, t0, L, rev, tid);
        }
    };
    const int c_begin = seg * p.cps, c_end = min(p.nchunks, c_begin + p.cps);
    bool cur_fast = fast_all && (c_begin + 1) * TL <= L;
    if (cur_fast) prefetch(c_begin);
    for (int c = c_begin; c < c_end; ++c) {
        const int t0 = c * TL;
        __syncthreads();   // tiles free (previous epilogue finished), tables initialised
        if (cur_fast) {
            tile_commit<IN_T, T, TL, NT, KR>(s_u, q_u, QT, rev, tid, ident);
            tile_commit<IN_T, T, TL, NT, KR>(s_dl, q_dl, QT, rev, tid, dl_xf);
            bc_commit<IN_T, TL, NT, K2>(s_B, q_B, NP, rev, tid);
            if (!AGG) {
                if (has_z) tile_commit<IN_T, T, TL, NT, KR>(s_z, q_z, QT, rev, tid, ident);
                bc_commit<IN_T, TL, NT, K2>(s_C, q_C, NP, rev, tid);
            }
        } else {
            load_tile<IN_T, T, TL, NT>(s_u, u_b, p.u_ds, row0, R, D, t0, L, rev, p.vec_mask & 1u, tid, 0.f, ident);
            load_tile<IN_T, T, TL, NT>(s_dl, dl_b, p.dl_ds, row0, R, D, t0, L, rev, p.vec_mask & 2u, tid, 0.f, dl_xf);
            bc_load_generic<IN_T, TL, NT>(s_B, B_b, p.B_ns, N, Ne, t0, L, rev, tid);
            if (!AGG) {
                if (has_z) load_tile<IN_T, T, TL, NT>(s_z, z_b, p.z_ds, row0, R, D, t0, L, rev, p.vec_mask & 4u, tid, 0.f, ident);
                bc_load_generic<IN_T, TL, NT>(s_C, C_b, p.C_ns, N, Ne, t0, L, rev, tid);
            }
        }
        __syncthreads();
        cur_fast = fast_all && (c + 1) < c_end && (c + 2) * TL <= L;
        if (cur_fast) prefetch(c + 1);

        // ---- per-(token,row) registers: delta and delta*u of my T tokens -------------------------------
        float dl[RD][T], dlu[RD][T];
#pragma unroll
        for (int r = 0; r < RD; ++r) {
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
                const float4 d4 = *reinterpret_cast<const float4 *>(s_dl + off_t[r][cc]);
                const float4 u4 = *reinterpret_cast<const float4 *>(s_u + off_t[r][cc]);
                dl[r][4 * cc + 0] = d4.x, dl[r][4 * cc + 1] = d4.y, dl[r][4 * cc + 2] = d4.z, dl[r][4 * cc + 3] = d4.w;
                dlu[r][4 * cc + 0] = d4.x * u4.x, dlu[r][4 * cc + 1] = d4.y * u4.y;
                dlu[r][4 * cc + 2] = d4.z * u4.z, dlu[r][4 * cc + 3] = d4.w * u4.w;
            }
            if (AGG) {
#pragma unroll
                for (int i = 0; i < T; ++i) dsum[r] += dl[r][i];
            }
        }
        float2 y2[RD][T];
#pragma unroll
        for (int r = 0; r < RD; ++r)
#pragma unroll
            for (int i = 0; i < T; ++i) y2[r][i] = make_float2(0.f, 0.f);

        const float *pB = s_B + pair0 * 2 * TL, *pC = s_C + pair0 * 2 * TL;
        const float *pA2 = s_A2 + lr0 * Ne + 2 * pair0;
        float *pcar = s_carry + lr0 * Ne + 2 * pair0;
        for (int pr = pair0; pr < pair1; ++pr, pB += 2 * TL, pC += 2 * TL, pA2 += 2, pcar += 2) {
            float2 A2[RD], P[RD], hl[RD];
            float2 cp[RD][T];
#pragma unroll
            for (int r = 0; r < RD; ++r) {
                A2[r] = *reinterpret_cast<const float2 *>(pA2 + r * Ne);
                P[r] = make_float2(1.f, 1.f);
                hl[r] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int cc = 0; cc < T / 2; ++cc) {
                const float4 b4 = *reinterpret_cast<const float4 *>(pB + off_bc[cc]);    // (B_2p, B_2p+1) of two tokens
                float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!AGG) c4 = *reinterpret_cast<const float4 *>(pC + off_bc[cc]);
                const float2 Bv[2] = {make_float2(b4.x, b4.y), make_float2(b4.z, b4.w)};
                const float2 Cv[2] = {make_float2(c4.x, c4.y), make_float2(c4.z, c4.w)};
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int i = 2 * cc + k;
#pragma unroll
                    for (int r = 0; r < RD; ++r) {
                        const float2 a = ex2(fmul2(splat(dl[r][i]), A2[r]));
                        const float2 bb = fmul2(splat(dlu[r][i]), Bv[k]);
                        hl[r] = ffma2(a, hl[r], bb);
                        P[r] = fmul2(P[r], a);
                        if (!AGG) {
                            y2[r][i] = ffma2(Cv[k], hl[r], y2[r][i]);
                            cp[r][i] = fmul2(Cv[k], P[r]);
                        }
                    }
                }
            }
            // ---- 8-lane inclusive scan of (P, hl), then apply the chunk carry -------------------------------
#pragma unroll
            for (int r = 0; r < RD; ++r) {
                float2 Pi = P[r], Hi = hl[r];
#pragma unroll
                for (int k = 1; k < 8; k <<= 1) {
                    const float2 Pn = shfl_up2(Pi, k, 8), Hn = shfl_up2(Hi, k, 8);
                    if (j >= k) {
                        Hi = ffma2(Pi, Hn, Hi);
                        Pi = fmul2(Pi, Pn);
                    }
                }
                const float2 hc = *reinterpret_cast<const float2 *>(pcar + r * Ne);
                const float2 send = ffma2(Pi, hc, Hi);   // state after my last token
                float2 hs = shfl_up2(send, 1, 8);        // state before my first token
                if (j == 0) hs = hc;
                __syncwarp();
                if (j == 7) *reinterpret_cast<float2 *>(pcar + r * Ne) = send;
                if (!AGG) {
#pragma unroll
                    for (int i = 0; i < T; ++i) y2[r][i] = ffma2(cp[r][i], hs, y2[r][i]);
                }
            }
        }
        if (AGG) continue;

        // ---- partial y of my dstate group -> shared -----------------------------------------------------------
#pragma unroll
        for (int r = 0; r < RD; ++r)
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc)
                *reinterpret_cast<float4 *>(s_yp + g * R * TL + off_t[r][cc]) =
                    make_float4(y2[r][4 * cc].x + y2[r][4 * cc].y, y2[r][4 * cc + 1].x + y2[r][4 * cc + 1].y,
                                y2[r][4 * cc + 2].x + y2[r][4 * cc + 2].y, y2[r][4 * cc + 3].x + y2[r][4 * cc + 3].y);
        __syncthreads();
        // ---- epilogue: sum groups, D skip, SiLU gate (selective_scan_fwd_kernel.cuh:158,280-298), store to global ----
        for (int idx = tid; idx < QT; idx += NT) {
            const int r = idx / (TL / 4), cs = idx - r * (TL / 4);
            const int off = idx * 4;   // raw (swizzled) position: all row tiles share the layout
            float4 y = *reinterpret_cast<const float4 *>(s_yp + off);
#pragma unroll
            for (int gg = 1; gg < NGW; ++gg) {
                const float4 t = *reinterpret_cast<const float4 *>(s_yp + gg * R * TL + off);
                y.x += t.x, y.y += t.y, y.z += t.z, y.w += t.w;
            }
            const float4 u4 = *reinterpret_cast<const float4 *>(s_u + off);
            const float dsk = s_D[r];
            y.x = fmaf(dsk, u4.x, y.x), y.y = fmaf(dsk, u4.y, y.y), y.z = fmaf(dsk, u4.z, y.z), y.w = fmaf(dsk, u4.w, y.w);
            if (y_b != nullptr) {   // pre-gate y for the backward (the reference's saved `out`)
                const int row_ = row0 + r, t_ = t0 + 4 * swz_chunk<T>(cs);
                if (row_ < D && t_ < L) {
                    const float v_[4] = {y.x, y.y, y.z, y.w};
                    store_quad<IN_T>(y_b + (int64_t)row_ * p.y_ds, t_, L, rev, p.vec_mask & 64u, v_);
                }
            }
            if (has_z) {
                const float4 z4 = *reinterpret_cast<const float4 *>(s_z + off);
                y.x *= z4.x * sigmoid_f(z4.x), y.y *= z4.y * sigmoid_f(z4.y);
                y.z *= z4.z * sigmoid_f(z4.z), y.w *= z4.w * sigmoid_f(z4.w);
            }
            const int row = row0 + r, t = t0 + 4 * swz_chunk<T>(cs);   // the swizzle is an involution
            if (row < D && t < L) {
                const float v[4] = {y.x, y.y, y.z, y.w};
                store_quad<IN_T>(o_b + (int64_t)row * p.o_ds, t, L, rev, p.vec_mask & 8u, v);
            }
        }
        // ---- the state after this chunk = x[b][row][c][:]  (chunk length == MMU_STATE_STRIDE) --------------------------
        if (p.x != nullptr && c < p.nx) {
            for (int i = tid; i < R * Ne; i += NT) {
                const int r = i / Ne, n = i - r * Ne, row = row0 + r;
                if (row < D && n < N) p.x[(((int64_t)b * D + row) * p.nx + c) * N + n] = s_carry[i];
            }
        }
    }

    __syncthreads();
    if (AGG) {
        for (int i = tid; i < R * Ne; i += NT) {
            const int r = i / Ne, n = i - r * Ne, row = row0 + r;
            if (row < D) p.seg_hend[(((int64_t)b * D + row) * p.nseg + seg) * Ne + n] = s_carry[i];
        }
        if (g == 0) {
#pragma unroll
            for (int r = 0; r < RD; ++r) {
                float s = dsum[r];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                const int row = row0 + lr0 + r;
                if (j == 0 && row < D) p.seg_dsum[((int64_t)b * D + row) * p.nseg + seg] = s;
            }
        }
    } else if (p.last_state != nullptr && seg == p.nseg - 1) {
        for (int i = tid; i < R * Ne; i += NT) {
            const int r = i / Ne, n = i - r * Ne, row = row0 + r;
            if (row < D && n < N) p.last_state[((int64_t)b * D + row) * N + n] = s_carry[i];
        }
    }
}

// chain the per-segment aggregates: hin[s] = state entering segment s
__global__ void scan_fwd_chain_kernel(const float *__restrict__ A, const float *__restrict__ seg_hend,
                                      const float *__restrict__ seg_dsum, float *__restrict__ hin, int B, int D, int N,
                                      int Ne, int nseg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * D * Ne) return;
    const int n = (int)(i % Ne);
    const int64_t bd = i / Ne;
    const int row = (int)(bd % D);
    const float a2 = n < N ? A[(int64_t)row * N + n] * kLog2e : 0.f;
    // the loads do not depend on the carry: fetch 8 segments at a time so that the serial part is 8 fmas per memory latency
    float h = 0.f;
    for (int s0 = 0; s0 < nseg; s0 += 8) {
        float pa[8], he[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int s = min(s0 + k, nseg - 1);
            pa[k] = seg_dsum[bd * nseg + s], he[k] = seg_hend[(bd * nseg + s) * Ne + n];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int s = s0 + k;
            if (s < nseg) {
                hin[(bd * nseg + s) * Ne + n] = h;
                if (s + 1 < nseg) h = fmaf(ex2(a2 * pa[k]), h, he[k]);
            }
        }
    }
}


// ---- host side ----------------------------------------------------------------------------------------
namespace {

struct FwdPlan {
    int cfg;    // index into the instantiation table
    int R, TL;  // rows per CTA, tokens per chunk
    int nseg, cps, nchunks;
};

// instantiation table: {RD, RQ, NGW}
constexpr int kNumFwdCfg = 5;
constexpr int kFwdCfg[kNumFwdCfg][3] = {
    {1, 2, 2},   // 0:  8 rows, 4 warps
    {1, 1, 4},   // 1:  4 rows, 4 warps (narrow D)
    {1, 4, 2},   // 2: 16 rows, 8 warps
    {1, 2, 4},   // 3:  8 rows, 8 warps
    {2, 2, 2},   // 4: 16 rows, 4 warps, two rows per lane
};

int env_int(const char *name, int dflt) { return knob(name, dflt); }

FwdPlan plan_fwd(int B, int D, int L, int /*N*/) {
    FwdPlan pl;
    pl.cfg = D <= 4 ? 1 : 0;
    pl.cfg = env_int("MMU_FWD_CFG", pl.cfg);
    if (pl.cfg < 0 || pl.cfg >= kNumFwdCfg) pl.cfg = 0;
    const int *c = kFwdCfg[pl.cfg];
    pl.R = c[1] * 4 * c[0];
    pl.TL = 64;
    pl.nchunks = (L + pl.TL - 1) / pl.TL;
    const int warps = B * ((D + pl.R - 1) / pl.R) * c[1] * c[2];
    const int target = 148 * 16;
    int nseg = warps >= 148 * 6 ? 1 : (target + warps - 1) / warps;   // splitting costs a second pass: only when starved
    nseg = std::min(nseg, std::max(1, pl.nchunks / 4));   // at least 4 chunks per segment
    nseg = std::max(1, std::min(nseg, 64));
    nseg = env_int("MMU_FWD_NSEG", nseg);
    nseg = std::max(1, std::min(nseg, pl.nchunks));
    pl.cps = (pl.nchunks + nseg - 1) / nseg;
    pl.nseg = (pl.nchunks + pl.cps - 1) / pl.cps;
    return pl;
}

template <typename IN_T, int RD, int RQ, int NGW>
int launch_fwd(const FwdArgs &a, bool agg, cudaStream_t st) {
    using Cfg = FwdCfg<RD, RQ, NGW>;
    const size_t smem = Cfg::smem_bytes(a.Ne);
    if (smem > 227 * 1024) return set_error(MMU_ERR_UNSUPPORTED, "selective_scan_fwd: dstate %d needs %zu B smem", a.N, smem);
    dim3 grid((a.D + Cfg::R - 1) / Cfg::R, a.B, a.nseg), block(Cfg::NT);
    if (agg) {
        auto k = scan_fwd_kernel<IN_T, RD, RQ, NGW, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, block, smem, st>>>(a);
    } else {
        auto k = scan_fwd_kernel<IN_T, RD, RQ, NGW, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, block, smem, st>>>(a);
    }
    count_launch();
    return check_launch("selective_scan_fwd");
}

template <typename IN_T> int dispatch_fwd(int cfg, const FwdArgs &a, bool agg, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_fwd<IN_T, 1, 2, 2>(a, agg, st);
        case 1: return launch_fwd<IN_T, 1, 1, 4>(a, agg, st);
        case 2: return launch_fwd<IN_T, 1, 4, 2>(a, agg, st);
        case 3: return launch_fwd<IN_T, 1, 2, 4>(a, agg, st);
        default: return launch_fwd<IN_T, 2, 2, 2>(a, agg, st);
    }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- v3 host side (scan3_fwd.cuh) ------------------------------------------------------------------------------------------
template <typename IN_T> bool row_aligned16(const void *p, int64_t s0, int64_t s1) {
    return reinterpret_cast<uintptr_t>(p) % 16 == 0 && (s0 * (int64_t)sizeof(IN_T)) % 16 == 0 && (s1 * (int64_t)sizeof(IN_T)) % 16 == 0;
}

template <typename IN_T> bool fwd3_eligible(const mmu_scan_fwd_params *p) {
    if (env_int("MMU_SCAN_V", 3) < 3) return false;
    if (p->dstate > 16 || p->seqlen % 8 != 0) return false;
    if (!row_aligned16<IN_T>(p->u, p->u_bs, p->u_ds) || !row_aligned16<IN_T>(p->delta, p->delta_bs, p->delta_ds) ||
        !row_aligned16<IN_T>(p->out, p->out_bs, p->out_ds) || !row_aligned16<IN_T>(p->B, p->B_bs, p->B_ns) ||
        !row_aligned16<IN_T>(p->C, p->C_bs, p->C_ns))
        return false;
    if (p->z && !row_aligned16<IN_T>(p->z, p->z_bs, p->z_ds)) return false;
    if (p->y && !row_aligned16<IN_T>(p->y, p->y_bs, p->y_ds)) return false;
    return true;
}

struct Fwd3Plan {
    int LPR, W, nseg, cps, nchunks;
};

Fwd3Plan plan_fwd3(int B, int D, int L, bool ordered = false) {
    Fwd3Plan pl;
    pl.LPR = (!ordered && env_int("MMU_FWD3_LPR", 32) == 16) ? 16 : 32;
    if (pl.LPR == 16) {
        pl.W = 4;
    } else {
        pl.W = env_int("MMU_FWD3_W", D <= 2 ? 1 : (D <= 4 ? 2 : (D % 8 != 0 && D % 6 == 0 ? 3 : 4)));
        if (pl.W < 1 || pl.W > 4) pl.W = 4;
    }
    const int RW = 2 * (32 / pl.LPR), R = pl.W * RW, CH = 8 * pl.LPR;
    pl.nchunks = (L + CH - 1) / CH;
    const int warps = B * ((D + R - 1) / R) * pl.W;
    int nseg = warps >= 148 * 4 ? 1 : (148 * 8 + warps - 1) / warps;   // splitting costs a second (aggregate) pass
    nseg = std::min(nseg, std::max(1, pl.nchunks / 2));
    nseg = std::max(1, std::min(nseg, 64));
    nseg = env_int("MMU_FWD_NSEG", nseg);
    nseg = std::max(1, std::min(nseg, pl.nchunks));
    pl.cps = (pl.nchunks + nseg - 1) / nseg;
    pl.nseg = (pl.nchunks + pl.cps - 1) / pl.cps;
    return pl;
}

template <typename IN_T, int LPR, int W, bool REV, bool AGG, bool ORD = false> int launch_fwd3(const Fwd3Args &a, cudaStream_t st) {
    using Cfg = Fwd3Cfg<IN_T, LPR, W>;
    dim3 grid((a.D + Cfg::R - 1) / Cfg::R, a.B, a.nseg), block(Cfg::NT);
    auto k = scan3_fwd_kernel<IN_T, LPR, W, REV, AGG, ORD>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    k<<<grid, block, Cfg::smem_bytes, st>>>(a);
    count_launch();
    return check_launch("selective_scan_fwd(v3)");
}

template <typename IN_T, bool AGG> int dispatch_fwd3(const Fwd3Args &a, const Fwd3Plan &pl, bool rev, cudaStream_t st) {
    if constexpr (!AGG) {
        if (a.ord.kind != MMU_ORDER_ROWMAJOR) {     // fused scan order: gate / output permuted by the kernel (never with rev)
            if (pl.W == 1) return launch_fwd3<IN_T, 32, 1, false, false, true>(a, st);
            if (pl.W == 2) return launch_fwd3<IN_T, 32, 2, false, false, true>(a, st);
            if (pl.W == 3) return launch_fwd3<IN_T, 32, 3, false, false, true>(a, st);
            return launch_fwd3<IN_T, 32, 4, false, false, true>(a, st);
        }
    }
#define MMU_F3(LPR_, W_) (rev ? launch_fwd3<IN_T, LPR_, W_, true, AGG>(a, st) : launch_fwd3<IN_T, LPR_, W_, false, AGG>(a, st))
    if (pl.LPR == 16) return MMU_F3(16, 4);
    if (pl.W == 1) return MMU_F3(32, 1);
    if (pl.W == 2) return MMU_F3(32, 2);
    if (pl.W == 3) return MMU_F3(32, 3);
    return MMU_F3(32, 4);
#undef MMU_F3
}

template <typename IN_T> int run_fwd3(const mmu_scan_fwd_params *p, cudaStream_t st) {
    const Fwd3Plan pl = plan_fwd3(p->batch, p->dim, p->seqlen, p->order != MMU_ORDER_ROWMAJOR);
    Fwd3Args a{};
    make_ordmap(a.ord, p->order, p->order_h, p->order_w, p->order_ns, p->seqlen);
    a.u = p->u, a.delta = p->delta, a.z = p->z, a.Bm = p->B, a.Cm = p->C;
    a.A = p->A, a.Dv = p->D, a.dbias = p->delta_bias;
    a.out = p->out, a.ysave = p->z ? p->y : nullptr, a.x = p->x, a.last_state = p->last_state;
    a.u_bs = p->u_bs, a.u_ds = p->u_ds, a.dl_bs = p->delta_bs, a.dl_ds = p->delta_ds;
    a.z_bs = p->z_bs, a.z_ds = p->z_ds, a.o_bs = p->out_bs, a.o_ds = p->out_ds, a.y_bs = p->y_bs, a.y_ds = p->y_ds;
    a.B_bs = p->B_bs, a.B_ns = p->B_ns, a.C_bs = p->C_bs, a.C_ns = p->C_ns;
    a.B = p->batch, a.D = p->dim, a.L = p->seqlen, a.N = p->dstate;
    a.nseg = pl.nseg, a.cps = pl.cps, a.nchunks = pl.nchunks;
    a.nx = (p->seqlen + MMU_STATE_STRIDE - 1) / MMU_STATE_STRIDE;
    a.softplus = p->delta_softplus;
    const bool rev = p->reverse != 0;
#if MMU_TMA_TILE
    if constexpr (sizeof(IN_T) == 4) {
        if (int rc = encode_bc_map(a.tmB, p->B, p->seqlen, p->dstate, p->batch, p->B_ns, p->B_bs)) return rc;
        if (int rc = encode_bc_map(a.tmC, p->C, p->seqlen, p->dstate, p->batch, p->C_ns, p->C_bs)) return rc;
    }
#endif
    if (pl.nseg > 1) {
        const size_t n_state = (size_t)a.B * a.D * pl.nseg * 16;
        const size_t need = 2 * align256(n_state * 4) + align256((size_t)a.B * a.D * pl.nseg * 4);
        if (p->workspace == nullptr || p->workspace_bytes < need)
            return set_error(MMU_ERR_WORKSPACE, "selective_scan_fwd: workspace %zu < %zu", p->workspace_bytes, need);
        char *w = static_cast<char *>(p->workspace);
        a.seg_hend = reinterpret_cast<float *>(w);
        float *hin = reinterpret_cast<float *>(w + align256(n_state * 4));
        a.seg_dsum = reinterpret_cast<float *>(w + 2 * align256(n_state * 4));
        a.hin = nullptr;
        int rc = dispatch_fwd3<IN_T, true>(a, pl, rev, st);
        if (rc) return rc;
        const int64_t tot = (int64_t)a.B * a.D * 16;
        scan_fwd_chain_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(a.A, a.seg_hend, a.seg_dsum, hin, a.B, a.D,
                                                                             a.N, 16, pl.nseg);
        count_launch();
        rc = check_launch("scan_fwd_chain");
        if (rc) return rc;
        a.hin = hin;
    }
    return dispatch_fwd3<IN_T, false>(a, pl, rev, st);
}


// ---- v4 host side (scan4.cuh): wide problems, rows in lanes ------------------------------------------------------------------
struct Fwd4Plan {
    int nseg, sps, nstage, nrg;
};

Fwd4Plan plan_fwd4(int B, int D, int L) {
    Fwd4Plan pl;
    pl.nstage = L / 8;
    pl.nrg = (D + kS4Rows - 1) / kS4Rows;
    const int wps = B * pl.nrg;                                          // warps per segment
    int nseg = (148 * env_int("MMU_V4_WPSM", 12) + wps - 1) / wps;       // about one wave of resident warps
    nseg = std::min(nseg, std::max(1, pl.nstage / 4));                   // at least 32 tokens per segment
    nseg = env_int("MMU_FWD_NSEG", nseg);
    nseg = std::max(1, std::min(nseg, std::min(pl.nstage, kS4MaxSeg)));
    pl.sps = (pl.nstage + nseg - 1) / nseg;
    pl.nseg = (pl.nstage + pl.sps - 1) / pl.sps;
    return pl;
}

template <typename IN_T> bool fwd4_eligible(const mmu_scan_fwd_params *p) {
    if (env_int("MMU_SCAN_V", 3) < 4) return false;
    if (p->dim < env_int("MMU_V4_MIN_DIM", 64)) return false;
    const int xs = p->x_stride ? p->x_stride : MMU_STATE_STRIDE;
    if (xs != 8 && xs != 64) return false;
    if (p->x && p->dstate == 16 && reinterpret_cast<uintptr_t>(p->x) % 16 != 0) return false;
    return fwd3_eligible<IN_T>(p);
}

template <typename IN_T, bool REV, bool AGG> int launch_fwd4(const Fwd4Args &a, cudaStream_t st) {
    using Sm = S4Fwd<IN_T, AGG ? 2 : 3>;
    const size_t smem = (size_t)kS4W * Sm::kWarpBytes;
    auto k = scan4_fwd_kernel<IN_T, REV, AGG>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(a.nitems + kS4W - 1) / kS4W, 32 * kS4W, smem, st>>>(a);
    count_launch();
    return check_launch("selective_scan_fwd(v4)");
}

template <typename IN_T> int run_fwd4(const mmu_scan_fwd_params *p, cudaStream_t st) {
    const Fwd4Plan pl = plan_fwd4(p->batch, p->dim, p->seqlen);
    Fwd4Args a{};
    a.u = p->u, a.delta = p->delta, a.z = p->z, a.Bm = p->B, a.Cm = p->C;
    a.A = p->A, a.Dv = p->D, a.dbias = p->delta_bias;
    a.out = p->out, a.ysave = p->z ? p->y : nullptr, a.x = p->x, a.last_state = p->last_state;
    a.u_bs = p->u_bs, a.u_ds = p->u_ds, a.dl_bs = p->delta_bs, a.dl_ds = p->delta_ds;
    a.z_bs = p->z_bs, a.z_ds = p->z_ds, a.o_bs = p->out_bs, a.o_ds = p->out_ds, a.y_bs = p->y_bs, a.y_ds = p->y_ds;
    a.B_bs = p->B_bs, a.B_ns = p->B_ns, a.C_bs = p->C_bs, a.C_ns = p->C_ns;
    a.B = p->batch, a.D = p->dim, a.L = p->seqlen, a.N = p->dstate;
    a.nseg = pl.nseg, a.sps = pl.sps, a.nstage = pl.nstage, a.nrg = pl.nrg;
    const int xs = p->x_stride ? p->x_stride : MMU_STATE_STRIDE;
    a.xs8 = xs / 8, a.nx = (p->seqlen + xs - 1) / xs;
    a.softplus = p->delta_softplus;
    const bool rev = p->reverse != 0;
    if (pl.nseg > 1) {
        const size_t n_state = (size_t)a.B * a.D * pl.nseg * 16;
        const size_t need = 2 * align256(n_state * 4) + align256((size_t)a.B * a.D * pl.nseg * 4);
        if (p->workspace == nullptr || p->workspace_bytes < need)
            return set_error(MMU_ERR_WORKSPACE, "selective_scan_fwd: workspace %zu < %zu", p->workspace_bytes, need);
        char *w = static_cast<char *>(p->workspace);
        a.seg_hend = reinterpret_cast<float *>(w);
        float *hin = reinterpret_cast<float *>(w + align256(n_state * 4));
        a.seg_dsum = reinterpret_cast<float *>(w + 2 * align256(n_state * 4));
        a.hin = nullptr;
        a.nitems = a.B * a.nrg * (pl.nseg - 1);                          // the last segment's aggregate is never used
        int rc = rev ? launch_fwd4<IN_T, true, true>(a, st) : launch_fwd4<IN_T, false, true>(a, st);
        if (rc) return rc;
        const int64_t tot = (int64_t)a.B * a.D * 16;
        scan_fwd_chain_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(a.A, a.seg_hend, a.seg_dsum, hin, a.B, a.D,
                                                                             a.N, 16, pl.nseg);
        count_launch();
        rc = check_launch("scan_fwd_chain");
        if (rc) return rc;
        a.hin = hin;
    }
    a.nitems = a.B * a.nrg * pl.nseg;
    return rev ? launch_fwd4<IN_T, true, false>(a, st) : launch_fwd4<IN_T, false, false>(a, st);
}


// ---- v5 host side (scan5_fwd.cuh): wide problems (fp32 / bf16), lane rings ----------------------------------------------------------------
// Ring warps per CTA.  A CTA's time is fixed by the sequence length and its width (measured per 128-token round: ~3.2 us with 4 ring
// warps, ~4.4 us with 6, ~6.2 us with 2 warps at two CTAs per SM), so the plan minimises waves x time per round; the B/C tiles are per
// batch element, so a CTA never spans two.
int plan_fwd5(int B, int D) {
    const int forced = env_int("MMU_V5_W", 0);
    if (forced == 2 || forced == 4 || forced == 6) return forced;
    int best = 4;
    double best_cost = 1e30;
    for (int W : {4, 6, 2}) {
        const int ctas = B * ((D + 4 * W - 1) / (4 * W)), slots = W == 2 ? 296 : 148;
        const double cost = (double)((ctas + slots - 1) / slots) * (W == 4 ? 3.2 : (W == 6 ? 4.4 : 6.2));
        if (cost < best_cost - 1e-9) best_cost = cost, best = W;
    }
    return best;
}

template <typename IN_T> bool fwd5_eligible(const mmu_scan_fwd_params *p) {
    if (env_int("MMU_RING", 1) == 0) return false;
    // fused scan orders: two-row (two 16-byte pieces per lane and row) and, tiled through the landing slots, fp32 nslices (the
    // element-wise form - eight 4-byte pieces per lane and row - measured 187 us against v3's 179 us at config 2; plain order 158 us)
    if (p->order != MMU_ORDER_ROWMAJOR) {
        if (p->order == MMU_ORDER_FLIP || p->reverse || !ordmap_fusable(p->order, p->order_h, p->order_w, p->order_ns, p->seqlen)) return false;
        if (p->order == MMU_ORDER_NSLICES &&        // tiled through the landing slots: 4-byte elements, 8 / 16 / 32 slices, whole chunks
            (sizeof(IN_T) != 4 || (p->order_ns != 8 && p->order_ns != 16 && p->order_ns != 32) || p->seqlen % 128 != 0))
            return false;
    }
    // 2-byte I/O: the ring is latency bound and gains nothing from the halved bytes, v3 does (B16 D128 L65536 bf16: 1 809 -> 1 689 us,
    // config 2 bf16 no change), and its one 200 KB CTA per SM keeps the other directions' kernels of a v3 Mamba off the SM: opt-in
    if (sizeof(IN_T) != 4 && env_int("MMU_RING_BF16", 0) == 0) return false;
    const int xs = p->x_stride ? p->x_stride : MMU_STATE_STRIDE;
    if (xs != 64) return false;
    if (p->x && reinterpret_cast<uintptr_t>(p->x) % 16 != 0) return false;
    // rows / 4 ring warps walk the whole sequence: below ~1 300 rows v3 (which is latency bound too, but per row pair) is faster
    // (scripts/probe_v5_thresh.py: B8 D128 L4096 v3 105 us / ring 113 us, B4 D384 136 / 113 us); short sequences pay the three
    // fill / drain rounds
    if (p->seqlen < 512) return false;
    if ((int64_t)p->batch * ((p->dim + 3) / 4) < env_int("MMU_V5_MIN_WARPS", 320)) return false;
    return fwd3_eligible<IN_T>(p);
}

template <typename IN_T, int WR> int launch_fwd5(const Fwd3Args &a, bool rev, cudaStream_t st) {
    using Cfg = Fwd5Cfg<IN_T, WR>;
    dim3 grid((a.D + Cfg::R - 1) / Cfg::R, a.B), block(Cfg::NT);
    auto k = rev ? scan5_fwd_kernel<IN_T, WR, true> : scan5_fwd_kernel<IN_T, WR, false>;
    if (a.ord.kind == MMU_ORDER_TWOROW) k = scan5_fwd_kernel<IN_T, WR, false, 1>;
    if constexpr (sizeof(IN_T) == 4) {
        if (a.ord.kind == MMU_ORDER_NSLICES) k = scan5_fwd_kernel<IN_T, WR, false, 2>;
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    k<<<grid, block, Cfg::smem_bytes, st>>>(a);
    count_launch();
    return check_launch("selective_scan_fwd(v5)");
}

template <typename IN_T> int run_fwd5(const mmu_scan_fwd_params *p, cudaStream_t st) {
    Fwd3Args a{};
    make_ordmap(a.ord, p->order, p->order_h, p->order_w, p->order_ns, p->seqlen);
    a.u = p->u, a.delta = p->delta, a.z = p->z, a.Bm = p->B, a.Cm = p->C;
    a.A = p->A, a.Dv = p->D, a.dbias = p->delta_bias;
    a.out = p->out, a.ysave = p->z ? p->y : nullptr, a.x = p->x, a.last_state = p->last_state;
    a.u_bs = p->u_bs, a.u_ds = p->u_ds, a.dl_bs = p->delta_bs, a.dl_ds = p->delta_ds;
    a.z_bs = p->z_bs, a.z_ds = p->z_ds, a.o_bs = p->out_bs, a.o_ds = p->out_ds, a.y_bs = p->y_bs, a.y_ds = p->y_ds;
    a.B_bs = p->B_bs, a.B_ns = p->B_ns, a.C_bs = p->C_bs, a.C_ns = p->C_ns;
    a.B = p->batch, a.D = p->dim, a.L = p->seqlen, a.N = p->dstate;
    a.nseg = 1, a.cps = 0, a.nchunks = (p->seqlen + 127) / 128;
    a.nx = (p->seqlen + MMU_STATE_STRIDE - 1) / MMU_STATE_STRIDE;
    a.softplus = p->delta_softplus;
    const bool rev = p->reverse != 0;
    switch (plan_fwd5(p->batch, p->dim)) {
        case 6: return launch_fwd5<IN_T, 6>(a, rev, st);
        case 4: return launch_fwd5<IN_T, 4>(a, rev, st);
        default: return launch_fwd5<IN_T, 2>(a, rev, st);
    }
}

template <typename IN_T> struct HasV3 { static constexpr bool value = false; };
template <> struct HasV3<float> { static constexpr bool value = true; };
template <> struct HasV3<__nv_bfloat16> { static constexpr bool value = true; };

template <typename IN_T> int run_fwd(const mmu_scan_fwd_params *p, cudaStream_t st) {
    if (p->order != MMU_ORDER_ROWMAJOR) {        // fused scan order: only on the dstate <= 16 kernels, only for the fusable maps
        bool ok = p->order != MMU_ORDER_FLIP && !p->reverse && ordmap_fusable(p->order, p->order_h, p->order_w, p->order_ns, p->seqlen);
        if constexpr (HasV3<IN_T>::value) ok = ok && fwd3_eligible<IN_T>(p);
        else ok = false;
        if (!ok)
            return set_error(MMU_ERR_UNSUPPORTED, "selective_scan_fwd: scan order %d (H=%d W=%d nslices=%d) cannot be fused for this problem "
                             "(see mmu_scan_order_fusable); permute with mmu_scan_order_gather / _scatter", p->order, p->order_h, p->order_w, p->order_ns);
        if constexpr (HasV3<IN_T>::value) {
            if (fwd5_eligible<IN_T>(p)) return run_fwd5<IN_T>(p, st);
            return run_fwd3<IN_T>(p, st);
        }
    }
    if constexpr (HasV3<IN_T>::value) {
        if (fwd5_eligible<IN_T>(p)) return run_fwd5<IN_T>(p, st);
    }
    if constexpr (HasV3<IN_T>::value) {
        if (fwd4_eligible<IN_T>(p)) return run_fwd4<IN_T>(p, st);
        if (p->x_stride != 0 && p->x_stride != MMU_STATE_STRIDE)
            return set_error(MMU_ERR_UNSUPPORTED, "selective_scan_fwd: x_stride %d is only supported by the wide (v4) kernels", p->x_stride);
        if (fwd3_eligible<IN_T>(p)) return run_fwd3<IN_T>(p, st);
    }
    if (p->x_stride != 0 && p->x_stride != MMU_STATE_STRIDE)
        return set_error(MMU_ERR_UNSUPPORTED, "selective_scan_fwd: x_stride %d is only supported by the wide (v4) kernels", p->x_stride);
    const FwdPlan pl = plan_fwd(p->batch, p->dim, p->seqlen, p->dstate);
    FwdArgs a{};
    a.u = p->u, a.delta = p->delta, a.z = p->z, a.Bm = p->B, a.Cm = p->C;
    a.A = p->A, a.Dv = p->D, a.dbias = p->delta_bias;
    a.out = p->out, a.ysave = p->z ? p->y : nullptr, a.x = p->x, a.last_state = p->last_state;
    a.u_bs = p->u_bs, a.u_ds = p->u_ds, a.dl_bs = p->delta_bs, a.dl_ds = p->delta_ds;
    a.z_bs = p->z_bs, a.z_ds = p->z_ds, a.o_bs = p->out_bs, a.o_ds = p->out_ds, a.y_bs = p->y_bs, a.y_ds = p->y_ds;
    a.B_bs = p->B_bs, a.B_ns = p->B_ns, a.C_bs = p->C_bs, a.C_ns = p->C_ns;
    a.B = p->batch, a.D = p->dim, a.L = p->seqlen, a.N = p->dstate, a.Ne = (p->dstate + 1) & ~1;
    a.nseg = pl.nseg, a.cps = pl.cps, a.nchunks = pl.nchunks;
    a.nx = (p->seqlen + MMU_STATE_STRIDE - 1) / MMU_STATE_STRIDE;
    a.softplus = p->delta_softplus, a.reverse = p->reverse;
    const bool rev = p->reverse != 0;
    const int L = p->seqlen;
    a.vec_mask = (quad_ok<IN_T>(p->u, p->u_bs, p->u_ds, L, rev) ? 1u : 0u) |
                 (quad_ok<IN_T>(p->delta, p->delta_bs, p->delta_ds, L, rev) ? 2u : 0u) |
                 ((p->z && quad_ok<IN_T>(p->z, p->z_bs, p->z_ds, L, rev)) ? 4u : 0u) |
                 (quad_ok<IN_T>(p->out, p->out_bs, p->out_ds, L, rev) ? 8u : 0u) |
                 (quad_ok<IN_T>(p->B, p->B_bs, p->B_ns, L, rev) ? 16u : 0u) |
                 (quad_ok<IN_T>(p->C, p->C_bs, p->C_ns, L, rev) ? 32u : 0u) |
                 ((p->y && quad_ok<IN_T>(p->y, p->y_bs, p->y_ds, L, rev)) ? 64u : 0u);
    const unsigned need = 1u | 2u | 16u | 32u | (p->z ? 4u : 0u);
    if ((a.vec_mask & need) == need && a.Ne <= 16 && env_int("MMU_NO_PREFETCH", 0) == 0) a.vec_mask |= 0x80u;
    if (pl.nseg > 1) {
        const size_t n_state = (size_t)a.B * a.D * pl.nseg * a.Ne;
        const size_t need = 2 * align256(n_state * 4) + align256((size_t)a.B * a.D * pl.nseg * 4);
        if (p->workspace == nullptr || p->workspace_bytes < need)
            return set_error(MMU_ERR_WORKSPACE, "selective_scan_fwd: workspace %zu < %zu", p->workspace_bytes, need);
        char *w = static_cast<char *>(p->workspace);
        a.seg_hend = reinterpret_cast<float *>(w);
        float *hin = reinterpret_cast<float *>(w + align256(n_state * 4));
        a.seg_dsum = reinterpret_cast<float *>(w + 2 * align256(n_state * 4));
        a.hin = nullptr;
        int rc = dispatch_fwd<IN_T>(pl.cfg, a, true, st);
        if (rc) return rc;
        const int64_t tot = (int64_t)a.B * a.D * a.Ne;
        scan_fwd_chain_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(a.A, a.seg_hend, a.seg_dsum, hin, a.B, a.D,
                                                                             a.N, a.Ne, pl.nseg);
        count_launch();
        rc = check_launch("scan_fwd_chain");
        if (rc) return rc;
        a.hin = hin;
    }
    return dispatch_fwd<IN_T>(pl.cfg, a, false, st);
}

}  // namespace
}  // namespace mmu

extern "C" size_t mmu_selective_scan_fwd_workspace(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate) {
    // upper bound over every plan the dispatcher can pick (nseg <= 64 unless forced by MMU_FWD_NSEG)
    const int Ne = dstate <= 16 ? 16 : ((dstate + 1) & ~1);
    const int nchunks = (seqlen + 63) / 64;
    const size_t nseg = (size_t)std::max(1, std::min(nchunks, std::max(64, mmu::env_int("MMU_FWD_NSEG", 1))));
    const size_t n_state = (size_t)batch * dim * nseg * Ne;
    size_t need = 2 * mmu::align256(n_state * 4) + mmu::align256((size_t)batch * dim * nseg * 4);
    if (dstate <= 16 && seqlen % 8 == 0 && seqlen >= 8) {   // the wide (v4) plan
        const mmu::Fwd4Plan pl = mmu::plan_fwd4(batch, dim, seqlen);
        const size_t n4 = (size_t)batch * dim * pl.nseg * 16;
        need = std::max(need, 2 * mmu::align256(n4 * 4) + mmu::align256((size_t)batch * dim * pl.nseg * 4));
    }
    return need;
}

extern "C" int32_t mmu_scan_order_fusable(int32_t order, int32_t H, int32_t W, int32_t nslices, int32_t dstate, int32_t dtype) {
    using namespace mmu;
    if (order == MMU_ORDER_ROWMAJOR) return 1;
    if (order == MMU_ORDER_FLIP) return 0;        // the `reverse` flag
    if (env_int("MMU_FUSE", 1) == 0 || env_int("MMU_SCAN_V", 3) < 3) return 0;
    if (dstate > 16 || (dtype != MMU_F32 && dtype != MMU_BF16)) return 0;
    const int64_t L = (int64_t)H * W;
    if (H <= 0 || W <= 0 || L > INT32_MAX) return 0;
    // NSLICES: a 256-token chunk touches 256/ns consecutive elements per slice - fuse only when that is at least a 32-byte sector
    // of 4-byte elements (measured: ns = 64 at L = 65 536 loses 15 % to the tiled gather / scatter kernels; 2-byte elements would
    // need 2-byte async copies)
    if (order == MMU_ORDER_NSLICES && (dtype != MMU_F32 || nslices > 32)) return 0;
    return ordmap_fusable(order, H, W, nslices, (int)L) ? 1 : 0;
}

extern "C" int mmu_selective_scan_fwd(const mmu_scan_fwd_params *p, void *stream) {
    using namespace mmu;
    if (p == nullptr) return set_error(MMU_ERR_INVALID, "selective_scan_fwd: null params");
    if (p->batch <= 0 || p->dim <= 0 || p->seqlen <= 0 || p->dstate <= 0)
        return set_error(MMU_ERR_INVALID, "selective_scan_fwd: empty shape (batch=%d dim=%d seqlen=%d dstate=%d)", p->batch,
                         p->dim, p->seqlen, p->dstate);
    if (p->dstate > 256) return set_error(MMU_ERR_INVALID, "selective_scan only supports state dimension <= 256");
    if (!p->u || !p->delta || !p->A || !p->B || !p->C || !p->out)
        return set_error(MMU_ERR_INVALID, "selective_scan_fwd: null tensor pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p->dtype) {
        case MMU_F32: return run_fwd<float>(p, st);
        case MMU_BF16: return run_fwd<__nv_bfloat16>(p, st);
        case MMU_F16: return run_fwd<__half>(p, st);
        default: return set_error(MMU_ERR_UNSUPPORTED, "selective_scan_fwd: dtype %d", p->dtype);
    }
}

// Selective scan backward, v4 kernels (wide problems; see scan4.cuh for the decomposition).  Recompute based: the L x N state
// is never materialised.  Math: SURVEY.md Appendix A; replaces selective_scan_bwd_kernel (selective_scan_bwd_kernel.cuh:75-489).
//
//   scan4_bwd_agg_kernel   lane = 2 rows x 16 states; walks the tokens of a segment last -> first and computes the reverse
//                          aggregate E = e at the segment's first token from a zero inflow (e_t = a_t (C_t dy_t + e_{t+1})),
//                          plus sum(delta); scan3_bwd_chain_kernel chains them right to left.
//   scan4_bwd_kernel       lane = 2 rows x 8 tokens (one stage) in registers, walks the 16 states; stages last -> first.
//                          Forward recompute h_i from the state x saved after the previous stage (x stride 8), reverse sweep with
//                          the carry e[n] in registers, contributions:  dB, dC (row sums in the lane, 16-shuffle transpose over the
//                          warp's 64 rows, one red.global per (8 tokens, state)), S1 = sum_n dh B, S2 = sum_n dh a h_prev A
//                          (registers, per token), dA / dD / ddelta_bias (registers, one atomic per row and segment).
#pragma once
#include "scan4.cuh"

namespace mmu {

struct Bwd4Args {
    const void *u, *delta, *z, *dout, *ysave, *Bm, *Cm;
    const float *A, *Dv, *dbias, *x;
    void *du, *ddelta, *dz;
    float *dA, *dB, *dC, *dD, *ddbias;
    float *seg_E, *seg_dsum;            // aggregate pass outputs: [b*D + row][nseg][16], [b*D + row][nseg]
    const float *ein;                   // main pass: reverse carry entering each segment from the right (NULL when nseg == 1)
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, g_bs, g_ds, y_bs, y_ds, B_bs, B_ns, C_bs, C_ns;
    int64_t du_bs, du_ds, ddl_bs, ddl_ds, dz_bs, dz_ds, dB_bs, dB_ns, dC_bs, dC_ns;
    int B, D, L, N;
    int nseg, sps, nstage, nrg, nx;
    int softplus;
    int nitems;
};

// ---- reverse aggregates -----------------------------------------------------------------------------------------------------------
template <typename IN_T, bool REV>
__global__ void __launch_bounds__(32 * kS4W, 3) scan4_bwd_agg_kernel(const __grid_constant__ Bwd4Args p) {
    constexpr int T = 8, NTEN = 3, SB = kS4SB;      // delta | dout | z
    using Sm = S4Fwd<IN_T, NTEN>;
    constexpr int NQ = Sm::NQ, P = Sm::P, ES = (int)sizeof(IN_T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kS4W + warp;
    if (item >= p.nitems) return;
    const int rg = item % p.nrg, it1 = item / p.nrg, b = it1 % p.B, seg = 1 + it1 / p.B;   // segment 0 has nothing to its left
    const int D = p.D, L = p.L, N = p.N;
    const bool has_z = p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_slot = smem_raw + (size_t)warp * Sm::kWarpBytes;
    float *s_tile = reinterpret_cast<float *>(s_slot + Sm::kSlotBytes);      // [8 tokens][16]  C only
    const unsigned slot_u32 = smem_u32(s_slot);

    const int row0 = rg * kS4Rows, nrows = min(kS4Rows, D - row0);
    int row[2];
    bool row_ok[2];
    float2 A2[16], e[16];
    float bias[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = row0 + lane + 32 * r;
        row_ok[r] = rr < D;
        row[r] = min(rr, D - 1);
        bias[r] = p.dbias != nullptr ? p.dbias[row[r]] : 0.f;
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        A2[n] = make_float2(n < N ? p.A[(int64_t)row[0] * N + n] * kLog2e : 0.f, n < N ? p.A[(int64_t)row[1] * N + n] * kLog2e : 0.f);
        e[n] = make_float2(0.f, 0.f);
    }
    const int s_begin = seg * p.sps, s_end = min(p.nstage, s_begin + p.sps);
    const int nblk = (s_end - s_begin + SB - 1) / SB;
    auto moff = [&](int s, int nst) { return REV ? L - T * (s + nst) : T * s; };
    const char *g_in[NTEN];
    int64_t rs_in[NTEN];
    g_in[0] = reinterpret_cast<const char *>(p.delta) + ((int64_t)b * p.dl_bs + (int64_t)row0 * p.dl_ds) * ES, rs_in[0] = p.dl_ds * ES;
    g_in[1] = reinterpret_cast<const char *>(p.dout) + ((int64_t)b * p.g_bs + (int64_t)row0 * p.g_ds) * ES, rs_in[1] = p.g_ds * ES;
    g_in[2] = reinterpret_cast<const char *>(p.z) + ((int64_t)b * p.z_bs + (int64_t)row0 * p.z_ds) * ES, rs_in[2] = p.z_ds * ES;
    const int tn = lane & 15;
    const bool t_live = tn < N && lane < 16;
    const IN_T *t_src = reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)tn * p.C_ns;
    uint4 treg[NQ];
    auto tile_ldg = [&](int s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            treg[q] = t_live ? ldg16_pf(reinterpret_cast<const uint4 *>(t_src + moff(s, 1)) + q) : make_uint4(0u, 0u, 0u, 0u);
    };
    auto tile_sts = [&]() {
        float v[8];
        Raw8<IN_T>::unpack(treg, v);
        if (lane < 16) {
#pragma unroll
            for (int k = 0; k < 8; ++k) s_tile[(REV ? 7 - k : k) * 16 + lane] = v[k];
        }
    };
    // block k (counted from the right end of the segment) covers stages [s_lo, s_lo + nst)
    auto blk_range = [&](int k, int &s_lo, int &nst) {
        const int s_hi = s_end - k * SB;
        nst = min(SB, s_hi - s_begin);
        s_lo = s_hi - nst;
    };
    auto issue_block = [&](int k, int par) {
        int s_lo, nst;
        blk_range(k, s_lo, nst);
        const int64_t mo = (int64_t)moff(s_lo, nst) * ES;
#pragma unroll
        for (int t = 0; t < NTEN; ++t) {
            if (t == 2 && !has_z) continue;
            rb_load_async<P>(slot_u32 + (par * NTEN + t) * Sm::kBlkBytes, g_in[t] + mo, rs_in[t], nrows, nst * NQ, lane);
        }
    };
    auto load_stage = [&](int par, int t, int r, int ms, float (&v)[T]) {
        uint4 q[NQ];
        const unsigned char *base = s_slot + (par * NTEN + t) * Sm::kBlkBytes;
#pragma unroll
        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(base + rb_unit<P>(lane + 32 * r, ms * NQ + k) * 16);
        float ev[8];
        Raw8<IN_T>::unpack(q, ev);
        order8<REV>(ev, v);
    };

    issue_block(0, 0);
    cp_async_commit();
    tile_ldg(s_end - 1);
    tile_sts();
    float dsum[2] = {0.f, 0.f};

    for (int blk = 0; blk < nblk; ++blk) {
        const int par = blk & 1;
        int s_lo, nst;
        blk_range(blk, s_lo, nst);
        cp_async_wait_all();
        __syncwarp();
        if (blk + 1 < nblk) issue_block(blk + 1, par ^ 1);
        cp_async_commit();
#pragma unroll 1
        for (int i = nst - 1; i >= 0; --i) {
            const int s = s_lo + i, ms = REV ? nst - 1 - i : i;
            __syncwarp();               // tile of stage s is visible
            if (s > s_begin) tile_ldg(s - 1);
            float dd[2][T], gy[2][T];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                load_stage(par, 0, r, ms, dd[r]);
                load_stage(par, 1, r, ms, gy[r]);
                float zz[T];
                if (has_z) load_stage(par, 2, r, ms, zz);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const float xx = dd[r][k] + bias[r];
                    dd[r][k] = sp ? softplus3(xx) : xx;
                    dsum[r] += dd[r][k];
                    if (has_z) gy[r][k] *= zz[k] * sigmoid3(zz[k]);
                }
            }
#pragma unroll
            for (int k = T - 1; k >= 0; --k) {
                const float2 dl = make_float2(dd[0][k], dd[1][k]);
                const float2 dy = make_float2(gy[0][k], gy[1][k]);
                const float4 *tb = reinterpret_cast<const float4 *>(s_tile + k * 16);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 c4 = tb[g];
                    const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = 4 * g + j;
                        const float2 a = ex2(fmul2(dl, A2[n]));
                        e[n] = fmul2(a, ffma2(dy, splat(cv[j]), e[n]));
                    }
                }
            }
            __syncwarp();
            if (s > s_begin) tile_sts();
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!row_ok[r]) continue;
        const int64_t o = ((int64_t)b * D + row[r]) * p.nseg + seg;
        float4 *ep = reinterpret_cast<float4 *>(p.seg_E + o * 16);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            ep[g] = r ? make_float4(e[4 * g].y, e[4 * g + 1].y, e[4 * g + 2].y, e[4 * g + 3].y)
                      : make_float4(e[4 * g].x, e[4 * g + 1].x, e[4 * g + 2].x, e[4 * g + 3].x);
        p.seg_dsum[o] = dsum[r];
    }
}

// ---- main pass ----------------------------------------------------------------------------------------------------------------------
// shared memory of one warp.  Every block is one 8-token stage of the warp's 64 rows, moved cooperatively (rb_* in scan4.cuh).
template <typename IN_T> struct S4Bwd {
    static constexpr int NQ = Raw8<IN_T>::kQuads;                    // 16-byte pieces per row per stage (fp32 2: lane pairs fetch full sectors)
    static constexpr int kInBytes = 64 * NQ * 16;                    // one input tensor
    static constexpr int kSlotBytes = 5 * kInBytes;                  // u | delta | dout | z | y (single buffered: consumed by the prologue)
    static constexpr int kKeepBytes = 2 * 64 * 2 * 16;               // u, sigmoid(delta_raw + bias) as fp32 (P = 2)
    // du / ddelta staging: fp32 pieces coincide with the owner's keep pieces (same P = 2 layout: staged in place);
    // 2-byte types have one piece per row, which would alias OTHER lanes' keep pieces: own area
    static constexpr int kStageBytes = NQ == 2 ? 0 : 2 * 64 * 16;
    static constexpr int kSeedBytes = 2 * 64 * 4 * 16;               // [parity] x 64 rows x 16 states fp32 (P = 4)
    static constexpr int kTileBytes = 32 * 8 * 4;                    // [B states 0..15 | C states 0..15][8 tokens] fp32, logical order
    static constexpr int kTabBytes = 16 * 32 * 8;                    // A*log2e of (row 0, row 1): [state][lane] float2
    static constexpr int kWarpBytes = kSlotBytes + kKeepBytes + kSeedBytes + kTileBytes + kTabBytes + kStageBytes;
};

template <typename IN_T, bool REV>
__global__ void __launch_bounds__(32 * kS4W, 2) scan4_bwd_kernel(const __grid_constant__ Bwd4Args p) {
    constexpr int T = 8;
    using Sm = S4Bwd<IN_T>;
    constexpr int NQ = Sm::NQ, ES = (int)sizeof(IN_T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kS4W + warp;
    if (item >= p.nitems) return;
    const int rg = item % p.nrg, it1 = item / p.nrg, b = it1 % p.B, seg = it1 / p.B;
    const int D = p.D, L = p.L;                                      // dstate == 16 (checked by the host)
    const bool has_z = p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_slot = smem_raw + (size_t)warp * Sm::kWarpBytes;
    unsigned char *s_keep = s_slot + Sm::kSlotBytes;
    unsigned char *s_seed = s_keep + Sm::kKeepBytes;
    float *s_tile = reinterpret_cast<float *>(s_seed + Sm::kSeedBytes);
    float2 *s_A = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(s_tile) + Sm::kTileBytes);
    unsigned char *s_stage0 = NQ == 2 ? s_keep : reinterpret_cast<unsigned char *>(s_A) + Sm::kTabBytes;      // du staging
    unsigned char *s_stage1 = NQ == 2 ? s_keep + 2048 : s_stage0 + 64 * 16;                                   // ddelta staging
    const unsigned slot_u32 = smem_u32(s_slot), seed_u32 = smem_u32(s_seed);

    const int row0 = rg * kS4Rows, nrows = min(kS4Rows, D - row0);
    int row[2];
    bool row_ok[2];
    float bias[2], Dsk[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = row0 + lane + 32 * r;
        row_ok[r] = rr < D;
        row[r] = min(rr, D - 1);
        bias[r] = p.dbias != nullptr ? p.dbias[row[r]] : 0.f;
        Dsk[r] = p.Dv != nullptr ? p.Dv[row[r]] : 0.f;
    }
    float2 e2[16], dA2[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        s_A[n * 32 + lane] = make_float2(p.A[(int64_t)row[0] * 16 + n] * kLog2e, p.A[(int64_t)row[1] * 16 + n] * kLog2e);
        float ev[2] = {0.f, 0.f};
        if (p.ein != nullptr && seg + 1 < p.nseg) {
#pragma unroll
            for (int r = 0; r < 2; ++r) ev[r] = p.ein[(((int64_t)b * D + row[r]) * p.nseg + seg) * 16 + n];
        }
        e2[n] = make_float2(ev[0], ev[1]);
        dA2[n] = make_float2(0.f, 0.f);
    }
    const int s_begin = seg * p.sps, s_end = min(p.nstage, s_begin + p.sps);
    auto moff = [&](int s) { return REV ? L - T * (s + 1) : T * s; };

    const char *g_in[5];
    int64_t rs_in[5];
    g_in[0] = reinterpret_cast<const char *>(p.u) + ((int64_t)b * p.u_bs + (int64_t)row0 * p.u_ds) * ES, rs_in[0] = p.u_ds * ES;
    g_in[1] = reinterpret_cast<const char *>(p.delta) + ((int64_t)b * p.dl_bs + (int64_t)row0 * p.dl_ds) * ES, rs_in[1] = p.dl_ds * ES;
    g_in[2] = reinterpret_cast<const char *>(p.dout) + ((int64_t)b * p.g_bs + (int64_t)row0 * p.g_ds) * ES, rs_in[2] = p.g_ds * ES;
    g_in[3] = reinterpret_cast<const char *>(p.z) + ((int64_t)b * p.z_bs + (int64_t)row0 * p.z_ds) * ES, rs_in[3] = p.z_ds * ES;
    g_in[4] = reinterpret_cast<const char *>(p.ysave) + ((int64_t)b * p.y_bs + (int64_t)row0 * p.y_ds) * ES, rs_in[4] = p.y_ds * ES;
    const char *g_x = reinterpret_cast<const char *>(p.x) + ((int64_t)b * D + row0) * p.nx * 64;          // x[b][row][k][16] fp32: 64 bytes per (row, k)
    const int64_t rs_x = (int64_t)p.nx * 64;
    // dB / dC: after the transpose lane pair (2k, 2k+1) holds token (k & 7) of dB (lanes 0..15) or dC (lanes 16..31)
    float *g_dbc = (lane & 16) ? p.dC + (int64_t)b * p.dC_bs : p.dB + (int64_t)b * p.dB_bs;
    const int64_t dbc_ns = (lane & 16) ? p.dC_ns : p.dB_ns;
    const int dbc_tok = REV ? 7 - ((lane >> 1) & 7) : ((lane >> 1) & 7);

    const int tn = lane & 15;
    const IN_T *t_src = lane < 16 ? reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)tn * p.B_ns
                                  : reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)tn * p.C_ns;
    uint4 treg[NQ];
    auto tile_ldg = [&](int s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) treg[q] = ldg16_pf(reinterpret_cast<const uint4 *>(t_src + moff(s)) + q);
    };
    auto tile_sts = [&]() {
        float ev[8], v[8];
        Raw8<IN_T>::unpack(treg, ev);
        order8<REV>(ev, v);
        float4 *d = reinterpret_cast<float4 *>(s_tile + lane * 8);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
    };
    auto issue_stage = [&](int s, int par) {
        const int64_t mo = (int64_t)moff(s) * ES;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            if (t >= 3 && !has_z) continue;
            rb_load_async<NQ>(slot_u32 + t * Sm::kInBytes, g_in[t] + mo, rs_in[t], nrows, NQ, lane);
        }
        if (s > 0) rb_load_async<4>(seed_u32 + par * (Sm::kSeedBytes / 2), g_x + (int64_t)(s - 1) * 64, rs_x, nrows, 4, lane);   // state entering stage s = x[s - 1]
    };
    auto load_slot = [&](int t, int r, float (&v)[T]) {
        uint4 q[NQ];
        const unsigned char *base = s_slot + t * Sm::kInBytes;
#pragma unroll
        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(base + rb_unit<NQ>(lane + 32 * r, k) * 16);
        float ev[8];
        Raw8<IN_T>::unpack(q, ev);
        order8<REV>(ev, v);
    };
    // 8 results of row r -> a staging area in the rb layout of an IN_T tensor (memory order), stored cooperatively afterwards
    auto stage_out = [&](unsigned char *area, int r, const float (&v)[T]) {
        float ev[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ev[REV ? 7 - i : i] = v[i];
        uint4 q[NQ];
        Raw8<IN_T>::pack(ev, q);
#pragma unroll
        for (int k = 0; k < NQ; ++k) *reinterpret_cast<uint4 *>(area + rb_unit<NQ>(lane + 32 * r, k) * 16) = q[k];
    };
    // keep: fp32, two 16-byte pieces per row (P = 2): [which][64 rows x 2]
    auto keep_ptr = [&](int which, int r, int q) { return reinterpret_cast<float4 *>(s_keep + which * 2048 + rb_unit<2>(lane + 32 * r, q) * 16); };

    issue_stage(s_end - 1, 0);
    cp_async_commit();
    tile_ldg(s_end - 1);
    tile_sts();
    float2 dDa = make_float2(0.f, 0.f), dba = make_float2(0.f, 0.f);

    for (int s = s_end - 1; s >= s_begin; --s) {
        const int par = (s_end - 1 - s) & 1;
        const int64_t mo = (int64_t)moff(s) * ES;
        cp_async_wait_all();
        __syncwarp();                   // stage s has landed (cooperative copies), tile of stage s is visible
        // ---- prologue: per (row, token) quantities of my 8 tokens -------------------------------------------------------------------
        float2 dl[T], dlu[T], dy[T];
        {
            float uu[2][T], dd[2][T], gg[2][T];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                load_slot(0, r, uu[r]);
                load_slot(1, r, dd[r]);
                load_slot(2, r, gg[r]);
                float sg[T];
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float xx = dd[r][i] + bias[r];
                    float ex;
                    const float spv = softplus3(xx, ex);
                    dd[r][i] = sp ? spv : xx;
                    sg[i] = sp ? sigmoid_from_e(xx, ex) : 1.f;
                    if (!row_ok[r]) uu[r][i] = 0.f, gg[r][i] = 0.f;
                }
                if (has_z) {
                    float zz[T], yy[T], dzv[T];
                    load_slot(3, r, zz);
                    load_slot(4, r, yy);
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        const float sz = sigmoid3(zz[i]);
                        const float g = gg[r][i];
                        dzv[i] = g * yy[i] * sz * (1.f + zz[i] * (1.f - sz));
                        gg[r][i] = g * zz[i] * sz;
                    }
                    stage_out(s_slot + 4 * Sm::kInBytes, r, dzv);        // dz takes the place of the (consumed) y piece
                }
                *keep_ptr(0, r, 0) = make_float4(uu[r][0], uu[r][1], uu[r][2], uu[r][3]);
                *keep_ptr(0, r, 1) = make_float4(uu[r][4], uu[r][5], uu[r][6], uu[r][7]);
                *keep_ptr(1, r, 0) = make_float4(sg[0], sg[1], sg[2], sg[3]);
                *keep_ptr(1, r, 1) = make_float4(sg[4], sg[5], sg[6], sg[7]);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (!row_ok[r]) {       // rows past dim: their slots were never filled - keep the garbage out of the cross-lane sums
#pragma unroll
                    for (int i = 0; i < T; ++i) dd[r][i] = 0.f, uu[r][i] = 0.f, gg[r][i] = 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < T; ++i) {
                dl[i] = make_float2(dd[0][i], dd[1][i]);
                dlu[i] = make_float2(dd[0][i] * uu[0][i], dd[1][i] * uu[1][i]);
                dy[i] = make_float2(gg[0][i], gg[1][i]);
                dDa = ffma2(dy[i], make_float2(uu[0][i], uu[1][i]), dDa);
            }
        }
        __syncwarp();                   // every lane has consumed its slots; the dz pieces are complete
        if (has_z)
            rb_store<NQ>(s_slot + 4 * Sm::kInBytes, reinterpret_cast<char *>(p.dz) + ((int64_t)b * p.dz_bs + (int64_t)row0 * p.dz_ds) * ES + mo,
                         p.dz_ds * ES, nrows, NQ, lane);
        __syncwarp();                   // dz has been read back: the slots may be refilled
        if (s > s_begin) {
            issue_stage(s - 1, par ^ 1);
            tile_ldg(s - 1);
        }
        cp_async_commit();

        float2 S1[T], S2[T];
#pragma unroll
        for (int i = 0; i < T; ++i) S1[i] = make_float2(0.f, 0.f), S2[i] = make_float2(0.f, 0.f);
        float *dbc_p = g_dbc + moff(s) + dbc_tok;
        float4 sd[2];

#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float2 An = s_A[n * 32 + lane];
            if ((n & 3) == 0) {         // the states entering my 8 tokens, 4 at a time
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    sd[r] = (s > 0 && row_ok[r]) ? *reinterpret_cast<const float4 *>(s_seed + par * (Sm::kSeedBytes / 2) + rb_unit<4>(lane + 32 * r, n >> 2) * 16)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float2 hs = (n & 3) == 0 ? make_float2(sd[0].x, sd[1].x) : (n & 3) == 1 ? make_float2(sd[0].y, sd[1].y)
                            : (n & 3) == 2 ? make_float2(sd[0].z, sd[1].z) : make_float2(sd[0].w, sd[1].w);
            float Bn[T], Cn[T];
            {
                const float4 *tb = reinterpret_cast<const float4 *>(s_tile + n * 8), *tc = reinterpret_cast<const float4 *>(s_tile + (16 + n) * 8);
                const float4 b0 = tb[0], b1 = tb[1], c0 = tc[0], c1 = tc[1];
                Bn[0] = b0.x, Bn[1] = b0.y, Bn[2] = b0.z, Bn[3] = b0.w, Bn[4] = b1.x, Bn[5] = b1.y, Bn[6] = b1.z, Bn[7] = b1.w;
                Cn[0] = c0.x, Cn[1] = c0.y, Cn[2] = c0.z, Cn[3] = c0.w, Cn[4] = c1.x, Cn[5] = c1.y, Cn[6] = c1.z, Cn[7] = c1.w;
            }
            // forward recompute
            float2 a[T], hh[T];
#pragma unroll
            for (int i = 0; i < T; ++i) {
                a[i] = ex2(fmul2(dl[i], An));
                hh[i] = ffma2(a[i], i ? hh[i - 1] : hs, fmul2(dlu[i], splat(Bn[i])));
            }
            // reverse sweep
            float2 en = e2[n], dAn = dA2[n];
            float v[16];
#pragma unroll
            for (int i = T - 1; i >= 0; --i) {
                const float2 dh = ffma2(dy[i], splat(Cn[i]), en);
                en = fmul2(a[i], dh);
                const float2 db = fmul2(dh, dlu[i]), dc = fmul2(dy[i], hh[i]);
                v[i] = db.x + db.y;
                v[8 + i] = dc.x + dc.y;
                S1[i] = ffma2(dh, splat(Bn[i]), S1[i]);
                const float2 t = fmul2(dh, fmul2(a[i], i ? hh[i - 1] : hs));
                S2[i] = ffma2(t, An, S2[i]);
                dAn = ffma2(t, dl[i], dAn);
            }
            e2[n] = en, dA2[n] = dAn;
            // sum over the warp's 64 rows: recursive-halving transpose, lane pair (2k, 2k+1) ends with value k
            {
                const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
                float w8[8], w4[4], w2[2];
#pragma unroll
                for (int k = 0; k < 8; ++k) w8[k] = (h4 ? v[k + 8] : v[k]) + __shfl_xor_sync(0xffffffffu, h4 ? v[k] : v[k + 8], 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) w4[k] = (h3 ? w8[k + 4] : w8[k]) + __shfl_xor_sync(0xffffffffu, h3 ? w8[k] : w8[k + 4], 8);
#pragma unroll
                for (int k = 0; k < 2; ++k) w2[k] = (h2 ? w4[k + 2] : w4[k]) + __shfl_xor_sync(0xffffffffu, h2 ? w4[k] : w4[k + 2], 4);
                float w1 = (h1 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h1 ? w2[0] : w2[1], 2);
                w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                if (!(lane & 1)) atomicAdd(dbc_p, w1);
                dbc_p += dbc_ns;
            }
        }
        __syncwarp();                   // every lane is done with the tile of stage s
        if (s > s_begin) tile_sts();

        // ---- epilogue: du, ddelta (staged over the u / sigmoid pieces they were computed from, then stored cooperatively) -----------
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float4 u0 = *keep_ptr(0, r, 0), u1 = *keep_ptr(0, r, 1), g0 = *keep_ptr(1, r, 0), g1 = *keep_ptr(1, r, 1);
            const float uu[T] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w}, sg[T] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            float duv[T], ddv[T];
            float dbs = 0.f;
#pragma unroll
            for (int i = 0; i < T; ++i) {
                const float s1 = r ? S1[i].y : S1[i].x, s2 = (r ? S2[i].y : S2[i].x) * 0.69314718056f;   // A was scaled by log2(e)
                const float dli = r ? dl[i].y : dl[i].x, dyi = r ? dy[i].y : dy[i].x;
                duv[i] = fmaf(dli, s1, Dsk[r] * dyi);
                ddv[i] = fmaf(uu[i], s1, s2) * sg[i];
                dbs += ddv[i];
            }
            if (r) dba.y += dbs; else dba.x += dbs;
            stage_out(s_stage0, r, duv);
            stage_out(s_stage1, r, ddv);
        }
        __syncwarp();
        rb_store<NQ>(s_stage0, reinterpret_cast<char *>(p.du) + ((int64_t)b * p.du_bs + (int64_t)row0 * p.du_ds) * ES + mo, p.du_ds * ES, nrows, NQ, lane);
        rb_store<NQ>(s_stage1, reinterpret_cast<char *>(p.ddelta) + ((int64_t)b * p.ddl_bs + (int64_t)row0 * p.ddl_ds) * ES + mo, p.ddl_ds * ES,
                     nrows, NQ, lane);
        // (the next prologue's keep writes come after the __syncwarp at the top of the loop)
    }

    // ---- per-row parameter gradients --------------------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!row_ok[r]) continue;
#pragma unroll
        for (int n = 0; n < 16; ++n) atomicAdd(p.dA + (int64_t)row[r] * 16 + n, r ? dA2[n].y : dA2[n].x);
        if (p.dD != nullptr) atomicAdd(p.dD + row[r], r ? dDa.y : dDa.x);
        if (p.ddbias != nullptr) atomicAdd(p.ddbias + row[r], r ? dba.y : dba.x);
    }
}

}  // namespace mmu

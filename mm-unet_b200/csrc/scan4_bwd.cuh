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
    constexpr int T = 8, NTEN = 3;                  // delta | dout | z
    using Sm = S4Fwd<IN_T, NTEN>;
    constexpr int NQ = Sm::NQ, EPQ = 16 / (int)sizeof(IN_T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kS4W + warp;
    if (item >= p.nitems) return;
    const int rg = item % p.nrg, it1 = item / p.nrg, b = it1 % p.B, seg = 1 + it1 / p.B;   // segment 0 has nothing to its left
    const int D = p.D, L = p.L, N = p.N;
    const bool has_z = p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *s_slot = smem_raw + (size_t)warp * Sm::kWarpBytes;
    float *s_tile = reinterpret_cast<float *>(s_slot + Sm::kSlotBytes);      // [8 tokens][16]  C only
    const unsigned slot_u32 = smem_u32(s_slot) + lane * 16;
    const unsigned char *slot_t = s_slot + lane * 16;

    int row[2];
    bool row_ok[2];
    float2 A2[16], e[16];
    float bias[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = rg * kS4Rows + lane + 32 * r;
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
    auto moff = [&](int s) { return REV ? L - T * (s + 1) : T * s; };
    const IN_T *src[NTEN][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        src[0][r] = reinterpret_cast<const IN_T *>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)row[r] * p.dl_ds;
        src[1][r] = reinterpret_cast<const IN_T *>(p.dout) + (int64_t)b * p.g_bs + (int64_t)row[r] * p.g_ds;
        src[2][r] = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs + (int64_t)row[r] * p.z_ds : nullptr;
    }
    const int tn = lane & 15;
    const bool t_live = tn < N && lane < 16;
    const IN_T *t_src = reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)tn * p.C_ns;
    uint4 treg[NQ];
    auto tile_ldg = [&](int s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            treg[q] = t_live ? ldg16_pf(reinterpret_cast<const uint4 *>(t_src + moff(s)) + q) : make_uint4(0u, 0u, 0u, 0u);
    };
    auto tile_sts = [&]() {
        float v[8];
        Raw8<IN_T>::unpack(treg, v);
        if (lane < 16) {
#pragma unroll
            for (int k = 0; k < 8; ++k) s_tile[(REV ? 7 - k : k) * 16 + lane] = v[k];
        }
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
        float ev[8];
        Raw8<IN_T>::unpack(q, ev);
        order8<REV>(ev, v);
    };

    issue_elems(s_end - 1, 0);
    cp_async_commit();
    tile_ldg(s_end - 1);
    tile_sts();
    float dsum[2] = {0.f, 0.f};

    for (int s = s_end - 1; s >= s_begin; --s) {
        const int par = (s_end - 1 - s) & 1;
        cp_async_wait_all();
        __syncwarp();
        if (s > s_begin) {
            issue_elems(s - 1, par ^ 1);
            tile_ldg(s - 1);
        }
        cp_async_commit();
        float dd[2][T], gy[2][T];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            load_slot(par, 0, r, dd[r]);
            load_slot(par, 1, r, gy[r]);
            float zz[T];
            if (has_z) load_slot(par, 2, r, zz);
#pragma unroll
            for (int i = 0; i < T; ++i) {
                const float xx = dd[r][i] + bias[r];
                dd[r][i] = sp ? softplus3(xx) : xx;
                dsum[r] += dd[r][i];
                if (has_z) gy[r][i] *= zz[i] * sigmoid3(zz[i]);
            }
        }
#pragma unroll
        for (int i = T - 1; i >= 0; --i) {
            const float2 dl = make_float2(dd[0][i], dd[1][i]);
            const float2 dy = make_float2(gy[0][i], gy[1][i]);
            const float4 *tb = reinterpret_cast<const float4 *>(s_tile + i * 16);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 c4 = tb[g];
                const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int n = 4 * g + k;
                    const float2 a = ex2(fmul2(dl, A2[n]));
                    e[n] = fmul2(a, ffma2(dy, splat(cv[k]), e[n]));
                }
            }
        }
        __syncwarp();
        if (s > s_begin) tile_sts();
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
// shared memory of one warp
template <typename IN_T> struct S4Bwd {
    static constexpr int NQ = Raw8<IN_T>::kQuads;
    static constexpr int kSlotBytes = 5 * 2 * NQ * 32 * 16;          // u | delta | dout | z | y : [tensor][row][quad][lane] x 16 B
    static constexpr int kKeepBytes = 2 * 2 * 2 * 32 * 16;           // u, sigmoid(delta_raw + bias) as fp32: [which][row][quad][lane] x 16 B
    static constexpr int kSeedBytes = 2 * 2 * 16 * 32 * 4;           // [parity][row][state][lane] fp32
    static constexpr int kTileBytes = 32 * 8 * 4;                    // [B states 0..15 | C states 0..15][8 tokens] fp32, logical order
    static constexpr int kTabBytes = 16 * 32 * 8;                    // A*log2e of (row 0, row 1): [state][lane] float2
    static constexpr int kWarpBytes = kSlotBytes + kKeepBytes + kSeedBytes + kTileBytes + kTabBytes;
};

template <typename IN_T, bool REV>
__global__ void __launch_bounds__(32 * kS4W, 2) scan4_bwd_kernel(const __grid_constant__ Bwd4Args p) {
    constexpr int T = 8;
    using Sm = S4Bwd<IN_T>;
    constexpr int NQ = Sm::NQ, EPQ = 16 / (int)sizeof(IN_T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kS4W + warp;
    if (item >= p.nitems) return;
    const int rg = item % p.nrg, it1 = item / p.nrg, b = it1 % p.B, seg = it1 / p.B;
    const int D = p.D, L = p.L, N = p.N;
    const bool has_z = p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *s_slot = smem_raw + (size_t)warp * Sm::kWarpBytes;
    unsigned char *s_keep = s_slot + Sm::kSlotBytes;
    float *s_seed = reinterpret_cast<float *>(s_keep + Sm::kKeepBytes);
    float *s_tile = s_seed + Sm::kSeedBytes / 4;
    float2 *s_A = reinterpret_cast<float2 *>(s_tile + Sm::kTileBytes / 4);
    const unsigned slot_u32 = smem_u32(s_slot) + lane * 16;
    const unsigned char *slot_t = s_slot + lane * 16;
    const unsigned seed_u32 = smem_u32(s_seed) + lane * 4;

    int row[2];
    bool row_ok[2];
    float bias[2], Dsk[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rr = rg * kS4Rows + lane + 32 * r;
        row_ok[r] = rr < D;
        row[r] = min(rr, D - 1);
        bias[r] = p.dbias != nullptr ? p.dbias[row[r]] : 0.f;
        Dsk[r] = p.Dv != nullptr ? p.Dv[row[r]] : 0.f;
    }
    float2 e2[16], dA2[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        s_A[n * 32 + lane] = make_float2(n < N ? p.A[(int64_t)row[0] * N + n] * kLog2e : 0.f, n < N ? p.A[(int64_t)row[1] * N + n] * kLog2e : 0.f);
        float ev[2] = {0.f, 0.f};
        if (p.ein != nullptr && seg + 1 < p.nseg && n < N) {
#pragma unroll
            for (int r = 0; r < 2; ++r) ev[r] = p.ein[(((int64_t)b * D + row[r]) * p.nseg + seg) * 16 + n];
        }
        e2[n] = make_float2(ev[0], ev[1]);
        dA2[n] = make_float2(0.f, 0.f);
    }
    const int s_begin = seg * p.sps, s_end = min(p.nstage, s_begin + p.sps);
    auto moff = [&](int s) { return REV ? L - T * (s + 1) : T * s; };

    const IN_T *src[5][2];
    const float *x_row[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        src[0][r] = reinterpret_cast<const IN_T *>(p.u) + (int64_t)b * p.u_bs + (int64_t)row[r] * p.u_ds;
        src[1][r] = reinterpret_cast<const IN_T *>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)row[r] * p.dl_ds;
        src[2][r] = reinterpret_cast<const IN_T *>(p.dout) + (int64_t)b * p.g_bs + (int64_t)row[r] * p.g_ds;
        src[3][r] = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs + (int64_t)row[r] * p.z_ds : nullptr;
        src[4][r] = has_z ? reinterpret_cast<const IN_T *>(p.ysave) + (int64_t)b * p.y_bs + (int64_t)row[r] * p.y_ds : nullptr;
        x_row[r] = p.x != nullptr ? p.x + ((int64_t)b * D + row[r]) * p.nx * N : nullptr;
    }
    const int tn = lane & 15;
    const bool t_live = tn < N;
    const IN_T *t_src = lane < 16 ? reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)tn * p.B_ns
                                  : reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)tn * p.C_ns;
    uint4 treg[NQ];
    auto tile_ldg = [&](int s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            treg[q] = t_live ? ldg16_pf(reinterpret_cast<const uint4 *>(t_src + moff(s)) + q) : make_uint4(0u, 0u, 0u, 0u);
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
        const int mo = moff(s);
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            if (t >= 3 && !has_z) continue;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < NQ; ++q) cp_async16_pf(slot_u32 + (((t * 2 + r) * NQ + q) * 32) * 16, src[t][r] + mo + q * EPQ);
        }
        if (s > 0) {        // state entering stage s = x[s - 1]
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float *xp = x_row[r] + (int64_t)(s - 1) * N;
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    if (n < N) cp_async4_pf(seed_u32 + (((par * 2 + r) * 16 + n) * 32) * 4, xp + n);
            }
        }
    };
    auto load_slot = [&](int t, int r, float (&v)[T]) {
        uint4 q[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(slot_t + (((t * 2 + r) * NQ + k) * 32) * 16);
        float ev[8];
        Raw8<IN_T>::unpack(q, ev);
        order8<REV>(ev, v);
    };
    float4 *keep = reinterpret_cast<float4 *>(s_keep) + lane;          // [which][row][quad] stride 32 float4

    issue_stage(s_end - 1, 0);
    cp_async_commit();
    tile_ldg(s_end - 1);
    tile_sts();
    float2 dDa = make_float2(0.f, 0.f), dba = make_float2(0.f, 0.f);

    for (int s = s_end - 1; s >= s_begin; --s) {
        const int par = (s_end - 1 - s) & 1;
        const int mo = moff(s);
        cp_async_wait_all();
        __syncwarp();
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
                    if (row_ok[r])
                        store8<IN_T, REV>(reinterpret_cast<IN_T *>(p.dz) + (int64_t)b * p.dz_bs + (int64_t)row[r] * p.dz_ds + mo, dzv);
                }
                keep[(0 * 2 + r) * 64] = make_float4(uu[r][0], uu[r][1], uu[r][2], uu[r][3]);
                keep[(0 * 2 + r) * 64 + 32] = make_float4(uu[r][4], uu[r][5], uu[r][6], uu[r][7]);
                keep[(1 * 2 + r) * 64] = make_float4(sg[0], sg[1], sg[2], sg[3]);
                keep[(1 * 2 + r) * 64 + 32] = make_float4(sg[4], sg[5], sg[6], sg[7]);
            }
#pragma unroll
            for (int i = 0; i < T; ++i) {
                dl[i] = make_float2(dd[0][i], dd[1][i]);
                dlu[i] = make_float2(dd[0][i] * uu[0][i], dd[1][i] * uu[1][i]);
                dy[i] = make_float2(gg[0][i], gg[1][i]);
                dDa = ffma2(dy[i], make_float2(uu[0][i], uu[1][i]), dDa);
            }
        }
        // the slots are consumed: fetch the next stage (one stage to the left)
        if (s > s_begin) {
            issue_stage(s - 1, par ^ 1);
            tile_ldg(s - 1);
        }
        cp_async_commit();

        float2 S1[T], S2[T];
#pragma unroll
        for (int i = 0; i < T; ++i) S1[i] = make_float2(0.f, 0.f), S2[i] = make_float2(0.f, 0.f);

#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float2 An = s_A[n * 32 + lane];
            float2 hs = make_float2(0.f, 0.f);
            if (s > 0) hs = make_float2(s_seed[((par * 2 + 0) * 16 + n) * 32 + lane], s_seed[((par * 2 + 1) * 16 + n) * 32 + lane]);
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
                if (!(lane & 1) && n < N) {
                    const int i = (lane >> 1) & 7;                      // token of my value; lanes 16.. hold dC
                    float *dst = (lane & 16) ? p.dC + (int64_t)b * p.dC_bs + (int64_t)n * p.dC_ns : p.dB + (int64_t)b * p.dB_bs + (int64_t)n * p.dB_ns;
                    atomicAdd(dst + mo + (REV ? 7 - i : i), w1);
                }
            }
        }
        __syncwarp();                   // every lane is done with the tile of stage s
        if (s > s_begin) tile_sts();

        // ---- epilogue: du, ddelta ----------------------------------------------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float4 u0 = keep[(0 * 2 + r) * 64], u1 = keep[(0 * 2 + r) * 64 + 32], g0 = keep[(1 * 2 + r) * 64], g1 = keep[(1 * 2 + r) * 64 + 32];
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
            if (row_ok[r]) {
                store8<IN_T, REV>(reinterpret_cast<IN_T *>(p.du) + (int64_t)b * p.du_bs + (int64_t)row[r] * p.du_ds + mo, duv);
                store8<IN_T, REV>(reinterpret_cast<IN_T *>(p.ddelta) + (int64_t)b * p.ddl_bs + (int64_t)row[r] * p.ddl_ds + mo, ddv);
            }
        }
    }

    // ---- per-row parameter gradients --------------------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!row_ok[r]) continue;
#pragma unroll
        for (int n = 0; n < 16; ++n)
            if (n < N) atomicAdd(p.dA + (int64_t)row[r] * N + n, r ? dA2[n].y : dA2[n].x);
        if (p.dD != nullptr) atomicAdd(p.dD + row[r], r ? dDa.y : dDa.x);
        if (p.ddbias != nullptr) atomicAdd(p.ddbias + row[r], r ? dba.y : dba.x);
    }
}

}  // namespace mmu

// Selective scan forward, v5 "lane ring" kernel (fp32 / bf16 I/O, dstate <= 16, seqlen % 8 == 0, 16-byte aligned rows, one segment,
// saved states every 64 tokens; plain, flipped (REV) or fused two-row (ORD) scan order).  Math: SURVEY.md Appendix A; replaces selective_scan_fwd_kernel (selective_scan_fwd_kernel.cuh:67-303).
//
// v3 (scan3_fwd.cuh) closes the chunk-level recurrence with a cross-lane Kogge-Stone scan of (prod a, h) plus a fix-up
// y_i += C_i Pc_i hs: 130 issue slots per (lane, state), 50 of them the scan, 23 the products it needs.  Here the lanes of a
// row pair form a RING that works as a systolic pipeline instead:
//   * 16 lanes own the 16 eight-token blocks of a 128-token chunk of a row pair (two rows packed in the FFMA2 halves);
//   * at step tau lane j works on item q = tau - j: chunk q / 16, state q % 16.  The state entering its block is what lane j-1
//     produced one step earlier (one SHFL.IDX per packed half); lane 0 receives lane 15's result, which is exactly the state
//     leaving the previous chunk for the same n - the ring carries the recurrence across chunks with no special case;
//   * so every lane runs the TRUE recurrence h = a h + delta u B, y += C h: no aggregates, no scan, no fix-up, one MUFU.EX2 per
//     (row, token, state), ~50 issue slots per (lane, state) for the recurrence itself.
// The price: the lanes of a ring sit in two different chunks at any time (lane j enters chunk k at step 16k + j), so the
// per-(row, token) quantities cannot simply be reloaded by the whole warp at a chunk boundary.  A lane picks its new delta /
// delta*u up from a shared-memory stage at ITS transition step (predicated loads), and swaps its y accumulators with the D*u
// of the new chunk there.  Work per warp is fixed by the problem (rows / 4 ring warps, each walking the whole sequence), so this
// kernel is for wide problems; narrow ones stay on v3, which can cut the sequence into segments.
//
// Warp specialisation.  A first version (git history: b08aac9) let every warp do its own element-wise work in a uniform "bulk"
// phase once per round of 16 steps: 195 us at BASELINE config 2 against v3's 172 us - ncu showed the steps running at 0.57
// instructions per clock and scheduler, but the bulk phase taking 40 % of the round and the schedulers of an SM carrying 2, 2, 1, 1
// ring warps.  Here a ring warp does NOTHING but steps; a HELPER warp per ring warp does every element-wise phase for it one round
// ahead / two rounds behind, through double-buffered lane slots (twice the warps per SM for the same work):
//   helper, round k:  gate + store of chunk k-2 from yout[(k-2)&1];  saved states of chunk k-2 from xbuf[(k-2)%3];  softplus /
//                     delta*u / D*u of chunk k+1 -> stage[(k+1)&1], du[(k+1)&1];  cp.async of u, delta (chunk k+2), z (chunk k-1)
//                     into its landing slots and of its share of the B/C tile k+1 (3-deep ring of tiles);
//   ring, round k:    16 steps on stage[k&1]; at the lane's transition the finished y -> yout[(k-1)&1] and D*u of the new chunk
//                     <- du[k&1]; the states leaving blocks 7 and 15 of a chunk (stride-64 saved states) -> xbuf[chunk%3].
// One __syncthreads per round orders everything.  The two rings of a warp are interleaved over the lanes (lane = 2 j + ring): the
// two lanes in transition at a step then sit in the same quarter-warp and their predicated 16-byte accesses are one wavefront.
// The step loop is a runtime loop over step PAIRS (operand double buffer): fully unrolled, ptxas split the live ranges of delta,
// delta*u and y across the 16 copies and paid ~50 predicated moves per step.
// Measured (B200, fp32, profiles/r2_v5_lane_ring.md): config 2 (B8 D384 L4096) 154-159 us vs v3 172 us; B16 D128 L 4k / 16k / 64k
// 117 / 433 / 1627 us vs 138 / 516 / 2034 us.  Time per round is ~3.2 us with 4 ring warps per CTA and ~4.4 us with 6 (6 ring warps
// sit 2, 2, 1, 1 on the four schedulers), independent of the number of CTAs up to one per SM.  2-byte I/O gains nothing over v3
// (latency bound): opt-in.
#pragma once
#include "scan3.cuh"
#include "scan3_fwd.cuh"

#ifndef MMU_S5_WHATIF
#define MMU_S5_WHATIF 0      // timing probe only (wrong results): half of the exponentials skipped
#endif

namespace mmu {

template <typename IN_T, int WR> struct Fwd5Cfg {
    static constexpr int T = kS3T, LPR = 16, CH = LPR * T, NRT = 32 * WR, NT = 2 * NRT, NRP = 2 * WR, R = 4 * WR;
    static constexpr bool kF32 = sizeof(IN_T) == 4;
    static constexpr int NQ = Raw8<IN_T>::kQuads;
    using Tl = BcTile<LPR>;
    static constexpr int kTiles = 3;
    static constexpr int kTileBytes = kTiles * Tl::kBytes;
    static constexpr int kRawBytes = kF32 ? 0 : 2 * (2 * 16 * CH * 2);     // 2-byte B/C rows as they sit in memory, double buffered
    static constexpr int kLandBytes = 3 * 2 * NQ * NRT * 16;   // u | delta | z : [tensor][row][quad][ring thread] x 16 B
    static constexpr int kStageBytes = 2 * 8 * NRT * 16;       // [parity][delta x4 | delta*u x4][ring thread]
    static constexpr int kDuBytes = 2 * 4 * NRT * 16;          // D*u of the chunk a lane is about to enter: [parity][quad][ring thread]
    static constexpr int kYBytes = 2 * 4 * NRT * 16;           // finished y of the chunk a lane left: [parity][quad][ring thread]
    static constexpr int kXBytes = 3 * 16 * NRP * 2 * 8;       // [chunk % 3][state][ring][block 7 | block 15] float2
    static constexpr int kABytes = NRP * 16 * 8;
    static constexpr size_t smem_bytes = (size_t)kTileBytes + kRawBytes + kLandBytes + kStageBytes + kDuBytes + kYBytes + kXBytes + kABytes;
};

// ORD: fused scan order (as scan3_fwd_kernel; the host selects it for TWOROW only) - the helpers gather the gate z and scatter out through p.ord
// (ord_issue8 / ord_store8 in scan3.cuh); everything the ring warps touch stays in scan order.  Never together with REV.
// ORD == 2: NSLICES for 4-byte elements with 8 / 16 / 32 slices and L % 128 == 0, through a transposition in the z landing slots: a
// chunk of a row is 128 / ns consecutive elements in each of the ns slices, i.e. 32 sixteen-byte pieces - two per lane, as in the plain
// order - which the lanes read / write at their logical tokens' positions (the generic element-wise form of ord_issue8 / ord_store8,
// eight 4-byte pieces per lane and row, measured 187 us at config 2 against 158 us for the plain order).
template <typename IN_T, int WR, bool REV, int ORD = 0>
__global__ void __launch_bounds__(64 * WR, 1) scan5_fwd_kernel(const __grid_constant__ Fwd3Args p) {
    static_assert(!ORD || !REV, "ordered gate / output: forward direction only");
    static_assert(ORD != 2 || sizeof(IN_T) == 4, "tiled nslices: 4-byte elements");
    using Cfg = Fwd5Cfg<IN_T, WR>;
    constexpr bool kF32 = Cfg::kF32;
    constexpr int NQ = Cfg::NQ, EPQ = 16 / (int)sizeof(IN_T);
    using Tl = typename Cfg::Tl;
    constexpr int T = Cfg::T, LPR = Cfg::LPR, CH = Cfg::CH, NRT = Cfg::NRT, NRP = Cfg::NRP, R = Cfg::R;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = lane >> 1, ring = lane & 1;               // interleaved rings
    const bool helper = warp >= WR;
    const int rw = helper ? warp - WR : warp;               // the ring warp I am / I serve
    const int rt = rw * 32 + lane;                          // ring-thread slot
    const int rp = rw * 2 + ring;
    const int b = blockIdx.y, row0 = blockIdx.x * R;
    const int D = p.D, L = p.L, N = p.N;
    const int NC = (L + CH - 1) / CH;
    const bool has_z = p.z != nullptr, sp = p.softplus != 0;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_tile = smem_raw;
    unsigned char *s_rawbc = s_tile + Cfg::kTileBytes;      // 2-byte inputs only
    unsigned char *s_land = s_rawbc + Cfg::kRawBytes;
    float4 *s_stage = reinterpret_cast<float4 *>(s_land + Cfg::kLandBytes);                                  // [2][8][NRT]
    float4 *s_du = reinterpret_cast<float4 *>(s_land + Cfg::kLandBytes + Cfg::kStageBytes);                  // [2][4][NRT]
    float4 *s_yout = s_du + 2 * 4 * NRT;                                                                     // [2][4][NRT]
    float2 *s_x = reinterpret_cast<float2 *>(s_yout + 2 * 4 * NRT);                                          // [3][16][NRP][2]
    float2 *s_A = s_x + 3 * 16 * NRP * 2;                                                                    // [NRP][16]

    for (int i = tid; i < (int)((Cfg::smem_bytes - Cfg::kABytes) / 16); i += Cfg::NT) reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < NRP * 16; i += Cfg::NT) {
        const int g = i >> 4, n = i & 15;
        float a[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = row0 + 2 * g + r;
            a[r] = (row < D && n < N) ? p.A[(int64_t)row * N + n] * kLog2e : 0.f;
        }
        s_A[i] = make_float2(a[0], a[1]);
    }
    __syncthreads();

    if (helper) {
        // ===================================== helper warp: every element-wise phase of ring warp rw ==================================
        const int htid = tid - NRT;
        const int rowA = row0 + 2 * rp;
        const bool rok[2] = {rowA < D, rowA + 1 < D};
        constexpr int STEP = REV ? -CH : CH;
        const int mo0 = REV ? L - T - T * j : T * j;        // memory index of the lane's 8 tokens in chunk 0
        const IN_T *u_p[2], *d_p[2], *z_p[2];               // stand on the chunk whose copy is issued next
        IN_T *o_p[2], *y_p[2];                              // stand on the chunk stored next
        float bias[2], Dsk[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = min(rowA + r, D - 1);
            u_p[r] = reinterpret_cast<const IN_T *>(p.u) + (int64_t)b * p.u_bs + (int64_t)row * p.u_ds + mo0;
            d_p[r] = reinterpret_cast<const IN_T *>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)row * p.dl_ds + mo0;
            // ORD: row bases (the token offset goes through the index map); else the lane's 8 tokens of chunk 0
            z_p[r] = has_z ? reinterpret_cast<const IN_T *>(p.z) + (int64_t)b * p.z_bs + (int64_t)row * p.z_ds + (ORD ? 0 : mo0) : nullptr;
            o_p[r] = reinterpret_cast<IN_T *>(p.out) + (int64_t)b * p.o_bs + (int64_t)row * p.o_ds + (ORD ? 0 : mo0);
            y_p[r] = p.ysave == nullptr ? nullptr : reinterpret_cast<IN_T *>(p.ysave) + (int64_t)b * p.y_bs + (int64_t)row * p.y_ds + mo0;
            bias[r] = p.dbias != nullptr ? p.dbias[row] : 0.f;
            Dsk[r] = p.Dv != nullptr ? p.Dv[row] : 0.f;
        }
        const IN_T *B_b = reinterpret_cast<const IN_T *>(p.Bm) + (int64_t)b * p.B_bs;
        const IN_T *C_b = reinterpret_cast<const IN_T *>(p.Cm) + (int64_t)b * p.C_bs;
        const unsigned s_tile_u32 = smem_u32(s_tile), s_raw_u32 = smem_u32(s_rawbc);
        const unsigned s_land_u32 = smem_u32(s_land) + rt * 16;
        const unsigned char *s_land_t = s_land + rt * 16;
        // tiled nslices (ORD == 2): 16-byte piece pc (0..31) of a chunk = slice pc / pps, quad pc % pps of its 128 / ns elements
        [[maybe_unused]] const int ns_sh = p.ord.ns_shift;
        [[maybe_unused]] auto ns_piece_off = [&](int pc, int c) {
            return (int64_t)(pc >> (5 - ns_sh)) * p.ord.Ls + ((c * CH) >> ns_sh) + 4 * (pc & ((32 >> ns_sh) - 1));
        };
        // landing-slot address (tensor z, row r) of the element that holds token i of my 8 logical tokens
        [[maybe_unused]] auto ns_elem = [&](int r, int i) {
            const int t = T * j + i, sl = t & ((1 << ns_sh) - 1), jj = t >> ns_sh;
            const int pc = (sl << (5 - ns_sh)) + (jj >> 2);
            return reinterpret_cast<float *>(s_land + ((2 * 2 + r) * NQ + (pc & 1)) * NRT * 16 + (rw * 32 + 2 * (pc >> 1) + ring) * 16) + (jj & 3);
        };
        auto issue_ud = [&](int c) {
            if (c * CH + T * j < L) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        cp_async16(s_land_u32 + ((0 * 2 + r) * NQ + q) * NRT * 16, u_p[r] + q * EPQ);
                        cp_async16(s_land_u32 + ((1 * 2 + r) * NQ + q) * NRT * 16, d_p[r] + q * EPQ);
                    }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) u_p[r] += STEP, d_p[r] += STEP;
        };
        auto issue_z = [&](int c) {
            if (!has_z) return;
            if constexpr (ORD == 2) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int q = 0; q < 2; ++q) cp_async16(s_land_u32 + ((2 * 2 + r) * NQ + q) * NRT * 16, z_p[r] + ns_piece_off(2 * j + q, c));
                return;
            } else if constexpr (ORD == 1) {
                if (c * CH + T * j < L) {
#pragma unroll
                    for (int r = 0; r < 2; ++r)
                        ord_issue8<IN_T>(p.ord, c * CH + T * j, z_p[r], s_land_u32 + ((2 * 2 + r) * NQ) * NRT * 16, s_land_u32 + ((2 * 2 + r) * NQ + NQ - 1) * NRT * 16);
                }
                return;
            }
            if (c * CH + T * j < L) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int q = 0; q < NQ; ++q) cp_async16(s_land_u32 + ((2 * 2 + r) * NQ + q) * NRT * 16, z_p[r] + q * EPQ);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) z_p[r] += STEP;
        };
        auto load_land = [&](int which, int r, float (&v)[T]) {
            uint4 q[NQ];
#pragma unroll
            for (int k = 0; k < NQ; ++k) q[k] = *reinterpret_cast<const uint4 *>(s_land_t + ((which * 2 + r) * NQ + k) * NRT * 16);
            float e[8];
            Raw8<IN_T>::unpack(q, e);
            order8<REV>(e, v);
        };
        // B/C tile of chunk c.  fp32: cp.async straight into tile c % 3.  2-byte inputs: the raw rows go to raw[c & 1] and are widened
        // into tile c % 3 by the helpers ONE ROUND LATER (widen_tile), so a tile is requested two rounds before its first use.
        auto issue_tile = [&](int c) {
            if constexpr (kF32) {
                tile_async_f32<LPR, NRT, REV, true>(s_tile_u32 + (unsigned)(c % Cfg::kTiles) * Tl::kBytes, reinterpret_cast<const float *>(B_b),
                                                    reinterpret_cast<const float *>(C_b), p.B_ns, p.C_ns, N, c * CH, L, htid);
            } else {
                raw_async_bf16<LPR, NRT, REV, true>(s_raw_u32 + (unsigned)(c & 1) * (Cfg::kRawBytes / 2), reinterpret_cast<const __nv_bfloat16 *>(B_b),
                                                    reinterpret_cast<const __nv_bfloat16 *>(C_b), p.B_ns, p.C_ns, N, c * CH, L, htid);
            }
        };
        auto widen_tile = [&](int c) {
            if constexpr (!kF32) widen_bf16_tile<LPR, NRT, true>(s_tile + (c % Cfg::kTiles) * Tl::kBytes, s_rawbc + (c & 1) * (Cfg::kRawBytes / 2), htid);
        };
        issue_ud(0);
        issue_tile(0);
        cp_async_commit();
        if constexpr (!kF32) {      // the raw tile of chunk 0 has to be widened before round 0: one extra start-up round trip
            cp_async_wait_all();
            // (helpers only: named barrier 1)
            asm volatile("bar.sync 1, %0;" ::"r"(NRT) : "memory");
            widen_tile(0);
            if (NC > 1) issue_tile(1);
            cp_async_commit();
        }

        for (int k = -1; k <= NC + 1; ++k) {
            cp_async_wait_all();
            __syncthreads();
            // ---- gate + store of chunk k-2 ------------------------------------------------------------------------------------------------
            const int ce = k - 2;
            if (ce >= 0) {
                const bool ok = ce * CH + T * j < L;
                const float4 *yo = s_yout + (ce & 1) * 4 * NRT + rt;
                float ya[2][T];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = yo[q * NRT];
                    ya[0][2 * q] = v.x, ya[1][2 * q] = v.y, ya[0][2 * q + 1] = v.z, ya[1][2 * q + 1] = v.w;
                }
                if constexpr (ORD == 2) {       // every lane runs the whole sequence: the __syncwarp()s are warp-wide
                    float zz[2][T];
                    if (has_z) {
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int i = 0; i < T; ++i) zz[r][i] = *ns_elem(r, i);
                    }
                    __syncwarp();               // z has been read: the slots now stage the output
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (rok[r] && y_p[r] != nullptr) store8<IN_T, REV>(y_p[r], ya[r]);
#pragma unroll
                        for (int i = 0; i < T; ++i) *ns_elem(r, i) = has_z ? ya[r][i] * (zz[r][i] * sigmoid3(zz[r][i])) : ya[r][i];
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (!rok[r]) continue;
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            *reinterpret_cast<uint4 *>(o_p[r] + ns_piece_off(2 * j + q, ce)) = *reinterpret_cast<const uint4 *>(s_land_t + ((2 * 2 + r) * NQ + q) * NRT * 16);
                    }
                    __syncwarp();               // before the next chunk's z lands in the slots
                } else if (ok) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (!rok[r]) continue;
                        if (y_p[r] != nullptr) store8<IN_T, REV>(y_p[r], ya[r]);
                        if (has_z) {
                            float zz[T];
                            if constexpr (ORD == 1) {
                                uint4 q[NQ];
#pragma unroll
                                for (int kq = 0; kq < NQ; ++kq) q[kq] = *reinterpret_cast<const uint4 *>(s_land_t + ((2 * 2 + r) * NQ + kq) * NRT * 16);
                                float e[8];
                                Raw8<IN_T>::unpack(q, e);
                                ord_to_tokens(p.ord, ce * CH + T * j, e, zz);
                            } else {
                                load_land(2, r, zz);
                            }
#pragma unroll
                            for (int i = 0; i < T; ++i) ya[r][i] *= zz[i] * sigmoid3(zz[i]);
                        }
                        if constexpr (ORD == 1) ord_store8<IN_T>(p.ord, ce * CH + T * j, o_p[r], ya[r]);
                        else store8<IN_T, REV>(o_p[r], ya[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (!ORD) o_p[r] += STEP;
                    if (y_p[r] != nullptr) y_p[r] += STEP;
                }
                // saved states: my float4 of the ring warp's 2 rings x {block 7, block 15} x 2 rows x 16 states
                if (p.x != nullptr || (p.last_state != nullptr && ce == NC - 1)) {
                    const int xr = lane >> 4, slot = (lane >> 3) & 1, r = (lane >> 2) & 1, quad = lane & 3;
                    const int rpx = rw * 2 + xr, row = row0 + 2 * rpx + r, kx = 2 * ce + slot;
                    const float2 *sx = s_x + (((ce % 3) * 16 + 4 * quad) * NRP + rpx) * 2 + slot;
                    float v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 t = sx[i * NRP * 2];
                        v[i] = r ? t.y : t.x;
                    }
                    if (row < D) {
                        if (p.x != nullptr && kx < p.nx) {
                            float *xp = p.x + (((int64_t)b * D + row) * p.nx + kx) * N + 4 * quad;
                            if (N == 16) {
                                *reinterpret_cast<float4 *>(xp) = make_float4(v[0], v[1], v[2], v[3]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (4 * quad + i < N) xp[i] = v[i];
                            }
                        }
                        if (p.last_state != nullptr && ce == NC - 1 && slot == 1) {     // blocks past L pass the final state through
                            float *lp = p.last_state + ((int64_t)b * D + row) * N + 4 * quad;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (4 * quad + i < N) lp[i] = v[i];
                        }
                    }
                }
            }
            // ---- softplus(delta + bias), delta*u, D*u of chunk k+1 -------------------------------------------------------------------
            const int cp = k + 1;
            if (cp <= NC) {
                float4 *st = s_stage + (cp & 1) * 8 * NRT + rt, *yi = s_du + (cp & 1) * 4 * NRT + rt;
                if (cp < NC) {
                    const bool ok = cp * CH + T * j < L;
                    float uu[2][T], dd[2][T];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        load_land(0, r, uu[r]);
                        load_land(1, r, dd[r]);
#pragma unroll
                        for (int i = 0; i < T; ++i) {
                            const float xx = dd[r][i] + bias[r];
                            const float v = sp ? softplus3(xx) : xx;
                            dd[r][i] = (ok && rok[r]) ? v : 0.f;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i0 = 2 * q, i1 = 2 * q + 1;
                        st[q * NRT] = make_float4(dd[0][i0], dd[1][i0], dd[0][i1], dd[1][i1]);
                        st[(4 + q) * NRT] = make_float4(dd[0][i0] * uu[0][i0], dd[1][i0] * uu[1][i0], dd[0][i1] * uu[0][i1], dd[1][i1] * uu[1][i1]);
                        yi[q * NRT] = make_float4(Dsk[0] * uu[0][i0], Dsk[1] * uu[1][i0], Dsk[0] * uu[0][i1], Dsk[1] * uu[1][i1]);
                    }
                } else {                // the drain round reads an all-zero stage (delta = 0: a = 1, b = 0)
#pragma unroll
                    for (int q = 0; q < 8; ++q) st[q * NRT] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) yi[q * NRT] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (k + 2 < NC) issue_ud(k + 2);
            if (k - 1 >= 0 && k - 1 < NC) issue_z(k - 1);
            if constexpr (kF32) {
                if (k + 1 >= 1 && k + 1 < NC) issue_tile(k + 1);
            } else {                // raw[(k+1)&1] landed before this round's barrier; raw[k&1] was widened last round
                if (k + 1 >= 1 && k + 1 < NC) widen_tile(k + 1);
                if (k + 2 >= 2 && k + 2 < NC) issue_tile(k + 2);
            }
            cp_async_commit();
        }
        return;
    }

    // ========================================= ring warp: nothing but steps =========================================================
    const int qa = REV ? 2 * (LPR - 1 - j) : 2 * j;         // my 8 tokens inside a tile row (memory order): two adjacent quads
    const float2 *s_A_rp = s_A + rp * 16;
    const int src_lane = 2 * ((j + 15) & 15) + ring;        // ring predecessor
    const bool xsave = (j & 7) == 7;                        // my block ends a 64-token group: its leaving state is a saved state
    float2 dl[T], dlu[T], ya[T];
#pragma unroll
    for (int i = 0; i < T; ++i) dl[i] = dlu[i] = ya[i] = make_float2(0.f, 0.f);
    float2 hout = make_float2(0.f, 0.f);
    float2 a[2][T], bb[2][T];
    float Cn[2][T];

    for (int k = -1; k <= NC + 1; ++k) {
        __syncthreads();
        if (k < 0 || k > NC) continue;
        const unsigned char *tb_cur = s_tile + (k % Cfg::kTiles) * Tl::kBytes + Tl::quad_off(qa);
        const unsigned char *tb_prev = s_tile + ((k + Cfg::kTiles - 1) % Cfg::kTiles) * Tl::kBytes + Tl::quad_off(qa);
        const float4 *st = s_stage + (k & 1) * 8 * NRT + tid, *yi = s_du + (k & 1) * 4 * NRT + tid;
        float4 *yo = s_yout + ((k & 1) ^ 1) * 4 * NRT + tid;
        float2 *sx_cur = s_x + ((k % 3) * 16 * NRP + rp) * 2 + (j >> 3), *sx_prev = s_x + (((k + 2) % 3) * 16 * NRP + rp) * 2 + (j >> 3);

#define MMU_S5_PREP(s_, an, bn, cn, WITH_CHAIN, ac, bc, cc)                                                                            \
        {                                                                                                                              \
            if ((s_) == j) {                                                                                                           \
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                                        \
                    const float4 v = st[q * NRT], w = st[(4 + q) * NRT];                                                               \
                    dl[2 * q] = make_float2(v.x, v.y), dl[2 * q + 1] = make_float2(v.z, v.w);                                          \
                    dlu[2 * q] = make_float2(w.x, w.y), dlu[2 * q + 1] = make_float2(w.z, w.w);                                        \
                }                                                                                                                      \
            }                                                                                                                          \
            const int n_ = ((s_) - j) & 15;                                                                                            \
            const unsigned char *rowB = ((s_) >= j ? tb_cur : tb_prev) + n_ * Tl::kRowBytes, *rowC = rowB + 16 * Tl::kRowBytes;        \
            const float2 A2 = s_A_rp[n_];                                                                                              \
            float Bn[T];                                                                                                               \
            {                                                                                                                          \
                const float4 b0 = *reinterpret_cast<const float4 *>(rowB), b1 = *reinterpret_cast<const float4 *>(rowB + 16);          \
                const float eb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};                                                  \
                order8<REV>(eb, Bn);                                                                                                   \
                const float4 c0 = *reinterpret_cast<const float4 *>(rowC), c1 = *reinterpret_cast<const float4 *>(rowC + 16);          \
                const float ec[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};                                                  \
                order8<REV>(ec, cn);                                                                                                   \
            }                                                                                                                          \
            _Pragma("unroll") for (int i = 0; i < T; ++i) {                                                                            \
                an[i] = (MMU_S5_WHATIF && (i & 1)) ? fmul2(dl[i], A2) : ex2(fmul2(dl[i], A2));                                         \
                bn[i] = fmul2(dlu[i], splat(Bn[i]));                                                                                   \
                if (WITH_CHAIN) {                                                                                                      \
                    h = ffma2(ac[i], h, bc[i]);                                                                                        \
                    ya[i] = ffma2(h, splat(cc[i]), ya[i]);                                                                             \
                }                                                                                                                      \
            }                                                                                                                          \
        }
        float2 h = make_float2(0.f, 0.f);
        MMU_S5_PREP(0, a[0], bb[0], Cn[0], false, a[0], bb[0], Cn[0])
        // one step: (transition) -> predecessor's state -> recurrence of step s_ interleaved with the operands of step s_ + 1.
        // The step loop is NOT unrolled beyond the cur / nxt pair: one copy of the code keeps delta, delta*u and y in fixed registers,
        // so the predicated transition loads land in them directly (the fully unrolled form split their live ranges and paid ~50
        // predicated moves per step), and the round fits the instruction cache.
#define MMU_S5_STEP(s_, cur, nxt, PREP)                                                                                                \
        {                                                                                                                              \
            if ((s_) == j) {   /* first step of my chunk k: the finished y of chunk k-1 goes out, D*u of chunk k comes in.  (Leaving  */  \
                               /* the accumulators running and letting the helper take differences of snapshots saves the reload:  */  \
                               /* 115 instead of 117 us at B16 D128 L4096, nothing at config 2 - not worth an error term that      */  \
                               /* grows with the number of chunks.)                                                                */  \
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                                        \
                    yo[q * NRT] = make_float4(ya[2 * q].x, ya[2 * q].y, ya[2 * q + 1].x, ya[2 * q + 1].y);                             \
                    const float4 v = yi[q * NRT];                                                                                      \
                    ya[2 * q] = make_float2(v.x, v.y), ya[2 * q + 1] = make_float2(v.z, v.w);                                          \
                }                                                                                                                      \
            }                                                                                                                          \
            h = make_float2(__shfl_sync(0xffffffffu, hout.x, src_lane), __shfl_sync(0xffffffffu, hout.y, src_lane));                   \
            if (PREP) {                                                                                                                \
                MMU_S5_PREP((s_) + 1, a[nxt], bb[nxt], Cn[nxt], true, a[cur], bb[cur], Cn[cur])                                        \
            } else {                                                                                                                   \
                _Pragma("unroll") for (int i = 0; i < T; ++i) {                                                                        \
                    h = ffma2(a[cur][i], h, bb[cur][i]);                                                                               \
                    ya[i] = ffma2(h, splat(Cn[cur][i]), ya[i]);                                                                        \
                }                                                                                                                      \
            }                                                                                                                          \
            hout = h;                                                                                                                  \
            if (xsave) ((s_) >= j ? sx_cur : sx_prev)[(((s_) - j) & 15) * NRP * 2] = h;                                                \
        }
#pragma unroll 1
        for (int s = 0; s < 16; s += 2) {
            MMU_S5_STEP(s, 0, 1, true)
            MMU_S5_STEP(s + 1, 1, 0, s + 2 < 16)
        }
#undef MMU_S5_STEP
#undef MMU_S5_PREP
    }
}

}  // namespace mmu

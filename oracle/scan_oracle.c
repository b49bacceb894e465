/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The CUDA product never calls it.
 *
 * Plain-C restatement of the reference's CPU algorithms for the Mamba-block hot path:
 *   - selective scan forward      : selective_scan_ref,
 *       requirements/Mamba/mamba/mamba_ssm/ops/selective_scan_interface.py:86-152
 *   - selective scan backward     : closed form of autograd through the same function
 *       (kernel line map selective_scan_bwd_kernel.cuh:186-453)
 *   - causal conv1d fwd / bwd     : causal_conv1d_ref,
 *       requirements/Mamba/causal-conv1d/causal_conv1d/causal_conv1d_interface.py:49-65
 *       (bwd = autograd of it; kernel line map causal_conv1d_bwd.cu:155-222)
 *   - scan-order index maps       : MMConv.two_row_columnwise_flatten_grad_safe / inverse
 *       (src/UM_Net/MMUNet.py:68-121), TFM flip / nslices interleave
 *       (requirements/mamba_simple.py:230,245-247,263)
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks every entry point against
 * golden vectors produced by running the unmodified reference functions in the build
 * container (oracle/gen_golden.py, outputs in tests/golden/).
 *
 * All arrays are dense row-major, float32 in/out; arithmetic is carried in double so the
 * oracle is at least as accurate as the fp32 torch reference it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); } /* F.softplus, threshold 20 */
static inline double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

/* ------------------------------------------------------------------------------------------
 * Selective scan forward.  selective_scan_interface.py:86-152
 *   u, delta, z, out : (B, D, L)        A : (D, N)       Bm, Cm : (B, G, N, L)   D, dbias : (D)
 *   z, Dv, dbias may be NULL.  last_state (B, D, N) may be NULL.
 * ------------------------------------------------------------------------------------------ */
int oracle_selective_scan_fwd(const float *u, const float *delta, const float *A, const float *Bm,
                              const float *Cm, const float *Dv, const float *z, const float *dbias,
                              int softplus, int64_t B, int64_t D, int64_t L, int64_t N, int64_t G,
                              float *out, float *last_state)
{
    if (G <= 0 || D % G) return -1;
    const int64_t H = D / G;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t d = 0; d < D; ++d) {
            double *h = (double *)calloc((size_t)N, sizeof(double));
            const int64_t g = d / H;
            const float *ur = u + (b * D + d) * L, *dr = delta + (b * D + d) * L;
            const float *zr = z ? z + (b * D + d) * L : NULL;
            float *orow = out + (b * D + d) * L;
            for (int64_t t = 0; t < L; ++t) {
                double dl = dr[t];
                if (dbias) dl += dbias[d];                 /* :104-105 */
                if (softplus) dl = softplus_d(dl);         /* :106-107 */
                double y = 0.0;
                for (int64_t n = 0; n < N; ++n) {
                    const double a = exp(dl * A[d * N + n]);                       /* :122 deltaA */
                    const double bu = dl * Bm[((b * G + g) * N + n) * L + t] * ur[t]; /* :126-130 */
                    h[n] = a * h[n] + bu;                                          /* :135 */
                    y += h[n] * Cm[((b * G + g) * N + n) * L + t];                 /* :139-142 */
                }
                if (Dv) y += (double)ur[t] * Dv[d];        /* :149 */
                if (zr) y *= zr[t] * sigmoid_d(zr[t]);     /* :150-151 */
                orow[t] = (float)y;
            }
            if (last_state) for (int64_t n = 0; n < N; ++n) last_state[(b * D + d) * N + n] = (float)h[n];
            free(h);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Selective scan backward (closed form).
 *   dout (B,D,L) -> du, ddelta (B,D,L); dA (D,N); dB, dC (B,G,N,L); dD, ddbias (D); dz (B,D,L)
 *   Optional outputs may be NULL.  dB/dC/dA/dD/ddbias are overwritten (not accumulated).
 * ------------------------------------------------------------------------------------------ */
int oracle_selective_scan_bwd(const float *u, const float *delta, const float *A, const float *Bm,
                              const float *Cm, const float *Dv, const float *z, const float *dbias,
                              const float *dout, int softplus, int64_t B, int64_t D, int64_t L,
                              int64_t N, int64_t G, float *du, float *ddelta, float *dA, float *dB,
                              float *dC, float *dD, float *dz, float *ddbias)
{
    if (G <= 0 || D % G) return -1;
    const int64_t H = D / G;
    double *dA_acc = (double *)calloc((size_t)(D * N), sizeof(double));
    double *dD_acc = (double *)calloc((size_t)D, sizeof(double));
    double *dbias_acc = (double *)calloc((size_t)D, sizeof(double));
    double *dB_acc = (double *)calloc((size_t)(B * G * N * L), sizeof(double));
    double *dC_acc = (double *)calloc((size_t)(B * G * N * L), sizeof(double));
    /* parallel over d so that dA/dD/dbias rows are private; dB/dC reduced under a critical section per (b,g) row block */
    #pragma omp parallel for schedule(dynamic)
    for (int64_t d = 0; d < D; ++d) {
        const int64_t g = d / H;
        double *h = (double *)malloc(sizeof(double) * (size_t)(L * N));   /* h[t][n] */
        double *dlv = (double *)malloc(sizeof(double) * (size_t)L);
        double *dBl = (double *)malloc(sizeof(double) * (size_t)(N * L));
        double *dCl = (double *)malloc(sizeof(double) * (size_t)(N * L));
        double *dh = (double *)malloc(sizeof(double) * (size_t)N);
        for (int64_t b = 0; b < B; ++b) {
            const float *ur = u + (b * D + d) * L, *dr = delta + (b * D + d) * L;
            const float *zr = z ? z + (b * D + d) * L : NULL;
            const float *gr = dout + (b * D + d) * L;
            const float *Bb = Bm + (b * G + g) * N * L, *Cb = Cm + (b * G + g) * N * L;
            /* forward recompute */
            for (int64_t t = 0; t < L; ++t) {
                double dl = dr[t];
                if (dbias) dl += dbias[d];
                if (softplus) dl = softplus_d(dl);
                dlv[t] = dl;
                for (int64_t n = 0; n < N; ++n) {
                    const double a = exp(dl * A[d * N + n]);
                    const double hp = t ? h[(t - 1) * N + n] : 0.0;
                    h[t * N + n] = a * hp + dl * Bb[n * L + t] * ur[t];
                }
            }
            for (int64_t n = 0; n < N; ++n) dh[n] = 0.0;
            for (int64_t t = L - 1; t >= 0; --t) {
                const double dl = dlv[t];
                double y = 0.0;
                for (int64_t n = 0; n < N; ++n) y += h[t * N + n] * Cb[n * L + t];
                if (Dv) y += (double)ur[t] * Dv[d];
                double dy = gr[t];
                if (zr) {
                    const double zz = zr[t], s = sigmoid_d(zz);
                    if (dz) dz[(b * D + d) * L + t] = (float)(gr[t] * y * s * (1.0 + zz * (1.0 - s)));
                    dy = gr[t] * zz * s;
                }
                double du_t = Dv ? dy * Dv[d] : 0.0;
                if (Dv) dD_acc[d] += dy * ur[t];
                double ddl = 0.0;
                for (int64_t n = 0; n < N; ++n) {
                    const double a_next = (t + 1 < L) ? exp(dlv[t + 1] * A[d * N + n]) : 0.0;
                    dh[n] = Cb[n * L + t] * dy + a_next * dh[n];
                    const double a = exp(dl * A[d * N + n]);
                    const double hp = t ? h[(t - 1) * N + n] : 0.0;
                    const double ahp = a * hp;
                    du_t += dh[n] * dl * Bb[n * L + t];
                    ddl += dh[n] * (Bb[n * L + t] * ur[t] + A[d * N + n] * ahp);
                    dA_acc[d * N + n] += dh[n] * dl * ahp;
                    dBl[n * L + t] = dh[n] * dl * ur[t];
                    dCl[n * L + t] = dy * h[t * N + n];
                }
                du[(b * D + d) * L + t] = (float)du_t;
                double ddraw = ddl;
                if (softplus) {
                    double x = dr[t] + (dbias ? dbias[d] : 0.0);
                    if (x <= 20.0) ddraw = ddl * sigmoid_d(x);
                }
                ddelta[(b * D + d) * L + t] = (float)ddraw;
                dbias_acc[d] += ddraw;
            }
            #pragma omp critical
            {
                double *pB = dB_acc + (b * G + g) * N * L, *pC = dC_acc + (b * G + g) * N * L;
                for (int64_t i = 0; i < N * L; ++i) { pB[i] += dBl[i]; pC[i] += dCl[i]; }
            }
        }
        free(h); free(dlv); free(dBl); free(dCl); free(dh);
    }
    if (dA) for (int64_t i = 0; i < D * N; ++i) dA[i] = (float)dA_acc[i];
    if (dD) for (int64_t i = 0; i < D; ++i) dD[i] = (float)dD_acc[i];
    if (ddbias) for (int64_t i = 0; i < D; ++i) ddbias[i] = (float)dbias_acc[i];
    if (dB) for (int64_t i = 0; i < B * G * N * L; ++i) dB[i] = (float)dB_acc[i];
    if (dC) for (int64_t i = 0; i < B * G * N * L; ++i) dC[i] = (float)dC_acc[i];
    free(dA_acc); free(dD_acc); free(dbias_acc); free(dB_acc); free(dC_acc);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Causal depthwise conv1d.  causal_conv1d_interface.py:49-65
 *   x, out (B, D, L);  w (D, W);  bias (D) or NULL;  silu flag.
 *   out[l] = act(bias + sum_k w[k] * x[l - (W-1-k)]), zero left padding.
 * ------------------------------------------------------------------------------------------ */
int oracle_causal_conv1d_fwd(const float *x, const float *w, const float *bias, int silu,
                             int64_t B, int64_t D, int64_t L, int64_t W, float *out)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t d = 0; d < D; ++d) {
            const float *xr = x + (b * D + d) * L;
            float *orow = out + (b * D + d) * L;
            for (int64_t l = 0; l < L; ++l) {
                double acc = bias ? bias[d] : 0.0;
                for (int64_t k = 0; k < W; ++k) {
                    const int64_t s = l - (W - 1 - k);
                    if (s >= 0) acc += (double)w[d * W + k] * xr[s];
                }
                if (silu) acc = acc * sigmoid_d(acc);
                orow[l] = (float)acc;
            }
        }
    return 0;
}

int oracle_causal_conv1d_bwd(const float *x, const float *w, const float *bias, const float *dout,
                             int silu, int64_t B, int64_t D, int64_t L, int64_t W, float *dx,
                             float *dw, float *dbias)
{
    #pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < D; ++d) {
        double dwacc[8] = {0}, dbacc = 0.0;
        double *dpre = (double *)malloc(sizeof(double) * (size_t)L);
        for (int64_t b = 0; b < B; ++b) {
            const float *xr = x + (b * D + d) * L, *gr = dout + (b * D + d) * L;
            for (int64_t l = 0; l < L; ++l) {
                double g = gr[l];
                if (silu) {
                    double pre = bias ? bias[d] : 0.0;
                    for (int64_t k = 0; k < W; ++k) {
                        const int64_t s = l - (W - 1 - k);
                        if (s >= 0) pre += (double)w[d * W + k] * xr[s];
                    }
                    const double sg = sigmoid_d(pre);
                    g = g * sg * (1.0 + pre * (1.0 - sg));
                }
                dpre[l] = g;
                dbacc += g;
                for (int64_t k = 0; k < W; ++k) {
                    const int64_t s = l - (W - 1 - k);
                    if (s >= 0) dwacc[k] += g * xr[s];
                }
            }
            for (int64_t l = 0; l < L; ++l) {
                double acc = 0.0;
                for (int64_t k = 0; k < W; ++k) {
                    const int64_t t = l + (W - 1 - k);
                    if (t < L) acc += (double)w[d * W + k] * dpre[t];
                }
                dx[(b * D + d) * L + l] = (float)acc;
            }
        }
        for (int64_t k = 0; k < W; ++k) dw[d * W + k] = (float)dwacc[k];
        if (dbias) dbias[d] = (float)dbacc;
        free(dpre);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Scan-order index maps.  idx[l] = source flat position (h*W + w) feeding token l.
 *   order 0: row-major (RCG, MMUNet.py:405)            idx[l] = l
 *   order 1: flip (mamba_simple.py:230)                idx[l] = L-1-l
 *   order 2: nslices interleave (mamba_simple.py:245-247)  idx[j*ns + s] = s*(L/ns) + j
 *   order 3: two-row column-interleave (MMUNet.py:68-93)   pairs of rows walked column by column,
 *            odd last row appended row-major.
 * ------------------------------------------------------------------------------------------ */
int oracle_scan_order_index(int order, int64_t H, int64_t W, int64_t ns, int64_t *idx)
{
    const int64_t L = H * W;
    switch (order) {
    case 0: for (int64_t l = 0; l < L; ++l) idx[l] = l; return 0;
    case 1: for (int64_t l = 0; l < L; ++l) idx[l] = L - 1 - l; return 0;
    case 2:
        if (ns <= 0 || L % ns) return -1;
        for (int64_t j = 0; j < L / ns; ++j)
            for (int64_t s = 0; s < ns; ++s) idx[j * ns + s] = s * (L / ns) + j;
        return 0;
    case 3: {
        const int64_t even_rows = (H / 2) * 2;
        int64_t l = 0;
        for (int64_t p = 0; p < even_rows / 2; ++p)        /* x_pair.permute(0,1,2,4,3): (pair, w, 2) */
            for (int64_t w = 0; w < W; ++w)
                for (int64_t r = 0; r < 2; ++r) idx[l++] = (2 * p + r) * W + w;
        for (int64_t w = 0; w < W && even_rows < H; ++w) idx[l++] = even_rows * W + w;   /* tail row */
        return 0;
    }
    default: return -2;
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the benchmark's CPU legs set the thread count explicitly. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

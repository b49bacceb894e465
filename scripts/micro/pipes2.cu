// Second round: LDS/STS caps with high ILP, MUFU || FFMA2 overlap, FMUL2 form.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
constexpr int ITERS = 1024, UNROLL = 16;

template <int W> __global__ void k_lds(float *out, long long *cyc, int stride4) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm + lane * stride4);
    float acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned a = addr + ((u & 7) * 512);
            if (W == 16) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                acc[u] += v.x + v.w;
            } else if (W == 8) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
                acc[u] += v.x;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                acc[u] += v;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += acc[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W> __global__ void k_sts(float *out, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm + lane * (W / 4));
    const float v = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned a = addr + ((u & 7) * 512);
            if (W == 16) asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(a), "f"(v) : "memory");
            else asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

// MODE 0: 8 MUFU only; 1: 8 FFMA2 only; 2: both interleaved (8 MUFU + 8 FFMA2); 3: 8 MUFU + 16 FFMA2; 4: 16 FFMA (scalar) + 8 MUFU
template <int MODE> __global__ void k_overlap(float *out, long long *cyc) {
    float x[8];
    float2 y[16];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = -0.001f * (threadIdx.x + u);
#pragma unroll
    for (int u = 0; u < 16; ++u) y[u] = make_float2(0.001f * (threadIdx.x + u), 0.5f);
    const float2 a = make_float2(0.999f, 1.001f), c = make_float2(0.001f, -0.001f);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE != 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
            if (MODE == 1 || MODE == 2) y[u] = __ffma2_rn(y[u], a, c);
            if (MODE == 3) {
                y[u] = __ffma2_rn(y[u], a, c);
                y[u + 8] = __ffma2_rn(y[u + 8], a, c);
            }
            if (MODE == 4) {
                y[u].x = fmaf(y[u].x, a.x, c.x);
                y[u].y = fmaf(y[u].y, a.y, c.y);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u];
#pragma unroll
    for (int u = 0; u < 16; ++u) s += y[u].x + y[u].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> void run(const char *name, F launch, int warps, double groups) {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    launch(warps, out, cyc);
    launch(warps, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += v;
    avg /= 148;
    printf("%-40s warps/SM %2d  cycles %9.0f  clk per warp-group/SM %.3f\n", name, warps, avg, avg / (groups * warps));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const double n = (double)ITERS * UNROLL;
    for (int warps : {8, 16, 32}) {
        run("LDS.128 distinct (16B stride)", [](int w, float *o, long long *c) { k_lds<16><<<148, 32 * w, 32768>>>(o, c, 4); }, warps, n);
        run("LDS.128 broadcast (same addr)", [](int w, float *o, long long *c) { k_lds<16><<<148, 32 * w, 32768>>>(o, c, 0); }, warps, n);
        run("LDS.64 distinct", [](int w, float *o, long long *c) { k_lds<8><<<148, 32 * w, 32768>>>(o, c, 2); }, warps, n);
        run("LDS.32 distinct", [](int w, float *o, long long *c) { k_lds<4><<<148, 32 * w, 32768>>>(o, c, 1); }, warps, n);
        run("STS.128 distinct", [](int w, float *o, long long *c) { k_sts<16><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("STS.32 distinct", [](int w, float *o, long long *c) { k_sts<4><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("8 MUFU", [](int w, float *o, long long *c) { k_overlap<0><<<148, 32 * w>>>(o, c); }, warps, ITERS);
        run("8 FFMA2", [](int w, float *o, long long *c) { k_overlap<1><<<148, 32 * w>>>(o, c); }, warps, ITERS);
        run("8 MUFU + 8 FFMA2", [](int w, float *o, long long *c) { k_overlap<2><<<148, 32 * w>>>(o, c); }, warps, ITERS);
        run("8 MUFU + 16 FFMA2", [](int w, float *o, long long *c) { k_overlap<3><<<148, 32 * w>>>(o, c); }, warps, ITERS);
        run("8 MUFU + 16 FFMA", [](int w, float *o, long long *c) { k_overlap<4><<<148, 32 * w>>>(o, c); }, warps, ITERS);
    }
    return 0;
}

// Microbenchmarks that pin the sm_100a per-SM pipe rates the scan kernels are budgeted against:
// LDS.128/64/32 passes (broadcast or not), SHFL, MUFU.EX2 (full / half-active warps), FFMA vs FFMA2, LDS+SHFL mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

constexpr int ITERS = 2048, UNROLL = 16;

template <int MODE> __global__ void k_lds(float *out, long long *cyc, int width) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int off;   // float index, multiple of 4
    if (MODE == 0) off = 0;                            // all lanes same address
    else if (MODE == 1) off = (lane >> 3) * 4;         // one address per quarter-warp
    else if (MODE == 2) off = (lane & 7) * 4;          // 8 distinct per quarter, same across quarters
    else off = lane * 4;                               // all distinct, conflict-free
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm + off);
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned a = addr + ((u & 7) * 512);
            if (width == 16) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                acc += v.x;
            } else if (width == 8) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
                acc += v.x;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                acc += v;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_shfl(float *out, long long *cyc) {
    float v = threadIdx.x, acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float r;
            asm volatile("shfl.sync.up.b32 %0, %1, 1, 0, 0xffffffff;" : "=f"(r) : "f"(v));
            acc += r;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_mix(float *out, long long *cyc) {   // alternating LDS.128 (distinct) and SHFL
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm + lane * 4);
    float v = threadIdx.x, acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL / 2; ++u) {
            float4 q;
            float r;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr + (u & 7) * 512));
            asm volatile("shfl.sync.up.b32 %0, %1, 1, 0, 0xffffffff;" : "=f"(r) : "f"(v));
            acc += q.x + r;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> __global__ void k_mufu(float *out, long long *cyc) {
    const int lane = threadIdx.x & 31;
    float x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) x[u] = -0.001f * (threadIdx.x + u);
    long long t0 = clock64();
    if (MODE == 0 || lane < 16) {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += x[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> __global__ void k_fma(float *out, long long *cyc) {   // 0: FFMA  1: FFMA2  2: FFMA2 with scalar-broadcast operand
    float2 x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) x[u] = make_float2(0.001f * (threadIdx.x + u), 0.5f);
    const float2 a = make_float2(0.999f, 1.001f), c = make_float2(0.001f, -0.001f);
    const float s = 0.9995f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE == 0) {
                x[u].x = fmaf(x[u].x, a.x, c.x);
            } else if (MODE == 1) {
                x[u] = __ffma2_rn(x[u], a, c);
            } else {
                x[u] = __ffma2_rn(x[u], make_float2(s, s), c);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += x[u].x + x[u].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_red(float *dst, long long *cyc, int vec, int nslots) {   // global fp32 atomics, spread addresses
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        const int slot = (tid * 4 + it * 4099 * 4) % nslots;
        float *p = dst + (slot & ~3);
        if (vec) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
        } else {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(1.0f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(1.0f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 3), "f"(1.0f) : "memory");
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double run(const char *name, void (*launch)(int, float *, long long *), int warps, double instr_per_thread) {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    launch(warps, out, cyc);
    launch(warps, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: %s\n", name, cudaGetErrorString(e));
        return 0;
    }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += v;
    avg /= 148;
    const double wi = instr_per_thread * warps;   // warp-instructions per SM
    printf("%-44s warps/SM %2d  cycles %9.0f  warp-instr/clk/SM %.3f  (clk per warp-instr %.2f)\n", name, warps, avg, wi / avg, avg / wi);
    cudaFree(out);
    cudaFree(cyc);
    return avg;
}

int main() {
    const double n = (double)ITERS * UNROLL;
    for (int warps : {4, 8, 16}) {
        run("LDS.128 all lanes same addr", [](int w, float *o, long long *c) { k_lds<0><<<148, 32 * w>>>(o, c, 16); }, warps, n);
        run("LDS.128 one addr per quarter-warp", [](int w, float *o, long long *c) { k_lds<1><<<148, 32 * w>>>(o, c, 16); }, warps, n);
        run("LDS.128 8 addrs, same in each quarter", [](int w, float *o, long long *c) { k_lds<2><<<148, 32 * w>>>(o, c, 16); }, warps, n);
        run("LDS.128 32 distinct", [](int w, float *o, long long *c) { k_lds<3><<<148, 32 * w>>>(o, c, 16); }, warps, n);
        run("LDS.64 all same", [](int w, float *o, long long *c) { k_lds<0><<<148, 32 * w>>>(o, c, 8); }, warps, n);
        run("LDS.64 8 addrs (x4 bcast)", [](int w, float *o, long long *c) { k_lds<2><<<148, 32 * w>>>(o, c, 8); }, warps, n);
        run("LDS.64 32 distinct (stride 16B)", [](int w, float *o, long long *c) { k_lds<3><<<148, 32 * w>>>(o, c, 8); }, warps, n);
        run("LDS.32 all same", [](int w, float *o, long long *c) { k_lds<0><<<148, 32 * w>>>(o, c, 4); }, warps, n);
        run("LDS.32 4 addrs (x8 bcast)", [](int w, float *o, long long *c) { k_lds<1><<<148, 32 * w>>>(o, c, 4); }, warps, n);
        run("SHFL.UP", [](int w, float *o, long long *c) { k_shfl<<<148, 32 * w>>>(o, c); }, warps, n);
        run("LDS.128 distinct + SHFL alternating", [](int w, float *o, long long *c) { k_mix<<<148, 32 * w>>>(o, c); }, warps, n);
        run("MUFU.EX2 full warps", [](int w, float *o, long long *c) { k_mufu<0><<<148, 32 * w>>>(o, c); }, warps, n);
        run("MUFU.EX2 half-active warps", [](int w, float *o, long long *c) { k_mufu<1><<<148, 32 * w>>>(o, c); }, warps, n);
        run("FFMA", [](int w, float *o, long long *c) { k_fma<0><<<148, 32 * w>>>(o, c); }, warps, n);
        run("FFMA2", [](int w, float *o, long long *c) { k_fma<1><<<148, 32 * w>>>(o, c); }, warps, n);
        run("FFMA2 scalar-broadcast operand", [](int w, float *o, long long *c) { k_fma<2><<<148, 32 * w>>>(o, c); }, warps, n);
    }
    // global atomics: 148*8 CTAs x 256 threads x 256 iters x 4 floats
    {
        float *dst;
        long long *cyc;
        const int nslots = 1 << 22;   // 16 MB of fp32 targets
        cudaMalloc(&dst, nslots * sizeof(float));
        cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
        cudaMemset(dst, 0, nslots * sizeof(float));
        for (int vec = 0; vec < 2; ++vec) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            k_red<<<148 * 8, 256>>>(dst, cyc, vec, nslots);
            cudaEventRecord(e0);
            k_red<<<148 * 8, 256>>>(dst, cyc, vec, nslots);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double floats = 148.0 * 8 * 256 * 256 * 4;
            printf("red.global.add %s: %s  %.3f ms  %.1f G float-adds/s  (%.1f GB/s payload)\n", vec ? "v4.f32" : "f32 x4", cudaGetErrorString(e), ms,
                   floats / ms * 1e-6, floats * 4 / ms * 1e-6);
        }
    }
    return 0;
}

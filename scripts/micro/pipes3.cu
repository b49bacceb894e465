// LDS/STS bandwidth with all vector components consumed and a loop-carried address (no CSE / narrowing).
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
constexpr int ITERS = 1024, UNROLL = 16;

// PAT 0: every lane distinct 16B-stride;  1: all lanes same address;  2: 8 distinct per quarter-warp, identical across quarters
template <int W, int PAT> __global__ void k_lds(float *out, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (i * 37) & 1023;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int l = PAT == 0 ? lane : PAT == 1 ? 0 : (lane & 7);
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    unsigned off = l * W;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned a = base + ((off + u * 1024) & 16383);
            if (W == 16) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            } else if (W == 8) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
                acc[0] += v.x; acc[1] += v.y;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                acc[u & 3] += v;
            }
        }
        off = (off + 512 + (it & 1) * 16384) & 16383;   // loop-carried
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

template <int W> __global__ void k_sts(float *out, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    unsigned off = lane * W;
    float v = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned a = base + ((off + u * 1024) & 16383);
            if (W == 16) asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(a), "f"(v) : "memory");
            else if (W == 8) asm volatile("st.shared.v2.f32 [%0], {%1,%1};" ::"r"(a), "f"(v) : "memory");
            else asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
        }
        off = (off + 512) & 16383;
        v += 1.f;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

template <typename F> void run(const char *name, F launch, int warps, double groups) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    launch(warps, out, cyc); launch(warps, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= 148;
    printf("%-40s warps/SM %2d  cycles %9.0f  clk per warp-instr per SM %.3f\n", name, warps, avg, avg / (groups * warps));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const double n = (double)ITERS * UNROLL;
    for (int warps : {8, 32}) {
        run("LDS.128 distinct", [](int w, float *o, long long *c) { k_lds<16, 0><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("LDS.128 all-same", [](int w, float *o, long long *c) { k_lds<16, 1><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("LDS.128 8 per quarter, x4 bcast", [](int w, float *o, long long *c) { k_lds<16, 2><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("LDS.64 distinct", [](int w, float *o, long long *c) { k_lds<8, 0><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("LDS.64 8 addrs x4 bcast", [](int w, float *o, long long *c) { k_lds<8, 2><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("LDS.32 distinct", [](int w, float *o, long long *c) { k_lds<4, 0><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("LDS.32 all-same", [](int w, float *o, long long *c) { k_lds<4, 1><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("STS.128 distinct", [](int w, float *o, long long *c) { k_sts<16><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("STS.64 distinct", [](int w, float *o, long long *c) { k_sts<8><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
        run("STS.32 distinct", [](int w, float *o, long long *c) { k_sts<4><<<148, 32 * w, 32768>>>(o, c); }, warps, n);
    }
    return 0;
}

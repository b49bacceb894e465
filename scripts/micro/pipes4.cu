// SHFL throughput with loop-carried operands; SHFL + LDS.128 sharing.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
constexpr int ITERS = 1024, UNROLL = 16;
template <int MODE> __global__ void k(float *out, long long *cyc) {   // 0: 16 SHFL   1: 8 SHFL + 8 LDS.128   2: 8 LDS.128
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (i * 37) & 1023;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm), off = lane * 16;
    float v[UNROLL], acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = threadIdx.x + u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE == 0 || (MODE == 1 && (u & 1))) {
                asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(v[u]));
            } else if (MODE == 1 || (MODE == 2 && (u & 1) == 0)) {
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(base + ((off + u * 1024) & 16383)));
                acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
            }
        }
        off = (off + 512) & 16383;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = acc[0] + acc[1] + acc[2] + acc[3];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += v[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> void run(const char *name, F launch, int warps) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    launch(warps, out, cyc); launch(warps, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (auto x : h) avg += x; avg /= 148;
    printf("%-28s warps/SM %2d  cycles %9.0f  clk per iteration-of-16 per warp per SM %.3f\n", name, warps, avg, avg / ((double)ITERS * warps));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int warps : {8, 32}) {
        run("16 SHFL", [](int w, float *o, long long *c) { k<0><<<148, 32 * w, 32768>>>(o, c); }, warps);
        run("8 SHFL + 8 LDS.128", [](int w, float *o, long long *c) { k<1><<<148, 32 * w, 32768>>>(o, c); }, warps);
        run("8 LDS.128", [](int w, float *o, long long *c) { k<2><<<148, 32 * w, 32768>>>(o, c); }, warps);
    }
}

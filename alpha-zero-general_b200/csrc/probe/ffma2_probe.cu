// ffma2_probe.cu -- throughput of packed fp32 FMA (fma.rn.f32x2 = FFMA2) against scalar FFMA on sm_100a.
// One CTA per SM; W warps per CTA; every thread runs N iterations of 16 independent accumulator updates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu ; run: ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
template <int MODE> __global__ void k(float* out, long long* cyc, int n, float w0, float x0) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = threadIdx.x * 0.001f + i;
    unsigned long long a2[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a2[i] = pk(acc[2 * i], acc[2 * i + 1]);
    float w = w0, x = x0;
    const unsigned long long w2 = pk(w, w * 1.0001f), x2 = pk(x, x);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = fmaf(w, acc[i], x);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(a2[i]) : "l"(w2), "l"(x2));
        }
    }
    const long long t1 = clock64();
    float s = 0;
    if (MODE == 0) { for (int i = 0; i < 32; i++) s += acc[i]; }
    else { for (int i = 0; i < 16; i++) { float a, b; upk(a2[i], a, b); s += a + b; } }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int n = 2000;
    for (int warps : {4, 8, 16}) {
        for (int mode = 0; mode < 2; mode++) {
            long long c = 0;
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, n, 0.999f, 0.001f); else k<1><<<148, warps * 32>>>(out, cyc, n, 0.999f, 0.001f);
                cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            }
            // 32 fp32 FMAs per thread per iteration in both modes
            printf("%2d warps/SM  %s: %lld cycles, %.2f FMA lanes per clock per SM\n", warps, mode ? "FFMA2" : "FFMA ", c, (double)n * 32 * warps * 32 / (double)c);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

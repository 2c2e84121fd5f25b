// shift_probe.cu -- does a K-major SWIZZLE_128B operand descriptor work when its start address is shifted by a number of ROWS that is
// not a multiple of 8 (i.e. not 1024-byte aligned)? The implicit-GEMM 3x3 convolution of the Santorini V89 kernel (net_v89_tc.cuh)
// reads its nine taps as nine row-shifted views of ONE activation buffer. Tested: shift s in 0..15 with the descriptor's base-offset
// field = 0 and = (address >> 7) & 7. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o shift_probe shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../umma.cuh"
using namespace azg::umma;

constexpr int ROWS = 160, K = 64, N = 64;
struct Args { const float* A; const float* B; float* D; int shift, use_base_offset; long long* cyc; int reps; };

__global__ void __launch_bounds__(128, 1) k_shift(Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_mma; __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t* sA = smem;                       // 2 atoms x ROWS x 128 B
    uint8_t* sB = smem + 2 * ROWS * 128;      // 2 atoms x N x 128 B
    if (t == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<128>(&tmem_base_s);
    for (int i = t; i < ROWS * K; i += 128) { const int r = i / K, k = i % K; *(float*)(sA + (k >> 5) * ROWS * 128 + sw128_off(r, k & 31)) = a.A[i]; }
    for (int i = t; i < N * K; i += 128) { const int r = i / K, k = i % K; *(float*)(sB + (k >> 5) * N * 128 + sw128_off(r, k & 31)) = a.B[i]; }
    fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_base_s;
    if (t == 0) {
        const uint32_t idesc = idesc_tf32(128, N);
        long long t0 = clock64();
        for (int rep = 0; rep < a.reps; rep++)
            for (int ks = 0; ks < K / 8; ks++) {
                const uint32_t aaddr = smem_u32(sA) + (ks >> 2) * ROWS * 128 + a.shift * 128 + (ks & 3) * 32;
                uint64_t da = desc_sw128(aaddr);
                if (a.use_base_offset) da |= (uint64_t)((aaddr >> 7) & 7u) << 49;
                const uint64_t db = desc_sw128(smem_u32(sB) + (ks >> 2) * N * 128 + (ks & 3) * 32);
                mma_tf32(tm, da, db, idesc, (rep | ks) != 0);
            }
        mma_commit(&bar_mma);
        mbar_wait(&bar_mma, 0);
        if (a.cyc) *a.cyc = clock64() - t0;
    }
    __syncthreads();
    mbar_wait(&bar_mma, 0); tc_fence_after();
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32]; tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + c, v); tmem_wait_ld();
        for (int j = 0; j < 32; j++) a.D[(size_t)(32 * warp + lane) * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}
static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
int main() {
    std::vector<float> A(ROWS * K), B(N * K);
    srand(7); for (auto& x : A) x = tf32_trunc((float)rand() / RAND_MAX * 2.f - 1.f); for (auto& x : B) x = tf32_trunc((float)rand() / RAND_MAX * 2.f - 1.f);
    float *dA, *dB, *dD; long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = 2 * ROWS * 128 + 2 * N * 128 + 1024;
    CK(cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int ubo = 0; ubo < 2; ubo++)
        for (int s = 0; s <= 16; s++) {
            Args a{dA, dB, dD, s, ubo, dC, 1};
            CK(cudaMemset(dD, 0, 128 * N * 4));
            k_shift<<<1, 128, smem>>>(a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift %d base_offset %d: KERNEL ERROR %s\n", s, ubo, cudaGetErrorString(e)); return 2; }
            std::vector<float> D(128 * N); CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double err = 0;
            for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
                double r = 0; for (int k = 0; k < K; k++) r += (double)A[(m + s) * K + k] * B[n * K + k];
                err = fmax(err, fabs(D[m * N + n] - r));
            }
            printf("shift %2d rows, base_offset field %s: max|D - expected| = %.3e  %s\n", s, ubo ? "(addr>>7)&7" : "0", err, err < 1e-4 ? "OK" : "MISMATCH");
        }
    { Args a{dA, dB, dD, 5, 0, dC, 64}; k_shift<<<1, 128, smem>>>(a); CK(cudaDeviceSynchronize()); long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
      printf("timing M=128 N=%d shifted by 5 rows: %.1f cycles per MMA\n", N, (double)c / (64 * 8)); }
    return 0;
}

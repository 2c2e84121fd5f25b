// common.cuh -- warp helpers, counter-based RNG, error plumbing shared by the engine's kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace azg {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int WARP = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- 64-bit mixing (splitmix64 finaliser) used by the board hash and the RNG key schedule --------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ---- Philox4x32-10: counter-based generator; key = (seed, stream), counter = (a, b) ---------------
struct Philox {
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t out[4];
    int have;
    __device__ __forceinline__ Philox(uint64_t seed, uint64_t stream, uint64_t ctr) {
        uint64_t k = mix64(seed ^ mix64(stream + 0x9E3779B97F4A7C15ULL));
        k0 = (uint32_t)k; k1 = (uint32_t)(k >> 32);
        c0 = (uint32_t)ctr; c1 = (uint32_t)(ctr >> 32); c2 = (uint32_t)stream; c3 = (uint32_t)(stream >> 32);
        have = 0;
    }
    __device__ __forceinline__ void round_(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t ka, uint32_t kb) {
        uint32_t hi0 = __umulhi(0xD2511F53u, a), lo0 = 0xD2511F53u * a;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c), lo1 = 0xCD9E8D57u * c;
        a = hi1 ^ b ^ ka; b = lo1; c = hi0 ^ d ^ kb; d = lo0;
    }
    __device__ __forceinline__ void refill() {
        uint32_t a = c0, b = c1, c = c2, d = c3, ka = k0, kb = k1;
#pragma unroll
        for (int i = 0; i < 10; i++) { round_(a, b, c, d, ka, kb); ka += 0x9E3779B9u; kb += 0xBB67AE85u; }
        out[0] = a; out[1] = b; out[2] = c; out[3] = d; have = 4;
        if (++c0 == 0) ++c1;
    }
    __device__ __forceinline__ uint32_t next() { if (!have) refill(); return out[--have]; }
    // uniform in [0,1) with 32 random bits
    __device__ __forceinline__ float uniformf() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
    __device__ __forceinline__ double uniform() {
        uint64_t x = ((uint64_t)next() << 21) ^ (uint64_t)(next() >> 11);
        return (double)(x & ((1ULL << 53) - 1)) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    // Marsaglia-Tsang gamma(alpha, 1)
    __device__ double gamma(double a) {
        double boost = 1.0;
        if (a < 1.0) { double u = uniform(); if (u < 1e-300) u = 1e-300; boost = pow(u, 1.0 / a); a += 1.0; }
        double d = a - 1.0 / 3.0, c = rsqrt(9.0 * d);
        for (int it = 0; it < 64; it++) {
            double x = normal(), v = 1.0 + c * x;
            if (v <= 0) continue;
            v = v * v * v;
            double u = uniform();
            if (u < 1.0 - 0.0331 * x * x * x * x || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v * boost;
        }
        return d * boost;
    }
};

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}


// Packed fp32 FMA (FFMA2): two IEEE fp32 FMAs per instruction, bit-identical to two fmaf. Same FMA-lane throughput as FFMA
// (csrc/probe/ffma2_probe.cu) but half the issue slots, which is what the register-tiled linears / 1x1 convolutions are bound by.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void fma2(uint64_t& acc, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
}  // namespace azg

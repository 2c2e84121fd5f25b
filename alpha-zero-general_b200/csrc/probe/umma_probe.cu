// umma_probe.cu -- standalone check of the tcgen05 / TMEM / bulk-copy primitives in ../umma.cuh against a CPU GEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../umma.cuh"
using namespace azg::umma;

struct Args {
    const float *Ahi, *Alo, *Bhi, *Blo;   // pre-swizzled images, [atoms][rows*32] floats
    const float *Arm, *Brm;               // row-major [rows][K]
    float* D;                             // [128][256] TMEM dump
    long long* cyc;
    int M, N, K, rowsA, rowsB, mode, reps;
};
// mode bits: 1 = operands from pre-swizzled images via cp.async.bulk (else threads write the swizzled layout from row-major)
//            2 = three-pass split (lo*hi + hi*lo + hi*hi), 4 = add 1.0 through a tcgen05.st / ld round trip, 8 = timing loop
__global__ void __launch_bounds__(128, 1) k_probe(Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int atoms = (a.K + 31) / 32;
    const uint32_t szA = (uint32_t)a.rowsA * 128u, szB = (uint32_t)a.rowsB * 128u;
    float* sAhi = (float*)smem; float* sAlo = (float*)(smem + atoms * szA);
    float* sBhi = (float*)(smem + 2 * atoms * szA); float* sBlo = (float*)(smem + 2 * atoms * szA + atoms * szB);
    if (t == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_base_s;
    {   // zero the 256 allocated columns
        uint32_t z[32];
        for (int j = 0; j < 32; j++) z[j] = 0;
        for (int c = 0; c < 256; c += 32) tmem_st32(tm + ((uint32_t)(32 * warp) << 16) + c, z);
        tmem_wait_st();
    }
    if (a.mode & 1) {
        if (t == 0) {
            mbar_expect_tx(&bar_load, 2 * atoms * (szA + szB));
            bulk_g2s(sAhi, a.Ahi, atoms * szA, &bar_load); bulk_g2s(sAlo, a.Alo, atoms * szA, &bar_load);
            bulk_g2s(sBhi, a.Bhi, atoms * szB, &bar_load); bulk_g2s(sBlo, a.Blo, atoms * szB, &bar_load);
        }
        mbar_wait(&bar_load, 0);
    } else {
        for (int i = t; i < (int)(2 * atoms * (szA + szB) / 4); i += 128) ((float*)smem)[i] = 0.f;
        __syncthreads();
        for (int i = t; i < a.M * a.K; i += 128) {
            int r = i / a.K, k = i % a.K; float hi, lo; split_tf32(a.Arm[i], hi, lo);
            uint32_t o = (k >> 5) * szA + sw128_off(r, k & 31);
            *(float*)((uint8_t*)sAhi + o) = (a.mode & 2) ? hi : a.Arm[i]; *(float*)((uint8_t*)sAlo + o) = lo;
        }
        for (int i = t; i < a.N * a.K; i += 128) {
            int r = i / a.K, k = i % a.K; float hi, lo; split_tf32(a.Brm[i], hi, lo);
            uint32_t o = (k >> 5) * szB + sw128_off(r, k & 31);
            *(float*)((uint8_t*)sBhi + o) = (a.mode & 2) ? hi : a.Brm[i]; *(float*)((uint8_t*)sBlo + o) = lo;
        }
        fence_async_smem();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (t == 0) {
        const uint32_t idesc = idesc_tf32(a.M, a.N);
        const uint64_t dAhi = desc_sw128(smem_u32(sAhi)), dAlo = desc_sw128(smem_u32(sAlo));
        const uint64_t dBhi = desc_sw128(smem_u32(sBhi)), dBlo = desc_sw128(smem_u32(sBlo));
        const int ksteps = (a.K + 7) / 8;
        long long t0 = clock64();
        const int reps = (a.mode & 8) ? a.reps : 1;
        bool acc = false;
        for (int rep = 0; rep < reps; rep++) {
            const int passes = (a.mode & 2) ? 3 : 1;
            for (int p = 0; p < passes; p++) {
                const uint64_t da = (passes == 3 && p == 0) ? dAlo : dAhi;
                const uint64_t db = (passes == 3 && p == 1) ? dBlo : dBhi;
                for (int ks = 0; ks < ksteps; ks++) {
                    const uint32_t offA = ((ks >> 2) * szA + (ks & 3) * 32) >> 4, offB = ((ks >> 2) * szB + (ks & 3) * 32) >> 4;
                    mma_tf32(tm, da + offA, db + offB, idesc, acc);
                    acc = true;
                }
            }
        }
        mma_commit(&bar_mma);
        mbar_wait(&bar_mma, 0);
        long long t1 = clock64();
        if (a.cyc) *a.cyc = t1 - t0;
    }
    __syncthreads();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        const uint32_t ta = tm + ((uint32_t)(32 * warp) << 16) + c;
        tmem_ld32(ta, v); tmem_wait_ld();
        if (a.mode & 4) {
            for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + 1.0f);
            tmem_st32(ta, v); tmem_wait_st();
            uint32_t u[8];
            for (int q = 0; q < 4; q++) { tmem_ld8(ta + 8 * q, u); tmem_wait_ld(); for (int j = 0; j < 8; j++) v[8 * q + j] = u[j]; }
        }
        for (int j = 0; j < 32; j++) a.D[(size_t)(32 * warp + lane) * 256 + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tm);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static std::vector<float> image(const std::vector<float>& rm, int rows, int K, int rowsPad, bool lo_part, bool split) {
    int atoms = (K + 31) / 32; std::vector<float> img((size_t)atoms * rowsPad * 32, 0.f);
    for (int r = 0; r < rows; r++) for (int k = 0; k < K; k++) {
        float x = rm[(size_t)r * K + k], hi = x, lo = 0.f; if (split) split_tf32(x, hi, lo);
        img[((size_t)(k >> 5) * rowsPad * 128 + sw128_off(r, k & 31)) / 4] = lo_part ? lo : hi;
    }
    return img;
}

static int run(const char* name, int M, int N, int K, int mode, bool exact_inputs, int nan_rows_from = -1, int reps = 1) {
    const int rowsA = 128, rowsB = (N + 7) / 8 * 8;
    std::vector<float> A((size_t)rowsA * K, 0.f), B((size_t)rowsB * K, 0.f);
    srand(1234 + M + N + K + mode);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (int i = 0; i < M * K; i++) A[i] = exact_inputs ? tf32_trunc(rnd()) : rnd();
    for (int i = 0; i < N * K; i++) B[i] = exact_inputs ? tf32_trunc(rnd()) : rnd();
    if (nan_rows_from >= 0) for (int r = nan_rows_from; r < rowsA; r++) for (int k = 0; k < K; k++) A[(size_t)r * K + k] = NAN;
    const bool split = mode & 2;
    auto Ahi = image(A, rowsA, K, rowsA, false, split), Alo = image(A, rowsA, K, rowsA, true, split);
    auto Bhi = image(B, rowsB, K, rowsB, false, split), Blo = image(B, rowsB, K, rowsB, true, split);
    float *dAhi, *dAlo, *dBhi, *dBlo, *dA, *dB, *dD; long long* dC;
    auto up = [](float** d, const std::vector<float>& h) { CK(cudaMalloc(d, h.size() * 4)); CK(cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); };
    up(&dAhi, Ahi); up(&dAlo, Alo); up(&dBhi, Bhi); up(&dBlo, Blo); up(&dA, A); up(&dB, B);
    CK(cudaMalloc(&dD, 128 * 256 * 4)); CK(cudaMemset(dD, 0, 128 * 256 * 4)); CK(cudaMalloc(&dC, 8));
    Args a{dAhi, dAlo, dBhi, dBlo, dA, dB, dD, dC, M, N, K, rowsA, rowsB, mode, reps};
    const int atoms = (K + 31) / 32; size_t smem = 2 * atoms * (rowsA + rowsB) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 128, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> D(128 * 256); long long cyc = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
    const int validM = nan_rows_from >= 0 ? nan_rows_from : M;
    std::vector<double> R((size_t)validM * N), Rt((size_t)validM * N), Rr((size_t)validM * N);
    for (int m = 0; m < validM; m++) for (int n = 0; n < N; n++) {
        double s = 0, st = 0, sr = 0;
        for (int k = 0; k < K; k++) {
            float x = A[(size_t)m * K + k], y = B[(size_t)n * K + k];
            s += (double)x * y; st += (double)tf32_trunc(x) * tf32_trunc(y); sr += (double)tf32_rn(x) * tf32_rn(y);
        }
        R[(size_t)m * N + n] = s; Rt[(size_t)m * N + n] = st; Rr[(size_t)m * N + n] = sr;
    }
    const double add = (mode & 4) ? 1.0 : 0.0;
    if (mode & 8) { printf("%-44s M=%d N=%d K=%d: %lld cycles for %d reps -> %.1f cycles per MMA\n", name, M, N, K, cyc, reps, (double)cyc / (reps * ((mode & 2) ? 3 : 1) * ((K + 7) / 8))); }
    else if (M == 128) {
        double e0 = 0, et = 0, er = 0, bias = 0, rel = 0, mag = 0;
        for (int m = 0; m < validM; m++) for (int n = 0; n < N; n++) {
            double d = D[(size_t)m * 256 + n] - add;
            e0 = fmax(e0, fabs(d - R[(size_t)m * N + n])); et = fmax(et, fabs(d - Rt[(size_t)m * N + n])); er = fmax(er, fabs(d - Rr[(size_t)m * N + n]));
            double ex = R[(size_t)m * N + n]; bias += (fabs(d) - fabs(ex)); rel += fabs(d - ex); mag += fabs(ex);
        }
        printf("    mean|err|/mean|D| = %.3e, mean(|D|-|exact|)/mean|D| = %.3e (negative = truncation toward zero)\n", rel / mag, bias / mag);
        printf("%-44s M=%d N=%d K=%d mode=%d: max|D-exact|=%.3e  |D-trunc_inputs|=%.3e  |D-rn_inputs|=%.3e  %s\n", name, M, N, K, mode, e0, et, er,
               (e0 < 1e-4 || et < 1e-4 || er < 1e-4) ? "OK" : "MISMATCH");
    } else {   // M = 64: find the lane each row landed on
        printf("%-44s M=%d N=%d K=%d: row -> lane map:", name, M, N, K);
        int prev = -2, start = -1;
        for (int m = 0; m < M; m++) {
            int found = -1;
            for (int l = 0; l < 128 && found < 0; l++) {
                double err = 0; for (int n = 0; n < N; n++) err = fmax(err, fabs(D[(size_t)l * 256 + n] - Rt[(size_t)m * N + n]));
                if (err < 1e-4) found = l;
            }
            if (found != prev + 1) { if (start >= 0) printf("..%d]", prev); printf(" [r%d->l%d", m, found); start = m; }
            prev = found;
        }
        printf("..%d]\n", prev);
    }
    cudaFree(dAhi); cudaFree(dAlo); cudaFree(dBhi); cudaFree(dBlo); cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return 0;
}

int main() {
    run("T1 bulk-loaded images, exact tf32 inputs", 128, 128, 64, 1, true);
    run("T1b same, K=56", 128, 128, 56, 1, true);
    run("T2 N=64", 128, 64, 64, 1, true);
    run("T2b N=176", 128, 176, 56, 1, true);
    run("T2c K=8 single MMA", 128, 128, 8, 1, true);
    run("T2d K=32", 128, 128, 32, 1, true);
    run("T2e K=96", 128, 64, 96, 1, true);
    run("T3 device-written swizzle, exact inputs", 128, 128, 64, 0, true);
    run("T4 raw fp32 inputs, single pass (rounding?)", 128, 128, 64, 0, false);
    run("T5 3xTF32 split, device-written", 128, 128, 56, 2, false);
    run("T5b 3xTF32 split, bulk images, K=96", 128, 64, 96, 3, false);
    run("T5c 3xTF32 split, K=8", 128, 128, 8, 3, false);
    run("T6 tcgen05.st/ld round trip (+1)", 128, 128, 64, 1 | 4, true);
    run("T7 NaN in A rows >= 56", 128, 128, 64, 1, true, 56);
    run("T8 M=64 lane map", 64, 64, 64, 1, true);
    run("T9 timing M=128 N=128", 128, 128, 64, 1 | 8, true, -1, 64);
    run("T9b timing M=128 N=64", 128, 64, 64, 1 | 8, true, -1, 64);
    run("T9c timing M=64 N=128", 64, 128, 64, 1 | 8, true, -1, 64);
    run("T9d timing M=128 N=176", 128, 176, 56, 1 | 8, true, -1, 64);
    run("T9e timing M=128 N=256", 128, 256, 64, 1 | 8, true, -1, 64);
    return 0;
}

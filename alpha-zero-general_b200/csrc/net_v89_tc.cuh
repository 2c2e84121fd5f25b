// net_v89_tc.cuh -- SantoriniNNet V89 eval-mode forward with the ten 64->64 3x3 convolutions of the trunk (99.5 % of its 18.5 MFLOP
// per leaf) on the 5th-generation tensor cores (tcgen05, sm_100a). Same function as k_v89_forward in net_v89.cuh
// (santorini/SantoriniNNet.py:70-84,194-217,273-279 behind GenericNNetWrapper.predict / predict_server, GenericNNetWrapper.py:94-157).
//
// A 3x3 convolution is an IMPLICIT GEMM over nine row-shifted views of one activation buffer:
//   D[position row r][cout] = sum over taps (ky, kx) and cin of  X[r + 6 (ky - 1) + (kx - 1)][cin] . W[ky][kx][cin][cout]
// Activations live in shared memory as the K-major, 128-byte-swizzled A operand X[row][cin] (cin = 64 = two K atoms). Rows are the
// positions of the 5x5 boards laid out with pitch 6 (one zero pad column) and one shared zero pad row between consecutive boards:
// leaf l, cell (y, x) is row 14 + 36 l + 6 y + x, so every tap is a plain ROW OFFSET of the operand's start address. The swizzle
// of the operand layout is a pure function of the shared-memory address, so a descriptor may start at any row (checked on the
// hardware for offsets 0..16, csrc/probe/shift_probe.cu, profiles/r02_shift_probe.txt). No im2col copy is ever made.
//   * one tile = 7 leaves = rows 14..269 = two M=128 accumulators in TMEM; 175 of the 256 rows are board cells (68 %)
//   * the parity bar is 1e-5 against the reference's fp32 forward, so every GEMM is error-compensated like the V80 kernel ("3xTF32"):
//     X is kept as two planes, H = the fp32 value itself (the tensor core truncates its operands to TF32) and L = x - trunc_tf32(x);
//     W as W_hi = rn_tf32(w) and W_lo = w - W_hi stacked along N:  D[:, 0:64] = H W_hi,  D[:, 64:128] = H W_lo + L W
//     (one N=128 TF32 MMA per 8-wide k-step; the small L W term runs as kind::f16 MMAs with FP16 operands, K = 16 per instruction:
//     |L| < 2^-10 |x| keeps it to 2^-21 of the main product, and a TF32 MMA with N = 64 costs the same ~76 cycles as N = 128)
//   * the accumulator lives in TMEM, so a layer's output overwrites its input IN PLACE in the epilogue (bias, residual, ReLU,
//     zero at the pad rows, split into H / L); the block input (the residual) waits in a per-CTA scratch in global memory (L2)
//   * the tensor core adds into its fp32 accumulator with round-toward-zero (profiles/r01_umma_probe.txt: -1.9e-8 relative per
//     accumulation step); over the 72 k-steps of a K = 576 convolution and ten layers that bias alone reaches 3e-5 on the value
//     output. The chain is therefore cut: taps 0-4 and taps 5-8 go to two separate accumulators (all 512 TMEM columns are used:
//     2 M tiles x 2 x 128), and the small L W_hi products go to the W_lo half of the columns, so the large H W_hi sums see 40 and 32
//     truncating additions instead of 144; the epilogue adds the four parts in fp32 round-to-nearest
//   * weights stream as pre-swizzled images, one 40 KB unit per tap (TF32 W_hi | W_lo + an FP16 copy of W), through a two-slot cp.async.bulk / mbarrier ring fed by a
//     producer thread that runs ahead of the MMA-issuing thread, across layer and tile boundaries
//   * measured (profiles/r02_v89tc_phases.txt): a convolution's 288 MMAs take 21.8 k cycles = 92 B/clk of operand fetch + 13 B/clk of
//     weight writes, which is the shared-memory rate the tensor core reaches in isolation (105-122 B/clk, csrc/probe); the MMA
//     issuer waits for weights 6 % of the time. The kernel is bound by shared-memory operand bandwidth, not by L2 or issue.
//   * first layer (2 -> 64 channels) and the heads (1x1 convolutions + small Linears) stay on the CUDA cores (0.5 % of the FLOPs)
#pragma once
#include "net_v89.cuh"
#include "net_v80_tc.cuh"   // tc_rn_tf32
#include "umma.cuh"
#include <cuda_fp16.h>

namespace azg {

constexpr int T89_CTHREADS = 256;         // 8 compute warps: TMEM lane quarter = warp & 3, M tile = warp >> 2
constexpr int T89_THREADS = 288;          // + one producer warp (weight ring); it only meets the others at CTA-wide barriers outside the tile loop
constexpr int T89_TB = 7;                 // leaves per tile
constexpr int T89_ROW0 = 14;              // first output row (leaf 0, cell (0, 0)); taps reach 7 rows either side
constexpr int T89_ROWS = 280;             // rows of an activation plane (35 groups of 8)
constexpr int T89_PLANE = T89_ROWS * 128; // bytes of one (plane, K atom)
constexpr int T89_ACT = 0;                // [H atom0 | H atom1 | L (fp16: 64 channels = one 128-byte atom)]
constexpr int T89_WRING = 3 * T89_PLANE;  // 107520: two slots x 40 KB; one unit = one tap: [K atom][128 rows: W_hi 0-63, W_lo 64-127][128 B] tf32, then [64 rows][128 B] W as fp16
constexpr int T89_UNIT_TF32 = 32768;      // (four 16 KB slots, one per (tap, K atom), were measured too: 25.1 k cycles per convolution instead of 21.8 k)
constexpr int T89_UNIT_BYTES = T89_UNIT_TF32 + 8192;
constexpr int T89_UNIT_FLOATS = T89_UNIT_BYTES / 4;
constexpr int T89_NSLOT = 2;
constexpr int T89_MISC = T89_WRING + T89_NSLOT * T89_UNIT_BYTES;   // biases [11][64], conv0 weights [2][9][64], head 1x1 weights [64][4], input planes, head scratch
constexpr int T89_MISC_FLOATS = 11 * 64 + 2 * 9 * 64 + 64 * 4 + 2 * T89_TB * 49 + T89_TB * (50 + 28 + 64 + V89_AP);
constexpr int T89_SMEM = T89_MISC + T89_MISC_FLOATS * 4 + 1024;
constexpr int T89_ACC = 0;                // TMEM columns: M tile m, tap group a (taps 0-4 / 5-8): [(2m + a) * 128, +128) = [H W_hi | H W_lo + L W_hi]
constexpr int T89_RES_FLOATS = 256 * 64;  // per-CTA residual scratch in global memory: [channel quad 0..15][row - 14][4] (a warp's 32 rows of one quad = 512 contiguous bytes)

struct V89TCImg { int conv[10]; int total; };                     // float offsets of the per-conv images (9 taps x T89_UNIT_FLOATS)
inline V89TCImg v89tc_layout() { V89TCImg I; int o = 0; for (int i = 0; i < 10; i++) { I.conv[i] = o; o += 9 * T89_UNIT_FLOATS; } I.total = o; return I; }
inline uint16_t v89_f32_to_f16(float x) {                          // round to nearest even, host side (weights are far inside the fp16 range)
    uint32_t u; memcpy(&u, &x, 4);
    const uint32_t sign = (u >> 16) & 0x8000u; const int e = (int)((u >> 23) & 0xFF) - 127 + 15; uint32_t m = u & 0x7FFFFFu;
    if (e >= 31) return (uint16_t)(sign | 0x7BFFu);                // clamp (not reached)
    if (e <= 0) {                                                  // subnormal or zero
        if (e < -10) return (uint16_t)sign;
        m |= 0x800000u; const int sh = 14 - e; uint32_t r = m >> sh; const uint32_t rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        if (rem > half || (rem == half && (r & 1u))) r++;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((uint32_t)e << 10) | (m >> 13); const uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
    return (uint16_t)(sign | r);
}
// Host: operand images of trunk conv i (1..10 of the prepared blob: [cin][tap][cout], BN folded) -> [tap][atom][row][k] swizzled
inline void v89tc_prepare(const float* blob, const V89Layout& L, const V89TCImg& I, float* img) {
    using umma::sw128_off;
    for (int i = 0; i < I.total; i++) img[i] = 0.f;
    for (int ci = 1; ci < V89_NCONV; ci++)
        for (int c = 0; c < 64; c++) for (int t = 0; t < 9; t++) for (int o = 0; o < 64; o++) {
            const float w = blob[L.conv[ci] + (c * 9 + t) * 64 + o], hi = tc_rn_tf32(w), lo = w - hi;
            const size_t unit = (size_t)I.conv[ci - 1] + (size_t)t * T89_UNIT_FLOATS, base = unit + (size_t)(c >> 5) * 4096;
            img[base + sw128_off(o, c & 31) / 4] = hi; img[base + sw128_off(64 + o, c & 31) / 4] = lo;
            // fp16 copy of w for the L . W term: B operand [cout row][cin], one 128-byte atom, same 16-byte-chunk swizzle
            uint16_t* h16 = reinterpret_cast<uint16_t*>(img + unit + T89_UNIT_TF32 / 4);
            h16[((o >> 3) * 1024 + (o & 7) * 128 + ((((c >> 3) ^ (o & 7)) & 7) << 4) + (c & 7) * 2) / 2] = v89_f32_to_f16(w);
        }
}

namespace t89 {
using namespace umma;
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }      // barrier of the 8 compute warps
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// row -> (leaf, y, x) of the pitch-6 layout; false for pad rows / pad columns / rows past the last leaf
__device__ __forceinline__ bool row_cell(int row, int& l, int& y, int& x) {
    const int r = row - T89_ROW0; l = r / 36; const int rem = r - 36 * l; y = rem / 6; x = rem - 6 * y;
    return r >= 0 && l < T89_TB && y < 5 && x < 5;
}
// y[16] (channels c0..c0+15 of one row) -> the operand planes: H = y (fp32; the tensor core truncates it to TF32) and
// L = y - trunc_tf32(y) as FP16 (|L| < 2^-10 |y|: its 11-bit mantissa keeps the product L . W to 2^-21 of y . w)
__device__ __forceinline__ void store_row16(uint8_t* act, int row, int c0, const float (&y)[16]) {
    uint8_t* base = act + (c0 >> 5) * T89_PLANE;
    uint32_t lp[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t o = sw128_off(row, (c0 & 31) + 4 * j);
        float lo[4];
#pragma unroll
        for (int i = 0; i < 4; i++) lo[i] = __fsub_rn(y[4 * j + i], __uint_as_float(__float_as_uint(y[4 * j + i]) & 0xFFFFE000u));
        *reinterpret_cast<float4*>(base + o) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        const __half2 a = __floats2half2_rn(lo[0], lo[1]), b = __floats2half2_rn(lo[2], lo[3]);
        lp[2 * j] = *reinterpret_cast<const uint32_t*>(&a); lp[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&b);
    }
    uint8_t* lrow = act + 2 * T89_PLANE + (row >> 3) * 1024 + (row & 7) * 128;        // 16 channels = two 16-byte chunks of the row's 128 bytes
#pragma unroll
    for (int h = 0; h < 2; h++)
        *reinterpret_cast<uint4*>(lrow + ((((c0 >> 3) + h) ^ (row & 7)) << 4)) = make_uint4(lp[4 * h], lp[4 * h + 1], lp[4 * h + 2], lp[4 * h + 3]);
}
}  // namespace t89

// First-layer weights [cin 2][tap 9][cout 64] and bias as a kernel parameter: with the loops unrolled every weight is a constant-bank
// operand of its FFMA. (Read from shared memory, the same weight for all lanes is a broadcast 128-bit load = one wavefront per
// quarter-warp, 17 wavefronts per 16 FMAs: the layer was bound by that.)
struct V89First { float w[2 * 9 * 64]; float b[64]; float wh[64 * 3]; };   // wh: the 1x1 head convolutions [cin][policy 0, policy 1, value]

__global__ void __launch_bounds__(T89_THREADS, 1)
k_v89_tc(const float* __restrict__ P, const float* __restrict__ IMG, float* __restrict__ RES, const __grid_constant__ V89Layout L, const __grid_constant__ V89TCImg I,
         const __grid_constant__ V89First F0, const int* count_ptr, const int* list, const int8_t* boards, int bstride, const uint32_t* masks, float* pi_out, float* v_out, int n_max, long long* prof) {
    using namespace t89;
    constexpr int TB = T89_TB, A = V89_A, MW = 6;
    extern __shared__ uint8_t smem_raw[];
    // 1024-aligned base as symbol + offset: the pointer stays in the shared state space (LDS / STS). Rounding the generic address instead
    // turns every access below into a generic LD.E / ST.E.
    uint8_t* sm = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_full[T89_NSLOT], bar_empty[T89_NSLOT], bar_acc;
    __shared__ uint32_t tmem_s;
    __shared__ int slot_of[TB];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, q = warp & 3, mt = warp >> 2;
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int ntiles = (count + TB - 1) / TB;
    if ((int)blockIdx.x >= ntiles) return;
    uint8_t* ACT = sm + T89_ACT; uint8_t* WR = sm + T89_WRING;
    float* BIAS = reinterpret_cast<float*>(sm + T89_MISC); float* W0 = BIAS + 11 * 64; float* WH = W0 + 2 * 9 * 64; float* IN = WH + 64 * 4;     // IN [2][TB][49]
    float* PF = IN + 2 * TB * 49; float* VF = PF + TB * 50; float* VH = VF + TB * 28; float* LG = VH + TB * 64;
    for (int i = t; i < 11 * 64; i += T89_THREADS) BIAS[i] = __ldg(P + L.cbias[i >> 6] + (i & 63));
    for (int i = t; i < 3 * T89_PLANE / 16; i += T89_THREADS) reinterpret_cast<uint4*>(ACT)[i] = make_uint4(0, 0, 0, 0);   // pad rows stay zero for good
    if (t == 0) { for (int i = 0; i < T89_NSLOT; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); } mbar_init(&bar_acc, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_s);
    fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_s;
    const uint32_t tl = tm + ((uint32_t)(32 * q) << 16);          // this warp's TMEM lane quarter
    const uint32_t acc = tl + T89_ACC + 256 * mt;
    const int row = T89_ROW0 + 128 * mt + 32 * q + lane;          // the row this thread owns in every epilogue
    float4* res = reinterpret_cast<float4*>(RES + (size_t)blockIdx.x * T89_RES_FLOATS) + (128 * mt + 32 * q + lane);   // its residual row: quad c4 at res[c4 * 256]
    int rl, ry, rx; const bool cell = row_cell(row, rl, ry, rx);
    const uint32_t act_a = smem_u32(ACT), wr_a = smem_u32(WR);
    constexpr uint32_t ID128 = idesc_tf32(128, 128), IDH64 = idesc_f16(128, 64);
    uint32_t g_mma = 0, g_load = 0, n_acc = 0;                    // global tap counters of the issuer / the producer, accumulator phases
    int prof_i = 0; long long wait_cyc = 0;
#define T89_STAMP() do { if (prof && t == 0 && blockIdx.x == 0 && prof_i < 60) prof[prof_i++] = clock64(); } while (0)

    if (warp == 8) {
        // ---- weight producer: one thread streams the 90 tap images of every tile through the two-slot ring, as far ahead as the ring
        //      allows (it only ever waits for the MMAs that read the slot it refills), across layer and tile boundaries
        if (lane == 0) {
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int k = 0; k < 90; k++, g_load++) {           // 10 convolutions x 9 taps, in the order the images are stored
                    const uint32_t s = g_load % T89_NSLOT, use = g_load / T89_NSLOT;
                    if (use >= 1) mbar_wait(&bar_empty[s], (use & 1u) ^ 1u);
                    mbar_expect_tx(&bar_full[s], T89_UNIT_BYTES);
                    bulk_g2s(WR + s * T89_UNIT_BYTES, IMG + (size_t)k * T89_UNIT_FLOATS, T89_UNIT_BYTES, &bar_full[s]);
                }
        }
        __syncwarp();
    } else
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tile0 = tile * TB;
        T89_STAMP();   /* tile start */
        if (t < TB) { const int j = tile0 + t; slot_of[t] = j < count ? (list ? list[j] : j) : -1; }
        for (int i = t; i < 2 * TB * 49; i += T89_CTHREADS) IN[i] = 0.f;
        csync();
        for (int k = t; k < TB * 50; k += T89_CTHREADS) {                              // input planes: workers, levels (7x7 zero-padded per leaf)
            const int l = k / 50, r = k - l * 50, c = r / 25, pos = r - c * 25, slot = slot_of[l];
            if (slot >= 0) IN[(c * TB + l) * 49 + (pos / 5 + 1) * 7 + pos % 5 + 1] = (float)boards[(size_t)slot * bstride + pos * 3 + c];
        }
        csync();
        {   // ---- first layer: Conv3x3(2 -> 64) + BN + ReLU on the CUDA cores; thread = row, 64 outputs in four 16-channel passes
            float vin[18];                                       // the 2 x 9 inputs of this row's cell (zero-padded planes)
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int tp = 0; tp < 9; tp++) vin[c * 9 + tp] = cell ? IN[(c * TB + rl) * 49 + (ry + tp / 3) * 7 + rx + tp % 3] : 0.f;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; j++) y[j] = 0.f;
                if (cell) {
#pragma unroll
                    for (int j = 0; j < 16; j++) y[j] = F0.b[c0 + j];
#pragma unroll
                    for (int k = 0; k < 18; k++)
#pragma unroll
                        for (int j = 0; j < 16; j++) y[j] = fmaf(F0.w[k * 64 + c0 + j], vin[k], y[j]);
#pragma unroll
                    for (int j = 0; j < 16; j++) y[j] = fmaxf(y[j], 0.f);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) __stcg(res + ((c0 >> 2) + j) * 256, make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]));   // block 0's residual
                store_row16(ACT, row, c0, y);
            }
        }
        fence_async_smem(); tc_fence_before(); csync(); tc_fence_after();
        T89_STAMP();   /* first layer done */
#pragma unroll 1
        for (int cv = 0; cv < 10; cv++) {
            // ---- implicit GEMM: 9 taps x 2 M tiles x 8 k-steps x (N=128 + N=64), issued by one thread
            if (t == 0) {
#pragma unroll 1
                for (int tap = 0; tap < 9; tap++, g_mma++) {
                    const uint32_t s = g_mma % T89_NSLOT;
                    const long long w0_ = prof ? clock64() : 0;
                    mbar_wait(&bar_full[s], (g_mma / T89_NSLOT) & 1u); tc_fence_after();
                    if (prof) wait_cyc += clock64() - w0_;
                    const int shift = 6 * (tap / 3 - 1) + (tap % 3 - 1);
                    const uint32_t wb = wr_a + s * T89_UNIT_BYTES;
                    const bool accum = !(tap == 0 || tap == 5);
#pragma unroll 1
                    for (int m = 0; m < 2; m++) {
                        const uint32_t arow = act_a + (uint32_t)(T89_ROW0 + 128 * m + shift) * 128u;
                        const uint32_t dcol = tm + T89_ACC + 256 * m + (tap >= 5 ? 128 : 0);
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const uint32_t ao = (ks >> 2) * T89_PLANE + (ks & 3) * 32;
                            mma_tf32(dcol, desc_sw128(arow + ao), desc_sw128(wb + (ks >> 2) * 16384 + (ks & 3) * 32), ID128, accum || ks != 0);   // H . (W_hi | W_lo)
                        }
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)                                                                       // L . W in FP16, K = 16 per MMA, next to H . W_lo
                            mma_f16(dcol + 64, desc_sw128(arow + 2 * T89_PLANE + ks * 32), desc_sw128(wb + T89_UNIT_TF32 + ks * 32), IDH64, true);
                    }
                    mma_commit(&bar_empty[s]);                     // slot free once these MMAs have read it
                }
                mma_commit(&bar_acc);
                mbar_wait(&bar_acc, n_acc & 1u);                  // ONLY the issuer polls the mbarrier: 255 threads spinning on a shared-memory
            }                                                     // barrier would compete with the tensor core's operand fetches for shared-memory bandwidth;
            n_acc++;                                              // everybody else sleeps in the hardware barrier below
            csync();
            tc_fence_after();
            T89_STAMP();   /* conv cv: MMAs done */
            // ---- epilogue, in place: bias (+ residual) + ReLU, zero at pad rows; conv2 of a block parks its output as the next residual
            const bool second = cv & 1;
            const float* bias = BIAS + (1 + cv) * 64;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t d0[16], d1[16], s0[16], s1[16];
                ld16(acc + c0, d0); ld16(acc + 128 + c0, d1); ld16(acc + 64 + c0, s0); ld16(acc + 192 + c0, s1);
                float xr[16];
                if (second) {
#pragma unroll
                    for (int j = 0; j < 4; j++) { const float4 r4 = __ldcg(res + ((c0 >> 2) + j) * 256); xr[4 * j] = r4.x; xr[4 * j + 1] = r4.y; xr[4 * j + 2] = r4.z; xr[4 * j + 3] = r4.w; }
                }
                tmem_wait_ld();
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    float v = (__uint_as_float(d0[j]) + __uint_as_float(d1[j])) + (__uint_as_float(s0[j]) + __uint_as_float(s1[j])) + bias[c0 + j];
                    if (second) v += xr[j];
                    y[j] = cell ? fmaxf(v, 0.f) : 0.f;
                }
                if (second) {
#pragma unroll
                    for (int j = 0; j < 4; j++) __stcg(res + ((c0 >> 2) + j) * 256, make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]));
                }
                store_row16(ACT, row, c0, y);
            }
            fence_async_smem(); tc_fence_before(); csync(); tc_fence_after();
            T89_STAMP();   /* conv cv: epilogue done */
        }
        // ---- heads: 1x1 convolutions (+BN, ReLU) to 2 + 1 planes from the trunk output (this thread's own residual row), flattened channel-major
        {
            float a0 = __ldg(P + L.pi_b), a1 = __ldg(P + L.pi_b + 1), a2 = __ldg(P + L.v_b);
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float xr[16];
#pragma unroll
                for (int j = 0; j < 4; j++) { const float4 r4 = __ldcg(res + ((c0 >> 2) + j) * 256); xr[4 * j] = r4.x; xr[4 * j + 1] = r4.y; xr[4 * j + 2] = r4.z; xr[4 * j + 3] = r4.w; }
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float x = xr[j];
                    a0 = fmaf(F0.wh[3 * (c0 + j)], x, a0); a1 = fmaf(F0.wh[3 * (c0 + j) + 1], x, a1); a2 = fmaf(F0.wh[3 * (c0 + j) + 2], x, a2);
                }
            }
            if (cell) { const int pos = ry * 5 + rx; PF[rl * 50 + pos] = fmaxf(a0, 0.f); PF[rl * 50 + 25 + pos] = fmaxf(a1, 0.f); VF[rl * 28 + pos] = fmaxf(a2, 0.f); }
        }
        tc_fence_before(); csync();
        for (int k = t; k < TB * (V89_AP / 4); k += T89_CTHREADS) {                    // policy Linear(50 -> 162)
            const int l = k / (V89_AP / 4), og = k - l * (V89_AP / 4);
            float4 a = __ldg(reinterpret_cast<const float4*>(P + L.pi_fcb + 4 * og));
            for (int i = 0; i < 50; i++) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(P + L.pi_fc + i * V89_AP + 4 * og)); const float x = PF[l * 50 + i];
                a.x = fmaf(w.x, x, a.x); a.y = fmaf(w.y, x, a.y); a.z = fmaf(w.z, x, a.z); a.w = fmaf(w.w, x, a.w);
            }
            *reinterpret_cast<float4*>(LG + l * V89_AP + 4 * og) = a;
        }
        for (int k = t; k < TB * 64; k += T89_CTHREADS) {                              // value Linear(25 -> 64) + ReLU
            const int l = k >> 6, j = k & 63;
            float a = __ldg(P + L.v_fc1b + j);
            for (int i = 0; i < 25; i++) a = fmaf(__ldg(P + L.v_fc1 + i * 64 + j), VF[l * 28 + i], a);
            VH[l * 64 + j] = fmaxf(a, 0.f);
        }
        csync();
        if (warp < TB) {   // masked softmax: where(valid, logits, -1e8) -> log_softmax -> exp (SantoriniNNet.py:279; GenericNNetWrapper.py:119)
            const int l = warp, slot = slot_of[l];
            if (slot >= 0) {
                float lg[MW]; float mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < MW; k++) {
                    const int a = lane + 32 * k;
                    const bool valid = a < A && (masks[(size_t)slot * MW + k] >> lane & 1);
                    lg[k] = a < A ? (valid ? LG[l * V89_AP + a] : -1e8f) : -INFINITY;
                    mx = fmaxf(mx, lg[k]);
                }
                mx = warp_max_f32(mx);
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < MW; k++) sum += expf(lg[k] - mx);
                sum = warp_sum_f32(sum);
                const float lse = logf(sum);
#pragma unroll
                for (int k = 0; k < MW; k++) { const int a = lane + 32 * k; if (a < A) pi_out[(size_t)slot * A + a] = expf(lg[k] - mx - lse); }
            }
        } else if (lane < TB * 2) {                                                    // warp 7: value Linear(64 -> 2), tanh
            const int l = lane >> 1, o = lane & 1, slot = slot_of[l];
            float a = __ldg(P + L.v_fc2b + o);
            for (int j = 0; j < 64; j++) a = fmaf(__ldg(P + L.v_fc2 + o * 64 + j), VH[l * 64 + j], a);
            if (slot >= 0) v_out[(size_t)slot * 2 + o] = tanhf(a);
        }
        csync();                                                                       // tile done: IN / PF / LG may be reused
        T89_STAMP();   /* tile end */
        if (prof && t == 0 && blockIdx.x == 0 && tile == 0) prof[63] = wait_cyc;      // cycles the MMA issuer spent waiting for weights in the first tile
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

}  // namespace azg

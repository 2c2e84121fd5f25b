// net_v80_tc.cuh -- SplendorNNet V80 eval-mode forward on the 5th-generation tensor cores (tcgen05, sm_100a).
// Same function as k_v80_forward in net_v80.cuh (splendor/SplendorNNet.py:149-204,259-280,397-404,440 behind
// GenericNNetWrapper.predict / predict_server, GenericNNetWrapper.py:94-157); the seven token-mixing Linear layers
// (first_layer, 3 x expand 56->168, 3 x project 168->56 = 83 % of the FLOPs) run as tcgen05.mma kind::tf32 with
// fp32 accumulators in TMEM.  The parity bar is 1e-5 on pi and v against the reference's fp32 forward, which one
// TF32 pass cannot meet (2e-4 .. 1e-3 measured), so every GEMM is error-compensated: x = hi + lo with hi = x rounded
// to TF32 and lo = x - hi (exact), and  D = W_lo X_hi + W_hi X_lo + W_hi X_hi  ("3xTF32", max error 2e-6 on the golden
// vectors incl. the measured round-toward-zero accumulation of the tensor core, profiles/r01_umma_probe.txt). In the expand
// GEMMs the two small terms (2^-12 of the main one) run as BF16 MMAs with K = 16 per instruction (W_lo . X and W . X_lo on
// bf16 copies: 2^-20 of the main term); first_layer has no activation and is folded into the trunk block's expand weights.
//
// One persistent CTA per SM, 512 threads, one tile = 16 leaves = 128 "columns" (leaf, feature f<8; f = 7 is padding):
//   * activations live in shared memory as K-major 128-byte-swizzled MMA operands  X[column][token]  (hi and lo planes)
//   * first_layer / project:  D[column (TMEM lane)][channel] = X . W^T     -- weights are the N-side operand; hi and lo
//       weight rows are stacked along N (one N=128 MMA gives W_hi X_hi | W_lo X_hi, one N=64 MMA adds W_hi X_lo)
//   * expand:                 D[channel (TMEM lane)][column] = W . X^T     -- channels on lanes so that the per-channel
//       Linear(7->7)+BN+act over the feature axis and the SE pooling are thread-local in the epilogue (one thread = one
//       channel, its registers = the 8 features of 4 leaves); 168 channels = two M=128 MMAs, the second one reads past
//       the 168 weight rows (rows are independent, the extra lanes are never read)
//   * weights arrive as pre-swizzled hi/lo images through cp.async.bulk + mbarrier; the SE / policy / value linears
//     (fp32 on CUDA cores) read their weights from shared memory rings fed the same way
//   * the expanded activations make one round trip through TMEM (tcgen05.st) between the depthwise pass and the SE-gated
//     operand pass, so shared memory only ever holds two 32-channel K-chunks of them (double buffered against the MMAs)
#pragma once
#include "net_v80.cuh"
#include "umma.cuh"
#include <cuda_bf16.h>

namespace azg {

#ifdef AZG_TC_MAXNREG
#define TC_BOUNDS __maxnreg__(AZG_TC_MAXNREG)
#else
#define TC_BOUNDS __launch_bounds__(TC_THREADS, 1)
#endif
constexpr int TC_THREADS = 512;           // 16 warps: TMEM lane quarter = warp & 3, sub-slice = warp >> 2
constexpr int TC_TB = 16;                 // leaves per tile (128 columns)
// ---- shared memory map (bytes from a 1024-aligned base) ----
constexpr int TC_XH = 0;                  // activations hi: [2 K atoms][128 columns][128 B]
constexpr int TC_XL = 32768;              // the small 3xTF32 terms of the expand run in BF16: [X as bf16: 128 columns x 128 B][X_lo = x - rn_tf32(x) as bf16], 16 KB each
constexpr int TC_ESTG = 65536;            // 2 stages x (EH 16 KB | EL 16 KB); also raw boards, SE fc weights, head activations
constexpr int TC_WRING = 131072;          // 64 KB: project weight slots 4 x 16 KB, first-layer image, policy/value weight ring
constexpr int TC_SQ = 196608;             // SE pooled values [leaf quad][169 (padded)][4] floats: consecutive channels 16 B apart (the depthwise threads store without
                                          // bank conflicts), odd pitch (a quarter-warp of fc1 reads the four leaf quads of one channel from different banks)
constexpr int TC_HID = TC_SQ + 4 * 169 * 16;   // SE hidden [40][16]
constexpr int TC_VH = TC_HID + 40 * 16 * 4;    // value-head accumulators [8][16]
constexpr int TC_SV = TC_VH + 8 * 16 * 4;        // small vectors (biases, BN scale/shift, value-head matrix), copied once per CTA
constexpr int SV_B0 = 0, SV_BLK = 64, SV_BLK_STRIDE = 800, SV_BE = 0, SV_SD = 168, SV_TD = 336, SV_B1 = 504, SV_B2 = 552, SV_BP = 720;
constexpr int SV_BPI2 = 2464, SV_BPI4 = 2560, SV_BV2 = 2656, SV_BV4 = 2660, SV_V4 = 2664, SV_FLOATS = 2680;
constexpr int TC_SMEM = TC_SV + SV_FLOATS * 4 + 1024;   // + alignment slack
constexpr int TC_WE_BYTES = 2 * 2 * 176 * 128;       // expand image: W_hi tf32 (2 atoms x 176 rows x 128 B), then W as bf16 and W_lo as bf16 (one atom each) = 90112 B (spans ESTG + WRING[0,24576))
constexpr int TC_WE_ATOM = 176 * 128;
constexpr int TC_PIRING_SLOT = 18816;     // 56 rows of the [392][84] policy matrix
// ---- TMEM columns ----
constexpr int TC_DE = 0;                  // expand accumulator: [0,128) channels 0-127, [128,256) channels 128-167 on lanes 0-39
constexpr int TC_DP = 256;                // first_layer / project accumulator: [256,320) hi-weight part, [320,384) lo-weight part
constexpr int TC_T = 384;                 // the block input (fp32, [column][token]) is parked here while the E stages borrow X's shared memory

struct V80TCImg { int w0; int we[3]; int wp[3]; int be0; int total; };   // float offsets into the image blob (be0: expand bias of block 0 with first_layer folded in)
inline V80TCImg v80tc_layout() {
    V80TCImg I; int o = 0;
    I.w0 = o; o += 32768 / 4;
    for (int b = 0; b < 3; b++) { I.we[b] = o; o += TC_WE_BYTES / 4; I.wp[b] = o; o += 6 * 16384 / 4; }
    I.be0 = o; o += 176;
    I.total = o; return I;
}
inline float tc_rn_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
// Host: build the operand images from the prepared fp32 blob (BN already folded, K-major [k][out], see v80_prepare).
inline void v80tc_prepare(const float* blob, const V80Layout& L, const V80TCImg& I, float* img) {
    using umma::sw128_off;
    for (int i = 0; i < I.total; i++) img[i] = 0.f;
    const int NV = 56, E = 168;
    auto put = [&](int base_f, size_t byte_off, float v) { img[base_f + byte_off / 4] = v; };
    for (int o = 0; o < NV; o++) for (int k = 0; k < NV; k++) {          // first_layer: rows o (hi) and 64 + o (lo), stacked along N
        const float w = blob[L.w0 + k * NV + o], hi = tc_rn_tf32(w), lo = w - hi;
        put(I.w0, (size_t)(k >> 5) * 16384 + sw128_off(o, k & 31), hi); put(I.w0, (size_t)(k >> 5) * 16384 + sw128_off(64 + o, k & 31), lo);
    }
    for (int b = 0; b < 3; b++) {
        const auto& B = L.blk[b];
        for (int c = 0; c < E; c++) for (int k = 0; k < NV; k++) {       // expand: M-side operand, rows = channels
            float w = blob[B.we + k * E + c];
            if (b == 0) {   // first_layer (Linear + BN, no activation) is folded into the trunk block's expand: W' = We . W0 (accumulated in double);
                double a = 0;   // the expand of block 0 then reads the raw board planes, which are exact in TF32 (no lo plane, two passes)
                for (int j = 0; j < NV; j++) a += (double)blob[L.w0 + k * NV + j] * (double)blob[B.we + j * E + c];
                w = (float)a;
            }
            const float hi = tc_rn_tf32(w), lo = w - hi;
            put(I.we[b], (size_t)(k >> 5) * TC_WE_ATOM + sw128_off(c, k & 31), hi);
            uint16_t* h16 = reinterpret_cast<uint16_t*>(img + I.we[b]);  // atoms 2, 3: W and W_lo as bf16 (K = 64 tokens = one atom) for the small terms (bf16: W_lo ~ 2^-12 |w| would be subnormal in fp16)
            h16[((size_t)2 * TC_WE_ATOM + umma::sw128_off_h(c, k)) / 2] = umma::f32_to_bf16_host(w);
            h16[((size_t)3 * TC_WE_ATOM + umma::sw128_off_h(c, k)) / 2] = umma::f32_to_bf16_host(lo);
        }
        if (b == 0) for (int c = 0; c < E; c++) {                        // be' = be + We . b0
            double a = blob[B.be + c];
            for (int j = 0; j < NV; j++) a += (double)blob[L.b0 + j] * (double)blob[B.we + j * E + c];
            img[I.be0 + c] = (float)a;
        }
        for (int o = 0; o < NV; o++) for (int c = 0; c < E; c++) {       // project: chunk = 32 input channels, rows o (hi) / 64 + o (lo)
            const float w = blob[B.wp + c * NV + o], hi = tc_rn_tf32(w), lo = w - hi;
            put(I.wp[b], (size_t)(c >> 5) * 16384 + sw128_off(o, c & 31), hi); put(I.wp[b], (size_t)(c >> 5) * 16384 + sw128_off(64 + o, c & 31), lo);
        }
    }
}

namespace tc {
using namespace umma;
enum { B_W0 = 0, B_WE, B_FC, B_WP0, B_WP1, B_WP2, B_WP3, B_EF0, B_EF1, B_EF2, B_EF3, B_MMA, B_PI0, B_PI1, B_PI2, B_PI3, B_V2, B_MM2, B_FL, B_N };
struct Phase {                            // per-thread parity of every barrier this thread waits on
    uint32_t bits = 0;
    __device__ __forceinline__ void wait(uint64_t* bars, int id) { mbar_wait(&bars[id], (bits >> id) & 1u); bits ^= 1u << id; }
};
__device__ __forceinline__ void load(uint64_t* bars, int id, void* dst, const void* src, uint32_t bytes) {
    mbar_expect_tx(&bars[id], bytes); bulk_g2s(dst, src, bytes, &bars[id]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// acc[ip][j] (+)= (w[2 ip], w[2 ip + 1]) * x[j] for a 4 x 4 register tile: 8 FFMA2
__device__ __forceinline__ void tile44_fma2(uint64_t (&acc)[2][4], const float4 w4, const float4 x4) {
    const uint64_t w01 = pk2(w4.x, w4.y), w23 = pk2(w4.z, w4.w);
    const uint64_t xd[4] = {pk2(x4.x, x4.x), pk2(x4.y, x4.y), pk2(x4.z, x4.z), pk2(x4.w, x4.w)};
#pragma unroll
    for (int j = 0; j < 4; j++) { fma2(acc[0][j], w01, xd[j]); fma2(acc[1][j], w23, xd[j]); }
}
__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {          // hi = x rounded to TF32, hi + lo == x exactly
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = __fsub_rn(x, hi);
}

// Depthwise pass of one block for one (channel = TMEM lane, 4 leaves = 32 columns) unit: expand bias + act, Linear(7->7) over
// the features + BN + act, SE pooling; the result replaces the accumulator in TMEM.
template <int ACT, bool SE_MAX>
__device__ __forceinline__ void depthwise_unit(uint32_t taddr, float be, float sd, float td, const float* __restrict__ dw, float* sq4, bool valid) {
    uint32_t r[32];
    tmem_ld32(taddr, r); tmem_wait_ld();
    float pl[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        float v[7];
#pragma unroll
        for (int f = 0; f < 7; f++) v[f] = act_apply(__uint_as_float(r[8 * l + f]) + be, ACT);
        float pool = SE_MAX ? -INFINITY : 0.f;
#pragma unroll
        for (int g = 0; g < 7; g++) {
            float a = 0.f;
#pragma unroll
            for (int f = 0; f < 7; f++) a = fmaf(dw[g * 7 + f], v[f], a);
            a = act_apply(fmaf(a, sd, td), ACT);
            r[8 * l + g] = __float_as_uint(a);
            pool = SE_MAX ? fmaxf(pool, a) : pool + a;
        }
        r[8 * l + 7] = 0u;
        pl[l] = SE_MAX ? pool : pool * (1.f / 7.f);
    }
    tmem_st32(taddr, r);
    if (valid) *reinterpret_cast<float4*>(sq4) = make_float4(pl[0], pl[1], pl[2], pl[3]);
}
}  // namespace tc

template <int NP>
__global__ void TC_BOUNDS
k_v80_tc(const float* __restrict__ P, const float* __restrict__ IMG, const __grid_constant__ V80Layout L, const __grid_constant__ V80TCImg I,
         const __grid_constant__ V80DW DW, const int* count_ptr, const int* list, const int8_t* boards, int bstride,
         const uint32_t* masks, float* pi_out, float* v_out, int n_max, long long* prof) {
    using namespace tc;
    constexpr int NV = 56, EC = 168, Q = V80_Q, TB = TC_TB, A = 81, PIP = V80Layout::PIP, SQP = EC + 1;
    // head inputs [token][feature][leaf] with a feature pitch of 20 floats: the 32 (leaf, feature) rows a warp stores in the project epilogue
    // fall into 32 different banks; rows stay 16-byte aligned for the policy head's 128-bit loads
    constexpr int HFP = 20, HLD = 7 * HFP, HEAD_X0 = NV * HLD * 4, LGP = 85;
    extern __shared__ uint8_t smem_raw[];
    // 1024-aligned base as symbol + offset: the pointer stays in the shared state space (LDS / STS). Rounding the generic address instead
    // turns every access below into a generic LD.E / ST.E.
    uint8_t* sm = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bars[B_N];
    __shared__ uint32_t tmem_s;
    __shared__ int slot_of[TB], slot_next[TB];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, q = warp & 3, sub = warp >> 2;
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int ntiles = (count + TB - 1) / TB;
    if ((int)blockIdx.x >= ntiles) return;
    if (prof && t == 0 && blockIdx.x < 160) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); prof[64 + 4 * blockIdx.x] = g; }   // debug: CTA timeline (ns)
    float* XH = reinterpret_cast<float*>(sm + TC_XH);
    float* SQ = reinterpret_cast<float*>(sm + TC_SQ); float* HID = reinterpret_cast<float*>(sm + TC_HID); float* VH = reinterpret_cast<float*>(sm + TC_VH);
    uint8_t* ESTG = sm + TC_ESTG; uint8_t* WRING = sm + TC_WRING;
    float* SV = reinterpret_cast<float*>(sm + TC_SV);
    {   // small vectors -> shared memory, once per (persistent) CTA
        auto cp = [&](int dst, int src, int n) { for (int i = t; i < n; i += TC_THREADS) SV[dst + i] = __ldg(P + src + i); };
        cp(SV_B0, L.b0, 56);
        for (int b = 0; b < 3; b++) {
            const int o = SV_BLK + b * SV_BLK_STRIDE; const V80Layout::Blk& B = L.blk[b];
            if (b == 0) { for (int i = t; i < 168; i += TC_THREADS) SV[o + SV_BE + i] = __ldg(IMG + I.be0 + i); } else cp(o + SV_BE, B.be, 168);
            cp(o + SV_SD, B.sd, 168); cp(o + SV_TD, B.td, 168); cp(o + SV_B1, B.b1, 40); cp(o + SV_B2, B.b2, 168); cp(o + SV_BP, B.bp, 56);
        }
        cp(SV_BPI2, L.bpi2, 84); cp(SV_BPI4, L.bpi4, 84); cp(SV_BV2, L.bv2, 4); cp(SV_BV4, L.bv4, 4); cp(SV_V4, L.v4, 16);
    }
    if (t == 0) { for (int i = 0; i < B_N; i++) mbar_init(&bars[i], 1u); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_s);
    for (int i = t; i < 65536 / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);   // token columns 56..63 only ever meet zero weights (they hold stale finite data once X's planes have served as E stages)
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_s;
    const uint32_t tlane = tm + ((uint32_t)(32 * q) << 16);       // this warp's TMEM lane quarter
    if (t == 0) { for (int i = 0; i < 4; i++) mbar_arrive(&bars[B_EF0 + i]); }   // the four E stages start out free
    if (prof && t == 0 && blockIdx.x < 160) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); prof[64 + 4 * blockIdx.x + 1] = g; }
    Phase ph;
    uint32_t rb[4] = {0u, 0u, 0u, 0u};                          // this thread's words of the next tile's raw boards (cross-tile prefetch)
    int prof_i = 0;
#define TC_STAMP() do { if (prof && t == 0 && blockIdx.x == 0 && prof_i < 64) prof[prof_i++] = clock64(); } while (0)
#ifdef AZG_TC_FINE_PROF
#define TC_FSTAMP() do { if (b == 0) TC_STAMP(); } while (0)      /* debug build: extra stamps inside the phases of block 0 */
#else
#define TC_FSTAMP() do { } while (0)
#endif
    const uint32_t xh_a = smem_u32(sm + TC_XH), xl_a = smem_u32(sm + TC_XL), estg_a = smem_u32(ESTG), wring_a = smem_u32(WRING);
    constexpr uint32_t ID128 = idesc_tf32(128, 128), ID64 = idesc_tf32(128, 64), IDH128 = idesc_bf16(128, 128);
    const float* IMGb = IMG;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tile0 = tile * TB;
        TC_STAMP();   /* 0: tile start */
        const bool prefetched = tile != (int)blockIdx.x;          // the previous tile's value head already fetched this tile's boards (rb)
        if (t < TB) { const int j = tile0 + t; slot_of[t] = prefetched ? slot_next[t] : (j < count ? (list ? list[j] : j) : -1); }
        if (t == 0) {                                             // first-layer image -> weight slots 2, 3; the trunk block's expand image -> ESTG .. WRING[0, 24 KB)
            load(bars, B_W0, WRING + 32768, IMGb + I.w0, 32768);
            load(bars, B_WE, ESTG, IMGb + I.we[0], TC_WE_BYTES);
        }
        if (!prefetched) __syncthreads();
        {   // raw boards -> the SE scratch (16 x 400 B; ESTG is receiving the expand image), 32-bit loads (board rows are 392 B, slots are 4-byte aligned)
            uint32_t* raw = reinterpret_cast<uint32_t*>(sm + TC_SQ);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = t + i * TC_THREADS;
                if (k < TB * 98) {
                    const int s = k / 98, w = k - s * 98;
                    if (!prefetched) { const int slot = slot_of[s]; rb[i] = slot >= 0 ? reinterpret_cast<const uint32_t*>(boards + (size_t)slot * bstride)[w] : 0u; }
                    raw[s * 100 + w] = rb[i];
                }
            }
        }
        __syncthreads();
        {   // XH[column][token] = (float)board[leaf][token][f]  (small integers: exact in TF32, no lo plane needed)
            const int8_t* raw = reinterpret_cast<const int8_t*>(sm + TC_SQ);
            for (int k = t; k < 128 * 14; k += TC_THREADS) {
                const int row = k & 127, kq = k >> 7, s = row >> 3, f = row & 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f < 7) { const int8_t* b = raw + s * 400 + (4 * kq) * 7 + f; v = make_float4((float)b[0], (float)b[7], (float)b[14], (float)b[21]); }
                *reinterpret_cast<float4*>(sm + TC_XH + (kq >> 3) * 16384 + sw128_off(row, (4 * kq) & 31)) = v;
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);       // the same planes as bf16 (small integers: exact)
                *reinterpret_cast<uint2*>(sm + TC_XL + sw128_off_h(row, 4 * kq)) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            }
            // token columns 56..63 of the bf16 plane must be ZERO: the region served as an E stage, and stale bits read as bf16 can be
            // Inf / NaN, which a zero weight does not cancel (stale fp32 data in XH is always finite)
            if (t < 256) *reinterpret_cast<uint2*>(sm + TC_XL + sw128_off_h(t & 127, 56 + 4 * (t >> 7))) = make_uint2(0u, 0u);
        }
        fence_async_smem(); __syncthreads();
        TC_STAMP();   /* 1: input staged */
        // ---------------- first_layer: D[column][channel] = X . W0^T, hi|lo weight rows stacked along N. Its output T is only needed as the
        // residual of the trunk block (first_layer has no activation, so the block's expand reads the raw planes through the folded
        // weights We . W0): the MMAs are queued here, the expand MMAs of block 0 right behind them, and T goes from the accumulator
        // straight to its parking columns in TMEM while the expand MMAs run. ----------------
        if (t == TC_THREADS - 32) {
            ph.wait(bars, B_W0); tc_fence_after();
#pragma unroll 1
            for (int ks = 0; ks < 7; ks++) {
                const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
                mma_tf32(tm + TC_DP, desc_sw128(xh_a + off), desc_sw128(wring_a + 32768 + off), ID128, ks > 0);
            }
            mma_commit(&bars[B_FL]);
        }
        __syncwarp();

#pragma unroll 1
        for (int b = 0; b < 3; b++) {
            const V80Layout::Blk& B = L.blk[b];
            const float* dw = DW.w[b];
            if (b == 2 && t < TB) {                              // leaf slots of this CTA's next tile: the load has the whole value block to land
                const int j = (tile + (int)gridDim.x) * TB + t;
                slot_next[t] = (tile + (int)gridDim.x < ntiles && j < count) ? (list ? list[j] : j) : -1;
            }
            const float* SB = SV + SV_BLK + b * SV_BLK_STRIDE;
            // ---------------- expand: D[channel][column] = We . X^T (3 passes x 7 k-steps x 2 channel halves) ----------------
            if (t == TC_THREADS - 32) {                           // issued from the last warp (one depthwise unit), not from warp 0 (two units): the
                ph.wait(bars, B_WE); tc_fence_after();            // issuing thread only reaches its own epilogue work once all 42 MMAs are queued
#pragma unroll 1
                for (int mh = 0; mh < 2; mh++) {
                    const uint32_t wa = estg_a + mh * 16384, dcol = tm + TC_DE + 128 * mh;
                    // the two small terms first, in BF16 (K = 16 per MMA, 4 k-steps for the 56 tokens): W_lo . X and, unless the input is
                    // the raw planes (block 0: exact, no lo part), W . X_lo; then the main term W_hi . X_hi in TF32 (7 k-steps)
#pragma unroll 1
                    for (int ks = 0; ks < 4; ks++)
                        mma_f16(dcol, desc_sw128(wa + 3 * TC_WE_ATOM + ks * 32), desc_sw128(xl_a + ks * 32), IDH128, ks != 0);
                    if (b != 0) {
#pragma unroll 1
                        for (int ks = 0; ks < 4; ks++)
                            mma_f16(dcol, desc_sw128(wa + 2 * TC_WE_ATOM + ks * 32), desc_sw128(xl_a + 16384 + ks * 32), IDH128, true);
                    }
#pragma unroll 1
                    for (int ks = 0; ks < 7; ks++)
                        mma_tf32(dcol, desc_sw128(wa + (ks >> 2) * TC_WE_ATOM + (ks & 3) * 32), desc_sw128(xh_a + (ks >> 2) * 16384 + (ks & 3) * 32), ID128, true);
                    mma_commit(&bars[mh == 0 ? B_MMA : B_MM2]);      // channels 0-127 complete first: their depthwise pass overlaps the MMAs of 128-167
                }
            }
            __syncwarp();
            if (b == 0) {   // trunk block: its residual T = first_layer(x) comes out of the first-layer accumulator (bias added) into the parking columns
                const int c0 = 16 * sub;
                ph.wait(bars, B_FL); tc_fence_after();
                if (t == 0) {                                     // the first-layer image is consumed: project chunks 0, 1 -> slots 2, 3
                    load(bars, B_WP2, WRING + 32768, IMGb + I.wp[0], 16384);
                    load(bars, B_WP3, WRING + 49152, IMGb + I.wp[0] + 4096, 16384);
                }
                __syncwarp();
                uint32_t dh[16], dl[16], xv[16];
                tmem_ld16(tlane + TC_DP + c0, dh); tmem_ld16(tlane + TC_DP + 64 + c0, dl); tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; j++) xv[j] = __float_as_uint(__uint_as_float(dh[j]) + __uint_as_float(dl[j]) + (c0 + j < NV ? SV[SV_B0 + c0 + j] : 0.f));
                tmem_st16(tlane + TC_T + c0, xv);
            }   // (blocks 1 and 2 find their input already parked: block 0's project epilogue put the trunk output there)
            ph.wait(bars, B_MMA); tc_fence_after();
            TC_STAMP();   /* b1: expand MMAs (channels 0-127) done */
            // ---------------- depthwise pass (thread = channel) ----------------
            {
                const int c = 32 * q + lane;
                const float be = SB[SV_BE + c], sd = SB[SV_SD + c], td = SB[SV_TD + c];
                const uint32_t ta = tlane + TC_DE + 32 * sub;
                if (b == 0) depthwise_unit<1, false>(ta, be, sd, td, dw, SQ + (sub * SQP + c) * 4, true);
                else depthwise_unit<2, true>(ta, be, sd, td, dw, SQ + (sub * SQP + c) * 4, true);
                TC_FSTAMP();   /* f: unit 1 done */
                ph.wait(bars, B_MM2); tc_fence_after();
                TC_FSTAMP();   /* f: MM2 wait done */
                // The expand image is dead: SE weights and project chunks 2, 3 take its place. Issued from quarters 2 and 3, which have no
                // second depthwise unit (a bulk-copy issue costs its thread a few hundred cycles).
                if (t == TC_THREADS - 32) {
                    mbar_expect_tx(&bars[B_FC], 2 * 26880);
                    bulk_g2s(ESTG, P + B.fc1, 26880, &bars[B_FC]); bulk_g2s(ESTG + 26880, P + B.fc2, 26880, &bars[B_FC]);
                }
                if (t == TC_THREADS - 64) {
                    load(bars, B_WP0, WRING, IMGb + I.wp[b] + 2 * 4096, 16384);
                    load(bars, B_WP1, WRING + 16384, IMGb + I.wp[b] + 3 * 4096, 16384);
                }
                __syncwarp();
                TC_FSTAMP();   /* f: parked */
                if (q < 2) {
                    const int c2 = 128 + c; const bool ok = c2 < EC; const int cc = ok ? c2 : 0;
                    const float be2 = SB[SV_BE + cc], sd2 = SB[SV_SD + cc], td2 = SB[SV_TD + cc];
                    if (b == 0) depthwise_unit<1, false>(ta + 128, be2, sd2, td2, dw, SQ + (sub * SQP + cc) * 4, ok);
                    else depthwise_unit<2, true>(ta + 128, be2, sd2, td2, dw, SQ + (sub * SQP + cc) * 4, ok);
                }
                tmem_wait_st();
                TC_FSTAMP();   /* f: unit 2 done */
            }
            __syncthreads();
            TC_STAMP();   /* b2: depthwise done */
            // ---------------- squeeze-excitation: fc1 (168 -> 40, ReLU), fc2 (40 -> 168, hardsigmoid) ----------------
            ph.wait(bars, B_FC);
            TC_STAMP();   /* b3: fc weights landed */
            // fc2 K-split partial sums [3][leaf quad][169][4] (32 448 B): consecutive channels 16 B apart for the gated pass (one thread per
            // channel), the odd pitch spreads the four leaf quads of the producers' stores over the banks
            float* GP = reinterpret_cast<float*>(sm + TC_XH); constexpr int GPP = EC + 1;
            {
                const float* W1 = reinterpret_cast<const float*>(ESTG); const float* W2 = reinterpret_cast<const float*>(ESTG + 26880);
                // Register-tiled: a thread owns 4 outputs x 4 leaves, 16 FMA per pair of 128-bit loads; the K range is split 12 / 3 ways so
                // that all 16 warps issue (these loops are FMA-issue bound). The partial sums meet in X's planes, which are free: the
                // block input is parked in TMEM. Fixed summation order: results do not depend on the launch.
                float* HP = reinterpret_cast<float*>(sm + TC_XH);                // fc1 partial sums [12][leaf quad][41 (padded)][4] (31 488 B)
                if (t < 480) {
                    const int part = t / 40, rem = t - part * 40, qg = rem >> 2, lq = rem & 3;
                    uint64_t a2[2][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) a2[0][j] = a2[1][j] = 0ull;
#pragma unroll 7
                    for (int kk = 14 * part; kk < 14 * part + 14; kk++)
                        tile44_fma2(a2, *reinterpret_cast<const float4*>(W1 + kk * Q + 4 * qg), *reinterpret_cast<const float4*>(SQ + (lq * SQP + kk) * 4));
                    float a[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) { upk2(a2[0][j], a[0][j], a[1][j]); upk2(a2[1][j], a[2][j], a[3][j]); }
#pragma unroll
                    for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(HP + ((part * 4 + lq) * 41 + 4 * qg + i) * 4) = make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
                }
                TC_FSTAMP();   /* f: fc1 partials */
                __syncthreads();
                for (int i = t; i < Q * TB; i += TC_THREADS) {
                    const int lq = i / 160, rem = i - lq * 160;      // rem = 4 * hidden unit + leaf of the quad
                    float h = HP[lq * 164 + rem];
#pragma unroll
                    for (int pt = 1; pt < 12; pt++) h += HP[(pt * 4 + lq) * 164 + rem];
                    HID[(rem >> 2) * TB + 4 * lq + (rem & 3)] = fmaxf(h + SB[SV_B1 + (rem >> 2)], 0.f);
                }
                __syncthreads();
                TC_FSTAMP();   /* f: hidden done + barrier */
                if (t < 504) {                                     // 3 K slices x 42 channel quads x 4 leaf quads
                    const int part = t / 168, rem = t - part * 168, cq = rem >> 2, lq = rem & 3;
                    const int k0 = part == 0 ? 0 : (part == 1 ? 14 : 27), k1 = part == 0 ? 14 : (part == 1 ? 27 : 40);
                    uint64_t g2[2][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) g2[0][j] = g2[1][j] = 0ull;
#pragma unroll 7
                    for (int kk = k0; kk < k1; kk++)
                        tile44_fma2(g2, *reinterpret_cast<const float4*>(W2 + kk * EC + 4 * cq), *reinterpret_cast<const float4*>(HID + kk * TB + 4 * lq));
                    float g[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) { upk2(g2[0][j], g[0][j], g[1][j]); upk2(g2[1][j], g[2][j], g[3][j]); }
#pragma unroll
                    for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(GP + ((part * 4 + lq) * GPP + 4 * cq + i) * 4) = make_float4(g[i][0], g[i][1], g[i][2], g[i][3]);
                }
                TC_FSTAMP();   /* f: fc2 done */
            }
            __syncthreads();
            TC_STAMP();   /* b4: SE done */
            // ---------------- SE-gated operand pass + project MMAs. Six 32-channel K chunks through FOUR stages (two in ESTG, two in X's
            //                  planes). Round A: warp quarter q writes chunk q into stage q (all 16 warps); round B: quarters 0 / 1 write
            //                  chunks 4 / 5 into stages 0 / 1 as soon as the MMAs of chunks 0 / 1 are done, while those of chunks 2, 3 run.
            //                  All MMAs are issued by one thread of the last warp (quarter 3: no round-B work). ----------------
            {
                auto gate_of = [&](int c) -> float4 {             // hardsigmoid(fc2 + b2) of channel c for this thread's 4 leaves
                    const float4 g0 = *reinterpret_cast<const float4*>(GP + (sub * GPP + c) * 4), g1 = *reinterpret_cast<const float4*>(GP + ((4 + sub) * GPP + c) * 4),
                                 g2 = *reinterpret_cast<const float4*>(GP + ((8 + sub) * GPP + c) * 4);
                    const float bb = SB[SV_B2 + c] + 3.f;
                    return make_float4(fminf(fmaxf(g0.x + g1.x + g2.x + bb, 0.f), 6.f) * (1.f / 6.f), fminf(fmaxf(g0.y + g1.y + g2.y + bb, 0.f), 6.f) * (1.f / 6.f),
                                       fminf(fmaxf(g0.z + g1.z + g2.z + bb, 0.f), 6.f) * (1.f / 6.f), fminf(fmaxf(g0.w + g1.w + g2.w + bb, 0.f), 6.f) * (1.f / 6.f));
                };
                const int cA = 32 * q + lane, cB = 128 + 32 * q + lane;
                const bool okB = q < 2 && cB < EC;
                const float4 gA = gate_of(cA), gB = gate_of(okB ? cB : 0);
                __syncthreads();                                  // every gate is in registers: the partial sums (stage 2) may be overwritten
                // element i of a thread -> row 32 sub + i, k = lane of the stage: byte offset (i >> 3) * 1024 + (i & 7) * 128 +
                // (((lane >> 2) ^ (i & 7)) << 4) + (lane & 3) * 4. The lane-dependent part only depends on i & 7: eight base pointers.
                auto gated_write = [&](uint32_t tcol, const float4 g4, int stg, bool ok) {
                    uint32_t d[32];
                    tmem_ld32(tlane + tcol, d);                   // warp-collective
                    tmem_wait_ld();
                    if (!ok) return;
                    const float gate[4] = {g4.x, g4.y, g4.z, g4.w};
                    uint8_t* eh = (stg < 2 ? ESTG + stg * 32768 : sm + TC_XH + (stg - 2) * 32768) + sub * 4096;   // rows 32 sub .. 32 sub + 31 of the stage
                    uint8_t* eb[8];
#pragma unroll
                    for (int v = 0; v < 8; v++) eb[v] = eh + ((((lane >> 2) ^ v) & 7) << 4) + ((lane & 3) << 2);
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        float hi, lo; split_rn(__uint_as_float(d[i]) * gate[i >> 3], hi, lo);
                        uint8_t* pdst = eb[i & 7] + (i >> 3) * 1024 + (i & 7) * 128;
                        *reinterpret_cast<float*>(pdst) = hi; *reinterpret_cast<float*>(pdst + 16384) = lo;
                    }
                };
                auto issue_chunk = [&](int j) {                   // chunk j: stage j & 3, weight slot (j + 2) & 3
                    const int slot = (j + 2) & 3, stg = j & 3;
                    ph.wait(bars, B_WP0 + slot); tc_fence_after();
                    const uint32_t ea = stg < 2 ? estg_a + stg * 32768 : xh_a + (stg - 2) * 32768, wa = wring_a + slot * 16384;
                    const uint64_t dh_ = desc_sw128(ea), dl_ = desc_sw128(ea + 16384), dw_ = desc_sw128(wa);
                    const int nks = j == 5 ? 1 : 4;
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        if (ks < nks) {
                            mma_tf32(tm + TC_DP, dh_ + 2 * ks, dw_ + 2 * ks, ID128, (j | ks) != 0);     // E_hi . (W_hi | W_lo); +32 B per k-step = +2 in the descriptor
                            mma_tf32(tm + TC_DP, dl_ + 2 * ks, dw_ + 2 * ks, ID64, true);               // E_lo . W_hi
                        }
                    }
                    mma_commit(&bars[B_EF0 + stg]);
                };
                // ---- round A
                ph.wait(bars, B_EF0 + q);                         // the MMAs that last read stage q are done
#ifdef AZG_TC_ROUND_PROF
                if (b == 0) TC_STAMP();
#endif
                gated_write(TC_DE + 32 * sub, gA, q, true);
#ifdef AZG_TC_ROUND_PROF
                if (b == 0) TC_STAMP();
#endif
                fence_async_smem(); tc_fence_before(); __syncthreads();
                if (t == TC_THREADS - 32) {
                    tc_fence_after();
#pragma unroll 1
                    for (int j = 0; j < 4; j++) issue_chunk(j);
                }
                __syncwarp();
                // ---- round B
                if (q < 2) {                                      // warp-uniform
                    ph.wait(bars, B_EF0 + q);                     // chunk q's MMAs are done: stage q and weight slot q + 2 are free
                    if (sub == 0 && lane == 0) load(bars, B_WP2 + q, WRING + (2 + q) * 16384, IMGb + I.wp[b] + (4 + q) * 4096, 16384);
                    __syncwarp();
#ifdef AZG_TC_ROUND_PROF
                    if (b == 0) TC_STAMP();
#endif
                    gated_write(TC_DE + 128 + 32 * sub, gB, q, okB);
#ifdef AZG_TC_ROUND_PROF
                    if (b == 0) TC_STAMP();
#endif
                }
                fence_async_smem(); tc_fence_before(); __syncthreads();
                if (t == TC_THREADS - 32) {
                    tc_fence_after();
                    issue_chunk(4); issue_chunk(5);
                    mma_commit(&bars[B_MMA]);
                }
                __syncwarp();
            }
            ph.wait(bars, B_MMA);                                 // committed after the last chunk by the thread that issued every MMA: all of them are done
            tc_fence_after();
            TC_STAMP();   /* b5: project MMAs done */
            if (lane == 0 && warp >= 12) {                        // everything in ESTG / WRING has been consumed; one bulk copy per warp (12 .. 15)
                const int w = warp - 12;
                if (b == 0) {
                    if (w == 0) load(bars, B_WE, ESTG, IMGb + I.we[1], TC_WE_BYTES);
                    if (w == 1) load(bars, B_WP2, WRING + 32768, IMGb + I.wp[1], 16384);
                    if (w == 2) load(bars, B_WP3, WRING + 49152, IMGb + I.wp[1] + 4096, 16384);
                } else if (b == 1) {
                    // policy weight ring: four slots (the last one runs over into SQ, whose gates are consumed), loads in flight ahead of the math
                    load(bars, B_PI0 + w, WRING + w * TC_PIRING_SLOT, P + L.pi2 + w * 56 * PIP, TC_PIRING_SLOT);
                } else {
                    if (w == 0) load(bars, B_V2, WRING, P + L.v2, NV * 7 * 4 * 4);
                }
            }
            __syncwarp();
            {   // project epilogue: bias + residual (the parked block input); trunk output -> X's operand planes, head outputs -> plain
                // [token][feature][leaf] in ESTG; after the policy block the parked trunk output goes back into X's planes for the value block
                const int row = 32 * q + lane, c0 = 16 * sub;
                uint32_t dh[16], dl[16], xr[16];
                tmem_ld16(tlane + TC_DP + c0, dh); tmem_ld16(tlane + TC_DP + 64 + c0, dl); tmem_ld16(tlane + TC_T + c0, xr); tmem_wait_ld();
                float* HO = reinterpret_cast<float*>(ESTG);
                uint32_t yv[16];
#pragma unroll
                for (int j = 0; j < 16; j++) yv[j] = 0u;
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const int c = c0 + 4 * j4;
                    if (c < NV) {
                        const uint32_t o = (c >> 5) * 16384 + sw128_off(row, c & 31);
                        float y[4], hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) y[j] = __uint_as_float(dh[4 * j4 + j]) + __uint_as_float(dl[4 * j4 + j]) + SB[SV_BP + c + j] + __uint_as_float(xr[4 * j4 + j]);
                        if (b != 0) {
#pragma unroll
                            for (int j = 0; j < 4; j++) { if ((row & 7) < 7) HO[(c + j) * HLD + (row & 7) * HFP + (row >> 3)] = y[j]; y[j] = __uint_as_float(xr[4 * j4 + j]); }   // [token][feature][leaf]
                        }
                        if (b != 2) {                             // b == 0: the trunk output; b == 1: the trunk output again (X's planes served as E stages)
#pragma unroll
                            for (int j = 0; j < 4; j++) { split_rn(y[j], hi[j], lo[j]); yv[4 * j4 + j] = __float_as_uint(y[j]); }
                            *reinterpret_cast<float4*>(sm + TC_XH + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                            const uint32_t oh = sw128_off_h(row, c);
                            const __nv_bfloat162 x0 = __floats2bfloat162_rn(y[0], y[1]), x1 = __floats2bfloat162_rn(y[2], y[3]), l0 = __floats2bfloat162_rn(lo[0], lo[1]), l1 = __floats2bfloat162_rn(lo[2], lo[3]);
                            *reinterpret_cast<uint2*>(sm + TC_XL + oh) = make_uint2(*reinterpret_cast<const uint32_t*>(&x0), *reinterpret_cast<const uint32_t*>(&x1));
                            *reinterpret_cast<uint2*>(sm + TC_XL + 16384 + oh) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
                        }
                    }
                }
                if (b != 2 && sub == 3) {                         // token columns 56..63 of both bf16 planes: zero (see the staging pass)
#pragma unroll
                    for (int cz = 56; cz < 64; cz += 4) {
                        *reinterpret_cast<uint2*>(sm + TC_XL + sw128_off_h(row, cz)) = make_uint2(0u, 0u);
                        *reinterpret_cast<uint2*>(sm + TC_XL + 16384 + sw128_off_h(row, cz)) = make_uint2(0u, 0u);
                    }
                }
                if (b == 0) { tmem_st16(tlane + TC_T + c0, yv); tmem_wait_st(); }   // the trunk output is the input AND the residual of both head blocks: parked once, here
            }
            fence_async_smem(); tc_fence_before(); __syncthreads();
            TC_STAMP();   /* b6: project epilogue done */

            if (b == 1) {
                // ---------------- policy head linears on CUDA cores: Linear(392 -> 81) + ReLU, Linear(81 -> 81), masked softmax ----------------
                const float* X0 = reinterpret_cast<const float*>(ESTG);
                float* H1 = reinterpret_cast<float*>(ESTG + HEAD_X0); float* LG = H1 + PIP * TB;     // LG: logits [leaf][85] (a warp reads one leaf's row)
                // Weight ring (7 chunks of the 392 x 84 matrix, then the 81 x 84 one in two parts) through four slots. "Full" is the slot's
                // mbarrier (bulk-copy transaction bytes). "Empty" is a NAMED barrier per slot (ids 1-4, 384 threads): the 11 computing warps
                // bar.arrive when they are done with a slot, the last warp (no policy work) bar.syncs on it and its last lane refills the
                // slot. No CTA-wide barrier per chunk, and the write-after-read edge is one compute-sanitizer's racecheck models.
                uint32_t mw[3] = {0u, 0u, 0u};                     // this warp's leaf: legal-move words for the softmax below (the global load lands during the linears)
                {
                    const int slot = slot_of[warp];
                    if (slot >= 0) { mw[0] = masks[(size_t)slot * 3]; mw[1] = masks[(size_t)slot * 3 + 1]; mw[2] = masks[(size_t)slot * 3 + 2]; }
                }
                int ent = 0;
                auto acquire = [&]() -> const float* {
                    const int s = ent & 3;
                    ph.wait(bars, B_PI0 + s);
                    ent++;
                    return reinterpret_cast<const float*>(WRING + s * TC_PIRING_SLOT);
                };
                auto release = [&]() { asm volatile("bar.arrive %0, 384;" ::"r"(1 + ((ent - 1) & 3)) : "memory"); };
                if (warp == TC_THREADS / 32 - 1) {
#pragma unroll 1
                    for (int e = 4; e < 9; e++) {
                        const int s2 = e & 3;
                        asm volatile("bar.sync %0, 384;" ::"r"(1 + s2) : "memory");   // the previous entry of this slot has been consumed by all 11 warps
                        if (lane == 31) {
                            fence_async_smem();                    // order those generic-proxy reads before the async-proxy refill
                            const float* src = e < 7 ? P + L.pi2 + e * 56 * PIP : (e == 7 ? P + L.pi4 : P + L.pi4 + 41 * PIP);
                            const uint32_t bytes = e < 7 ? TC_PIRING_SLOT : (e == 7 ? 41 * PIP * 4 : 40 * PIP * 4);
                            load(bars, B_PI0 + s2, WRING + s2 * TC_PIRING_SLOT, src, bytes);
                        }
                        __syncwarp();
                    }
                }
                // 352 tasks = 8 K-slices x 11 output octets x 4 leaf quads (register tile 8 x 4: 32 FMA per three 128-bit loads; the loads,
                // not the FMAs, are the limit, so the work is spread over 11 warps). Slices s and s + 4 sit in lanes l and l ^ 16 of one warp
                // and are summed by shuffle; the four remaining partial sums meet in PP and are added in a fixed order.
                float* PP = reinterpret_cast<float*>(ESTG + HEAD_X0 + (PIP + LGP) * TB * 4);   // [4][84][16]
                static_assert(HEAD_X0 + (PIP + LGP) * TC_TB * 4 + 4 * PIP * TC_TB * 4 <= 65536, "policy head scratch must fit the E stages");
                const bool live = t < 352;
                const int unit = min((t >> 5) * 16 + (t & 15), 175), half = (t >> 4) & 1;
                const int part = unit / 44, task = unit - part * 44, o8 = task >> 2, lq = task & 3, slice = part + 4 * half;
                const bool hi_ok = o8 < 10;                       // the last octet only has outputs 80..83
                uint64_t acc2a[2][4], acc2b[2][4];                // outputs 0-3 / 4-7 of the octet x 4 leaves, as FFMA2 pairs (2 outputs x 1 leaf)
#pragma unroll
                for (int j = 0; j < 4; j++) acc2a[0][j] = acc2a[1][j] = acc2b[0][j] = acc2b[1][j] = 0ull;
                auto fma_row = [&](const float* wrow, const float4 x4) {
                    const float4 wa = *reinterpret_cast<const float4*>(wrow), wb = hi_ok ? *reinterpret_cast<const float4*>(wrow + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tile44_fma2(acc2a, wa, x4); tile44_fma2(acc2b, wb, x4);
                };
                auto flush = [&]() {                              // lanes l and l ^ 16 hold the two K-slices of one task
                    if (!live) return;                            // warp-uniform (352 = 11 warps)
                    float acc[8][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        upk2(acc2a[0][j], acc[0][j], acc[1][j]); upk2(acc2a[1][j], acc[2][j], acc[3][j]);
                        upk2(acc2b[0][j], acc[4][j], acc[5][j]); upk2(acc2b[1][j], acc[6][j], acc[7][j]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[i][j] += __shfl_xor_sync(FULL, acc[i][j], 16);
                    if (half == 0) {
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            if (i < 4 || hi_ok) *reinterpret_cast<float4*>(PP + (part * PIP + 8 * o8 + i) * TB + 4 * lq) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) acc2a[0][j] = acc2a[1][j] = acc2b[0][j] = acc2b[1][j] = 0ull;
                };
                for (int ch = 0; ch < NV / 8 && live; ch++) {
                    const float* W = acquire();
                    {                                             // K-slice `slice` = token `slice` of this 8-token chunk: 7 rows
                        const float* xp = X0 + (8 * ch + slice) * HLD + 4 * lq;
                        const float* wp = W + (slice * 7) * PIP + 8 * o8;
#pragma unroll
                        for (int f = 0; f < 7; f++) fma_row(wp + f * PIP, *reinterpret_cast<const float4*>(xp + f * HFP));
                    }
                    release();
                }
                flush();
                __syncthreads();
#ifdef AZG_TC_POLICY_PROF
                TC_STAMP();
#endif
                for (int i = t; i < PIP * TB; i += TC_THREADS)
                    H1[i] = fmaxf(PP[i] + PP[PIP * TB + i] + PP[2 * PIP * TB + i] + PP[3 * PIP * TB + i] + SV[SV_BPI2 + (i >> 4)], 0.f);
                __syncthreads();                                  // H1 complete (and PP consumed)
#ifdef AZG_TC_POLICY_PROF
                TC_STAMP();
#endif
                for (int ch = 0; ch < 2 && live; ch++) {
                    const float* W = acquire();
#ifdef AZG_TC_POLICY_PROF
                    TC_STAMP();
#endif
                    {
                        const int k0 = 41 * ch + 6 * slice, k1 = min(k0 + 6, ch == 0 ? 41 : 81);   // every K slice has rows in both chunks: no half-idle warps
#pragma unroll 2
                        for (int k = k0; k < k1; k++) fma_row(W + (k - 41 * ch) * PIP + 8 * o8, *reinterpret_cast<const float4*>(H1 + k * TB + 4 * lq));
                    }
                    release();
                }
#ifdef AZG_TC_POLICY_PROF
                TC_STAMP();
#endif
                if (warp == TC_THREADS / 32 - 1) {                   // the last entry of every slot has been released too: every named barrier is back at a fresh generation
                    asm volatile("bar.sync 1, 384;" ::: "memory"); asm volatile("bar.sync 2, 384;" ::: "memory");
                    asm volatile("bar.sync 3, 384;" ::: "memory"); asm volatile("bar.sync 4, 384;" ::: "memory");
                }
                __syncwarp();
                flush();
#ifdef AZG_TC_POLICY_PROF
                TC_STAMP();
#endif
                __syncthreads();
#ifdef AZG_TC_POLICY_PROF
                TC_STAMP();
#endif
                for (int i = t; i < PIP * TB; i += TC_THREADS) LG[(i & 15) * LGP + (i >> 4)] = PP[i] + PP[PIP * TB + i] + PP[2 * PIP * TB + i] + PP[3 * PIP * TB + i] + SV[SV_BPI4 + (i >> 4)];
                __syncthreads();
#ifdef AZG_TC_POLICY_PROF
                TC_STAMP();
#endif
                {   // masked softmax: where(valid, logits, -1e8) -> log_softmax -> exp (SplendorNNet.py:404,440; GenericNNetWrapper.py:119)
                    const int sl = warp, slot = slot_of[sl];
                    if (slot >= 0) {
                        float l[3]; float mx = -INFINITY;
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int a = lane + 32 * k;
                            const bool valid = a < A && (mw[k] >> lane & 1);
                            l[k] = a < A ? (valid ? LG[sl * LGP + a] : -1e8f) : -INFINITY;
                            mx = fmaxf(mx, l[k]);
                        }
                        mx = warp_max_f32(mx);
                        float sum = 0.f;
#pragma unroll
                        for (int k = 0; k < 3; k++) sum += expf(l[k] - mx);
                        sum = warp_sum_f32(sum);
                        const float lse = logf(sum);
#pragma unroll
                        for (int k = 0; k < 3; k++) { const int a = lane + 32 * k; if (a < A) pi_out[(size_t)slot * A + a] = expf(l[k] - mx - lse); }
                    }
                }
                __syncthreads();                                  // head activations and the ring are dead: value block's images
                TC_STAMP();   /* b7 (b == 1 only): policy head done */
                if (t == 0) {
                    load(bars, B_WE, ESTG, IMGb + I.we[2], TC_WE_BYTES);
                    load(bars, B_WP2, WRING + 32768, IMGb + I.wp[2], 16384);
                    load(bars, B_WP3, WRING + 49152, IMGb + I.wp[2] + 4096, 16384);
                }
                __syncwarp();
            }
        }
        // ---------------- value head: Linear(392 -> np) + ReLU, Linear(np -> np), tanh ----------------
        {
            const float* X0 = reinterpret_cast<const float*>(ESTG);
            ph.wait(bars, B_V2);
            __syncthreads();
            if (tile + (int)gridDim.x < ntiles) {                // next tile's boards -> registers; they are stored to shared memory after this tile's last barrier
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = t + i * TC_THREADS;
                    if (k < TB * 98) { const int s = k / 98, w = k - s * 98, slot = slot_next[s]; rb[i] = slot >= 0 ? reinterpret_cast<const uint32_t*>(boards + (size_t)slot * bstride)[w] : 0u; }
                }
            }
            const float* W = reinterpret_cast<const float*>(WRING);
            if (t < 384) {
                const int s = t & 15, kp = t >> 4;               // 24 token slices x 16 leaves
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int i = kp; i < NV; i += 24) {
                    const float* xp = X0 + i * HLD + s;
#pragma unroll
                    for (int f = 0; f < 7; f++) {
                        const float4 w4 = *reinterpret_cast<const float4*>(W + (i * 7 + f) * 4);
                        const float x = xp[f * HFP];
                        a0 = fmaf(w4.x, x, a0); a1 = fmaf(w4.y, x, a1); a2 = fmaf(w4.z, x, a2); a3 = fmaf(w4.w, x, a3);
                    }
                }
                float* VP = reinterpret_cast<float*>(ESTG + HEAD_X0);   // per-slice partial sums [24][4][16], summed in a fixed order below
                VP[(kp * 4 + 0) * TB + s] = a0; VP[(kp * 4 + 1) * TB + s] = a1; VP[(kp * 4 + 2) * TB + s] = a2; VP[(kp * 4 + 3) * TB + s] = a3;
            }
            __syncthreads();
            if (t < 4 * TB) {
                const float* VP = reinterpret_cast<const float*>(ESTG + HEAD_X0);
                float a = 0.f;
                for (int kp = 0; kp < 24; kp++) a += VP[kp * 4 * TB + t];
                VH[t] = a;
            }
            __syncthreads();
            if (t < NP * TB) {
                const int o = t / TB, sl = t - o * TB, slot = slot_of[sl];
                float a = SV[SV_BV4 + o];
#pragma unroll
                for (int i = 0; i < NP; i++) a = fmaf(SV[SV_V4 + o * 4 + i], fmaxf(VH[i * TB + sl] + SV[SV_BV2 + i], 0.f), a);
                if (slot >= 0) v_out[(size_t)slot * NP + o] = tanhf(a);
            }
        }
        __syncthreads();                                          // tile done: every shared region may be reused
        TC_STAMP();   /* tile end */
    }
    tc_fence_before(); __syncthreads();
    if (prof && t == 0 && blockIdx.x < 160) {
        long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); prof[64 + 4 * blockIdx.x + 2] = g;
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); prof[64 + 4 * blockIdx.x + 3] = smid;
    }
    if (warp == 0) tmem_dealloc<512>(tm);
}

}  // namespace azg

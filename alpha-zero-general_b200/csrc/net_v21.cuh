// net_v21.cuh -- batched forward of AbaloneNNet version 21 (abalone/AbaloneNNet.py:117-156 layers, :173-202 forward;
// torchvision InvertedResidual 24->48->24, kernel 3, no SE, ReLU) in eval mode, replacing onnxruntime's
// InferenceSession.run behind GenericNNetWrapper.predict / predict_server (GenericNNetWrapper.py:94-157) for Abalone.
//
//   x[B,9,9,4] -> channels 0..2 as NCHW -> Conv3x3(3->24)+BN+ReLU -> 4 x { 1x1 24->48 +BN+ReLU, depthwise 3x3 +BN+ReLU,
//   1x1 48->24 +BN, +residual } -> policy: 1x1 24->42 +BN, permuted to [r][q][plane] = 3402 logits, mask, log_softmax,
//   exp | value: 1x1 24->4 +BN+ReLU, flatten 324, cat Linear(6->16)+ReLU of the misc cells, Linear(340->64)+ReLU,
//   Linear(64->2), tanh.   2.1 MFLOP per leaf, 35 862 parameters.
//
// One CTA (512 threads) evaluates 4 leaves (or 2: the tail of a batch is cut into half-size tiles when that fills the last wave of
// CTAs better, see v21_plan); the three activation planes [24|48|48][4 x 81] stay in shared memory
// (156 KB) next to ALL convolution weights (53 KB, copied once per CTA; only the 340 x 64 value matrix stays in L1/L2); the
// dense 3402-wide logits of the tile overwrite the two 48-channel planes once the trunk is done. The layers are
// register-tiled so that shared-memory wavefronts, not FMAs, stop being the limit: a 1x1 task is one board position
// in all four leaves x 4..14 output channels (lanes = consecutive positions, weights = warp-uniform 128-bit broadcasts),
// a 3x3 task is one board row (9 outputs from 27 inputs held in registers, lane stride 9 words = conflict-free).
// The summation order of every convolution output is the reference's (k ascending). BatchNorm folded on the host.
// fp32 on CUDA cores (parity bar 1e-5).
#pragma once
#include "common.cuh"

namespace azg {

constexpr int V21_THREADS = 512;          // one CTA per SM, up to 128 registers per thread for the register tiles
constexpr int V21_TB = 4;
constexpr int V21_N = V21_TB * 81;        // positions per tile
constexpr int V21_A = 3402, V21_MW = 107;

struct V21Layout {
    int wf, bf;                                         // first conv: [(c*9+t)][24], [24]
    struct Blk { int we, be, wd, bd, wp, bp; } blk[4];  // [24][48],[48] | [9][48],[48] | [48][24],[24]
    int wm, bm, wpi, bpi, wvc, bvc, f1, f1b, f2, f2b;   // [6][16],[16] | [24][42],[42] | [24][4],[4] | [340][64],[64] | [2][64],[2]
    int total;
};
inline V21Layout v21_layout() {
    V21Layout L; int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
    L.wf = take(27 * 24); L.bf = take(24);
    for (int i = 0; i < 4; i++) { auto& B = L.blk[i]; B.we = take(24 * 48); B.be = take(48); B.wd = take(9 * 48); B.bd = take(48); B.wp = take(48 * 24); B.bp = take(24); }
    L.wm = take(6 * 16); L.bm = take(16); L.wpi = take(24 * 42); L.bpi = take(42); L.wvc = take(24 * 4); L.bvc = take(4);
    L.f1 = take(340 * 64); L.f1b = take(64); L.f2 = take(2 * 64); L.f2b = take(2);
    L.total = o; return L;
}
inline size_t v21_src_floats() {
    return (size_t)24 * 27 + 96 + 4 * ((48 * 24 + 192) + (48 * 9 + 192) + (24 * 48 + 96)) + (96 + 16) + (42 * 24 + 168) + (4 * 24 + 16) + (64 * 340 + 64 + 128 + 2);
}
// Host: fold BN and transpose to K-major. `src` = state_dict tensors in V21_TENSOR_ORDER (nnet.py).
inline void v21_prepare(const float* src, const V21Layout& L, float* dst) {
    const float* p = src;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    for (int i = 0; i < L.total; i++) dst[i] = 0.f;
    auto conv = [&](int cout, int kin, int w_off, int b_off) {              // weight [cout][kin] (+BN cout) -> [kin][cout], bias
        const float *W = take((size_t)cout * kin), *g = take(cout), *b = take(cout), *m = take(cout), *v = take(cout);
        for (int o = 0; o < cout; o++) {
            const float s = g[o] / sqrtf(v[o] + 1e-5f);
            dst[b_off + o] = b[o] - m[o] * s;
            for (int k = 0; k < kin; k++) dst[w_off + k * cout + o] = W[o * kin + k] * s;
        }
    };
    conv(24, 27, L.wf, L.bf);
    for (int i = 0; i < 4; i++) { const auto& B = L.blk[i]; conv(48, 24, B.we, B.be); conv(48, 9, B.wd, B.bd); conv(24, 48, B.wp, B.bp); }
    { const float *W = take(16 * 6), *b = take(16); for (int o = 0; o < 16; o++) { dst[L.bm + o] = b[o]; for (int k = 0; k < 6; k++) dst[L.wm + k * 16 + o] = W[o * 6 + k]; } }
    conv(42, 24, L.wpi, L.bpi);
    conv(4, 24, L.wvc, L.bvc);
    { const float *W = take((size_t)64 * 340), *b = take(64); for (int o = 0; o < 64; o++) { dst[L.f1b + o] = b[o]; for (int k = 0; k < 340; k++) dst[L.f1 + k * 64 + o] = W[o * 340 + k]; } }
    { const float *W = take(2 * 64), *b = take(2); for (int o = 0; o < 2; o++) { dst[L.f2b + o] = b[o]; for (int k = 0; k < 64; k++) dst[L.f2 + o * 64 + k] = W[o * 64 + k]; } }
}

// out[o][n] = act(bias[o] + sum_k W[k][o] * in[k][n]) (+ out if RES) over the N = 81*TB positions of the tile; task = (group of OG
// outputs) x (TB consecutive positions: one 128-bit (TB = 4) or 64-bit (TB = 2) activation load per input channel)
template <int TB> struct V21Vec;
template <> struct V21Vec<4> { typedef float4 T; };
template <> struct V21Vec<2> { typedef float2 T; };
template <int TB, int CIN, int COUT, int OG, int ACT, bool RES>
__device__ __forceinline__ void conv1x1(const float* W, const float* bias, const float* in, float* out) {
    static_assert(COUT % OG == 0 && OG % 4 == 0 && COUT % 4 == 0, "128-bit weight broadcasts");
    typedef typename V21Vec<TB>::T VT;
    constexpr int N = TB * 81, TASKS = (COUT / OG) * 81;
    for (int t = threadIdx.x; t < TASKS; t += V21_THREADS) {
        const int og = t / 81, nq = t - og * 81, o0 = OG * og;
        uint64_t acc2[OG / 2][TB];                                 // FFMA2 pairs: outputs (2 jp, 2 jp + 1) x one position
#pragma unroll
        for (int j = 0; j < OG / 2; j++)
#pragma unroll
            for (int l = 0; l < TB; l++) acc2[j][l] = 0ull;
        const float* xin = in + TB * nq; const float* wk = W + o0;
#pragma unroll 4
        for (int k = 0; k < CIN; k++) {
            const VT xv = *reinterpret_cast<const VT*>(xin + k * N);
            float x[TB];
            memcpy(x, &xv, sizeof(xv));
            uint64_t xd[TB];
#pragma unroll
            for (int l = 0; l < TB; l++) xd[l] = pk2(x[l], x[l]);
#pragma unroll
            for (int j4 = 0; j4 < OG / 4; j4++) {
                const float4 w4 = *reinterpret_cast<const float4*>(wk + k * COUT + 4 * j4);
                const uint64_t w01 = pk2(w4.x, w4.y), w23 = pk2(w4.z, w4.w);
#pragma unroll
                for (int l = 0; l < TB; l++) { fma2(acc2[2 * j4][l], w01, xd[l]); fma2(acc2[2 * j4 + 1][l], w23, xd[l]); }
            }
        }
        float acc[OG][TB];
#pragma unroll
        for (int j = 0; j < OG / 2; j++)
#pragma unroll
            for (int l = 0; l < TB; l++) upk2(acc2[j][l], acc[2 * j][l], acc[2 * j + 1][l]);
#pragma unroll
        for (int j = 0; j < OG; j++) {
            const float b = bias[o0 + j];
            VT* o = reinterpret_cast<VT*>(out + (o0 + j) * N + TB * nq);
            float r[TB], v[TB];
            if (RES) { const VT rv = *o; memcpy(r, &rv, sizeof(rv)); }
#pragma unroll
            for (int l = 0; l < TB; l++) {
                v[l] = acc[j][l] + b;
                if (ACT) v[l] = fmaxf(v[l], 0.f);
                if (RES) v[l] += r[l];
            }
            VT ov; memcpy(&ov, v, sizeof(ov));
            *o = ov;
        }
    }
}

// Shared memory of a CTA (sized for the 4-leaf tile): activations | VC | VH | VP | WS. WS mirrors the parameter blob up to (not including) f1.
constexpr int V21_WS_OFF = 120 * V21_N + V21_TB * (340 + 64) + 8 * V21_TB * 64;       // floats
inline size_t v21_smem_bytes() { return sizeof(float) * (size_t)(V21_WS_OFF + v21_layout().f1); }

// Tiling of a batch of n leaves on n_sm SMs (one CTA per SM at a time, dispatched in blockIdx order): 4-leaf tiles, except that the
// leaves left over after the last full round of 4-leaf tiles go into 2-leaf tiles when those fit in ONE round (a 2-leaf tile takes
// ~0.6 of a 4-leaf one: 2048 leaves on 148 SMs = 3 rounds + 0.6 instead of 4 rounds).
struct V21Plan { int n_big, n_small; };
inline V21Plan v21_plan(int n, int n_sm) {
    V21Plan p; const int full = n / (4 * n_sm) * n_sm, rest = n - 4 * full;
    if ((rest + 1) / 2 <= n_sm) { p.n_big = full; p.n_small = (rest + 1) / 2; }
    else { p.n_big = full + (rest + 3) / 4; p.n_small = 0; }
    return p;
}

// One tile of TB leaves (TB = 4 or 2). slot_of[TB], red[32] in static shared memory; WS already holds the weights.
template <int TB>
__device__ __forceinline__ void v21_tile(const float* __restrict__ P, const V21Layout& L, float* smem, const int* slot_of, float* red,
                                         const int8_t* boards, int bstride, const uint32_t* masks, float* pi_out, float* v_out) {
    constexpr int N = TB * 81, A = V21_A, MW = V21_MW, ROWS = TB * 9;
    constexpr int WPL = (V21_THREADS / 32) / TB, SW = (MW + WPL - 1) / WPL;           // softmax: warps per leaf, mask words per warp
    static_assert(SW <= 32 && TB * 64 <= V21_THREADS && 3 * 81 <= 256, "task maps");
    float* X = smem; float* E = X + 24 * N; float* D = E + 48 * N; float* VC = D + 48 * N;     // VC [TB][340]: value features + meta
    float* VH = VC + TB * 340; float* VP = VH + TB * 64; const float* WS = smem + V21_WS_OFF;  // VP [8][TB][64]: K-slice partials of the value Linear
    float* LG = E;                                                                             // LG [TB][3402] over E and D
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // input planes (channels 0..2) into D, misc cells -> meta embedding Linear(6->16)+ReLU into VC[.][324..339]
    for (int k = t; k < 3 * N; k += V21_THREADS) {
        const int c = k / N, n = k - c * N, l = n / 81, pos = n - l * 81, slot = slot_of[l];
        D[k] = slot >= 0 ? (float)boards[(size_t)slot * bstride + pos * 4 + c] : 0.f;
    }
    if (t < TB * 16) {
        const int l = t >> 4, o = t & 15, slot = slot_of[l];
        float a = WS[L.bm + o];
        if (slot >= 0) for (int j = 0; j < 6; j++) a = fmaf(WS[L.wm + j * 16 + o], (float)boards[(size_t)slot * bstride + j * 4 + 3], a);
        VC[l * 340 + 324 + o] = fmaxf(a, 0.f);
    }
    __syncthreads();
    // first_layer: Conv3x3(3->24) + BN + ReLU; task = (group of 4 outputs) x (leaf, board row): 9 x 4 outputs from 3 x 27 inputs
    for (int k = t; k < 6 * ROWS; k += V21_THREADS) {
        const int og = k / ROWS, m = k - ROWS * og, r = m % 9, o0 = 4 * og;
        float acc[9][4];
#pragma unroll
        for (int q = 0; q < 9; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int dr = 0; dr < 3; dr++) {
                const int rr = r + dr - 1; const bool ok = rr >= 0 && rr < 9;
                float row[9];
#pragma unroll
                for (int q = 0; q < 9; q++) row[q] = ok ? D[c * N + 9 * m + (dr - 1) * 9 + q] : 0.f;
#pragma unroll
                for (int dq = 0; dq < 3; dq++) {
                    const float4 w = *reinterpret_cast<const float4*>(WS + L.wf + (c * 9 + dr * 3 + dq) * 24 + o0);
#pragma unroll
                    for (int q = 0; q < 9; q++) {
                        const int qq = q + dq - 1;
                        if (qq < 0 || qq >= 9) continue;
                        acc[q][0] = fmaf(w.x, row[qq], acc[q][0]); acc[q][1] = fmaf(w.y, row[qq], acc[q][1]);
                        acc[q][2] = fmaf(w.z, row[qq], acc[q][2]); acc[q][3] = fmaf(w.w, row[qq], acc[q][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float b = WS[L.bf + o0 + j];
#pragma unroll
            for (int q = 0; q < 9; q++) X[(o0 + j) * N + 9 * m + q] = fmaxf(acc[q][j] + b, 0.f);
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int b = 0; b < 4; b++) {                                                    // trunk: InvertedResidual x 4
        const V21Layout::Blk B = L.blk[b];
        conv1x1<TB, 24, 48, 8, 1, false>(WS + B.we, WS + B.be, X, E);
        __syncthreads();
        for (int k = t; k < 48 * ROWS; k += V21_THREADS) {                            // depthwise 3x3 + BN + ReLU; task = (channel, leaf, board row)
            const int c = k / ROWS, r = (k - ROWS * c) % 9;
            const float* e = E + 9 * k;                                               // = E + c*N + leaf*81 + r*9
            float w[9], a[9];
#pragma unroll
            for (int tp = 0; tp < 9; tp++) w[tp] = WS[B.wd + tp * 48 + c];
            const float bd = WS[B.bd + c];
#pragma unroll
            for (int q = 0; q < 9; q++) a[q] = bd;
#pragma unroll
            for (int dr = 0; dr < 3; dr++) {
                const int rr = r + dr - 1; const bool ok = rr >= 0 && rr < 9;
                float row[9];
#pragma unroll
                for (int q = 0; q < 9; q++) row[q] = ok ? e[(dr - 1) * 9 + q] : 0.f;
#pragma unroll
                for (int dq = 0; dq < 3; dq++)
#pragma unroll
                    for (int q = 0; q < 9; q++) {
                        const int qq = q + dq - 1;
                        if (qq < 0 || qq >= 9) continue;
                        a[q] = fmaf(w[dr * 3 + dq], row[qq], a[q]);
                    }
            }
#pragma unroll
            for (int q = 0; q < 9; q++) D[9 * k + q] = fmaxf(a[q], 0.f);
        }
        __syncthreads();
        conv1x1<TB, 48, 24, 4, 0, true>(WS + B.wp, WS + B.bp, D, X);                  // project + residual (X holds the block input)
        __syncthreads();
    }
    // heads (both read X only; E and D are free): policy logits on warps 0..7 overwrite E/D; value features and value Linear on warps 8..15
    if (t < 3 * 81) {                                                                 // 1x1 24->42 + BN -> logits[r][q][plane]; task = 14 planes x position
        const int og = t / 81, pos = t - og * 81, o0 = 14 * og;
        uint64_t acc2[7][TB];                                                          // FFMA2 pairs: planes (2 j2, 2 j2 + 1) x one leaf
#pragma unroll
        for (int j = 0; j < 7; j++)
#pragma unroll
            for (int l = 0; l < TB; l++) acc2[j][l] = 0ull;
#pragma unroll 2
        for (int i = 0; i < 24; i++) {
            uint64_t xd[TB];
#pragma unroll
            for (int l = 0; l < TB; l++) { const float x = X[i * N + l * 81 + pos]; xd[l] = pk2(x, x); }
            const float2* w2 = reinterpret_cast<const float2*>(WS + L.wpi + i * 42 + o0);
#pragma unroll
            for (int j2 = 0; j2 < 7; j2++) {
                const float2 w = w2[j2]; const uint64_t wp = pk2(w.x, w.y);
#pragma unroll
                for (int l = 0; l < TB; l++) fma2(acc2[j2][l], wp, xd[l]);
            }
        }
        float acc[14][TB];
#pragma unroll
        for (int j = 0; j < 7; j++)
#pragma unroll
            for (int l = 0; l < TB; l++) upk2(acc2[j][l], acc[2 * j][l], acc[2 * j + 1][l]);
#pragma unroll
        for (int j2 = 0; j2 < 7; j2++) {
            const float b0 = WS[L.bpi + o0 + 2 * j2], b1 = WS[L.bpi + o0 + 2 * j2 + 1];
#pragma unroll
            for (int l = 0; l < TB; l++) *reinterpret_cast<float2*>(LG + l * A + pos * 42 + o0 + 2 * j2) = make_float2(acc[2 * j2][l] + b0, acc[2 * j2 + 1][l] + b1);
        }
    } else if (t >= 256) {                                                            // warps 8..15: the value path, under the policy conv
        if (t < 256 + 81) {                                                           // 1x1 24->4 + BN + ReLU, flattened channel-major
            const int pos = t - 256;
            float acc[4][TB];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int l = 0; l < TB; l++) acc[c][l] = WS[L.bvc + c];
#pragma unroll 4
            for (int i = 0; i < 24; i++) {
                const float4 w = *reinterpret_cast<const float4*>(WS + L.wvc + i * 4);
#pragma unroll
                for (int l = 0; l < TB; l++) {
                    const float x = X[i * N + l * 81 + pos];
                    acc[0][l] = fmaf(w.x, x, acc[0][l]); acc[1][l] = fmaf(w.y, x, acc[1][l]);
                    acc[2][l] = fmaf(w.z, x, acc[2][l]); acc[3][l] = fmaf(w.w, x, acc[3][l]);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int l = 0; l < TB; l++) VC[l * 340 + c * 81 + pos] = fmaxf(acc[c][l], 0.f);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");                                // VC complete (warps 8..15 only; warps 0..7 are in the policy conv)
        // value Linear(340 -> 64): K split in 4 slices of 85, one thread per (slice, output), all leaves; unrolled so that the weight
        // loads (L2: the 87 KB matrix does not fit the 28 KB left to L1) are in flight together
        const int ks = (t - 256) >> 6, j = t & 63, i0 = 85 * ks;
        float a[TB];
#pragma unroll
        for (int l = 0; l < TB; l++) a[l] = 0.f;
#pragma unroll
        for (int ii = 0; ii < 85; ii++) {
            const float w = __ldg(P + L.f1 + (i0 + ii) * 64 + j);
#pragma unroll
            for (int l = 0; l < TB; l++) a[l] = fmaf(w, VC[l * 340 + i0 + ii], a[l]);
        }
#pragma unroll
        for (int l = 0; l < TB; l++) VP[(ks * TB + l) * 64 + j] = a[l];
    }
    __syncthreads();
    // masked softmax over 3402 actions: where(valid, logits, -1e8) -> log_softmax -> exp; WPL warps per leaf, mask words interleaved
    // (warp wv owns words wv, wv+WPL, ...). Mask words are loaded once (lane i holds word wv+WPL*i) and the masked logits stay in registers.
    const int sl = warp / WPL, wv = warp % WPL, slot = slot_of[sl];
    float val[SW];
    {
        const uint32_t mword = (slot >= 0 && wv + WPL * lane < MW) ? __ldg(masks + (size_t)slot * MW + wv + WPL * lane) : 0u;
        const float* lg = LG + sl * A;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < SW; i++) {
            const int k = wv + WPL * i, a = lane + 32 * k;
            const uint32_t m = __shfl_sync(FULL, mword, i);
            val[i] = (k < MW && a < A) ? ((m >> lane & 1) ? lg[a] : -1e8f) : -INFINITY;   // -inf past the end of the action space
            mx = fmaxf(mx, val[i]);
        }
        mx = warp_max_f32(mx);
        if (lane == 0) red[warp] = mx;
    }
    __syncthreads();
    if (t < TB * 64) {                                                                // value Linear: bias + the 4 partials in a fixed order, ReLU
        const int l = t >> 6, j = t & 63;
        float a = __ldg(P + L.f1b + j);
#pragma unroll
        for (int ks = 0; ks < 4; ks++) a += VP[(ks * TB + l) * 64 + j];
        VH[l * 64 + j] = fmaxf(a, 0.f);
    }
    float mx = red[sl * WPL];
#pragma unroll
    for (int i = 1; i < WPL; i++) mx = fmaxf(mx, red[sl * WPL + i]);
    {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < SW; i++) sum += expf(val[i] - mx);
        sum = warp_sum_f32(sum);
        if (lane == 0) red[16 + warp] = sum;
    }
    __syncthreads();
    if (slot >= 0) {
        float tot = red[16 + sl * WPL];
#pragma unroll
        for (int i = 1; i < WPL; i++) tot += red[16 + sl * WPL + i];
        const float lse = logf(tot);
        float* po = pi_out + (size_t)slot * A + lane;
#pragma unroll
        for (int i = 0; i < SW; i++) { const int k = wv + WPL * i; if (k < MW && lane + 32 * k < A) po[32 * k] = expf(val[i] - mx - lse); }
    }
    if (warp < TB * 2) {                                                              // value Linear(64 -> 2), tanh: one warp per (leaf, output)
        const int l = warp >> 1, o = warp & 1, vslot = slot_of[l];
        float a = fmaf(__ldg(P + L.f2 + o * 64 + lane), VH[l * 64 + lane], __ldg(P + L.f2 + o * 64 + 32 + lane) * VH[l * 64 + 32 + lane]);
        a = warp_sum_f32(a);
        if (lane == 0 && vslot >= 0) v_out[(size_t)vslot * 2 + o] = tanhf(a + __ldg(P + L.f2b + o));
    }
}

// boards: int8[.][324] HWC with `bstride` bytes between boards; masks: 107 words per slot; list/count as in k_v80_forward.
// CTAs [0, n_big) take 4 leaves each, CTAs [n_big, gridDim.x) 2 leaves each (v21_plan).
__global__ void __launch_bounds__(V21_THREADS, 1)
k_v21_forward(const float* __restrict__ P, const __grid_constant__ V21Layout L, const int* count_ptr, const int* list,
              const int8_t* boards, int bstride, const uint32_t* masks, float* pi_out, float* v_out, int n_max, int n_big) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int slot_of[4];
    __shared__ float red[32];
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const bool big = (int)blockIdx.x < n_big;
    const int tile0 = big ? 4 * blockIdx.x : 4 * n_big + 2 * ((int)blockIdx.x - n_big), tb = big ? 4 : 2;
    if (tile0 >= count) return;
    const int t = threadIdx.x;
    if (t < 4) { const int j = tile0 + t; slot_of[t] = (t < tb && j < count) ? (list ? list[j] : j) : -1; }
    for (int i = t; i < L.f1 / 4; i += V21_THREADS) reinterpret_cast<float4*>(smem + V21_WS_OFF)[i] = __ldg(reinterpret_cast<const float4*>(P) + i);
    __syncthreads();
    if (big) v21_tile<4>(P, L, smem, slot_of, red, boards, bstride, masks, pi_out, v_out);
    else     v21_tile<2>(P, L, smem, slot_of, red, boards, bstride, masks, pi_out, v_out);
}

}  // namespace azg

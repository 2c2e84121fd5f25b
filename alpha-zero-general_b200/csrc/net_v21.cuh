// net_v21.cuh -- batched forward of AbaloneNNet version 21 (abalone/AbaloneNNet.py:117-156 layers, :173-202 forward;
// torchvision InvertedResidual 24->48->24, kernel 3, no SE, ReLU) in eval mode, replacing onnxruntime's
// InferenceSession.run behind GenericNNetWrapper.predict / predict_server (GenericNNetWrapper.py:94-157) for Abalone.
//
//   x[B,9,9,4] -> channels 0..2 as NCHW -> Conv3x3(3->24)+BN+ReLU -> 4 x { 1x1 24->48 +BN+ReLU, depthwise 3x3 +BN+ReLU,
//   1x1 48->24 +BN, +residual } -> policy: 1x1 24->42 +BN, permuted to [r][q][plane] = 3402 logits, mask, log_softmax,
//   exp | value: 1x1 24->4 +BN+ReLU, flatten 324, cat Linear(6->16)+ReLU of the misc cells, Linear(340->64)+ReLU,
//   Linear(64->2), tanh.   2.1 MFLOP per leaf, 35 862 parameters.
//
// One CTA (256 threads) evaluates 4 leaves; the three activation planes [24|48|48][4 x 81] stay in shared memory
// (156 KB), the dense 3402-wide logits of the tile overwrite the two 48-channel planes once the trunk is done.
// Lanes run over consecutive board positions, output-channel groups of 6 are warp-uniform (broadcast weight loads
// from L1/L2; the whole net is 144 KB). BatchNorm folded on the host. fp32 on CUDA cores (parity bar 1e-5).
#pragma once
#include "common.cuh"

namespace azg {

constexpr int V21_THREADS = 1024;         // 32 warps per (single-CTA) SM: the layers are latency-bound weight broadcasts, more warps hide it
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

// out[o][n] = act(bias[o] + sum_k W[k][o] * in[k][n]) (+ out[o][n] if RES); tasks = (group of 6 outputs) x (position n).
template <int CIN, int COUT, int ACT, bool RES>
__device__ __forceinline__ void conv1x1(const float* __restrict__ W, const float* __restrict__ bias, const float* in, float* out) {
    static_assert(COUT % 6 == 0, "output channels in groups of 6");
    for (int t = threadIdx.x; t < (COUT / 6) * V21_N; t += V21_THREADS) {
        const int og = t / V21_N, n = t - og * V21_N, o0 = 6 * og;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int k = 0; k < CIN; k++) {
            const float x = in[k * V21_N + n];
            const float2 w0 = __ldg(reinterpret_cast<const float2*>(W + k * COUT + o0)), w1 = __ldg(reinterpret_cast<const float2*>(W + k * COUT + o0 + 2)),
                         w2 = __ldg(reinterpret_cast<const float2*>(W + k * COUT + o0 + 4));
            acc[0] = fmaf(w0.x, x, acc[0]); acc[1] = fmaf(w0.y, x, acc[1]); acc[2] = fmaf(w1.x, x, acc[2]);
            acc[3] = fmaf(w1.y, x, acc[3]); acc[4] = fmaf(w2.x, x, acc[4]); acc[5] = fmaf(w2.y, x, acc[5]);
        }
#pragma unroll
        for (int j = 0; j < 6; j++) {
            float v = acc[j] + __ldg(bias + o0 + j);
            if (ACT) v = fmaxf(v, 0.f);
            if (RES) v += out[(o0 + j) * V21_N + n];
            out[(o0 + j) * V21_N + n] = v;
        }
    }
}

constexpr size_t v21_smem_bytes() { return sizeof(float) * (size_t)(120 * V21_N + V21_TB * (340 + 64 + 8)); }

// boards: int8[.][324] HWC with `bstride` bytes between boards; masks: 107 words per slot; list/count as in k_v80_forward.
__global__ void __launch_bounds__(V21_THREADS, 1)
k_v21_forward(const float* __restrict__ P, const __grid_constant__ V21Layout L, const int* count_ptr, const int* list,
              const int8_t* boards, int bstride, const uint32_t* masks, float* pi_out, float* v_out, int n_max) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TB = V21_TB, N = V21_N, A = V21_A, MW = V21_MW;
    float* X = smem; float* E = X + 24 * N; float* D = E + 48 * N; float* VC = D + 48 * N;     // VC [TB][340]: value features + meta
    float* VH = VC + TB * 340; float* LG = E;                                                  // LG [TB][3402] over E and D
    __shared__ int slot_of[TB];
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int tile0 = blockIdx.x * TB;
    if (tile0 >= count) return;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t < TB) { const int j = tile0 + t; slot_of[t] = j < count ? (list ? list[j] : j) : -1; }
    __syncthreads();
    // input planes (channels 0..2) into D, misc cells -> meta embedding Linear(6->16)+ReLU into VC[.][324..339]
    for (int k = t; k < 3 * N; k += V21_THREADS) {
        const int c = k / N, n = k - c * N, l = n / 81, pos = n - l * 81, slot = slot_of[l];
        D[k] = slot >= 0 ? (float)boards[(size_t)slot * bstride + pos * 4 + c] : 0.f;
    }
    if (t < TB * 16) {
        const int l = t >> 4, o = t & 15, slot = slot_of[l];
        float a = __ldg(P + L.bm + o);
        if (slot >= 0) for (int j = 0; j < 6; j++) a = fmaf(__ldg(P + L.wm + j * 16 + o), (float)boards[(size_t)slot * bstride + j * 4 + 3], a);
        VC[l * 340 + 324 + o] = fmaxf(a, 0.f);
    }
    __syncthreads();
    // first_layer: Conv3x3(3->24) + BN + ReLU; tasks = (group of 6 outputs) x position
    for (int k = t; k < 4 * N; k += V21_THREADS) {
        const int og = k / N, n = k - og * N, l = n / 81, pos = n - l * 81, r = pos / 9, q = pos - 9 * r, o0 = 6 * og;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int tp = 0; tp < 9; tp++) {
                const int rr = r + tp / 3 - 1, qq = q + tp % 3 - 1;
                if (rr < 0 || rr >= 9 || qq < 0 || qq >= 9) continue;
                const float x = D[c * N + l * 81 + rr * 9 + qq];
                const float* w = P + L.wf + (c * 9 + tp) * 24 + o0;
#pragma unroll
                for (int j = 0; j < 6; j++) acc[j] = fmaf(__ldg(w + j), x, acc[j]);
            }
#pragma unroll
        for (int j = 0; j < 6; j++) X[(o0 + j) * N + n] = fmaxf(acc[j] + __ldg(P + L.bf + o0 + j), 0.f);
    }
    __syncthreads();
    for (int b = 0; b < 4; b++) {                                                    // trunk: InvertedResidual x 4
        const V21Layout::Blk B = L.blk[b];
        conv1x1<24, 48, 1, false>(P + B.we, P + B.be, X, E);
        __syncthreads();
        for (int k = t; k < 48 * N; k += V21_THREADS) {                               // depthwise 3x3 + BN + ReLU
            const int c = k / N, n = k - c * N, l = n / 81, pos = n - l * 81, r = pos / 9, q = pos - 9 * r;
            float a = __ldg(P + B.bd + c);
#pragma unroll
            for (int tp = 0; tp < 9; tp++) {
                const int rr = r + tp / 3 - 1, qq = q + tp % 3 - 1;
                if (rr < 0 || rr >= 9 || qq < 0 || qq >= 9) continue;
                a = fmaf(__ldg(P + B.wd + tp * 48 + c), E[c * N + l * 81 + rr * 9 + qq], a);
            }
            D[k] = fmaxf(a, 0.f);
        }
        __syncthreads();
        conv1x1<48, 24, 0, true>(P + B.wp, P + B.bp, D, X);                           // project + residual (X holds the block input)
        __syncthreads();
    }
    // value features first (they read X only; E and D are free), then the policy logits overwrite E/D
    for (int k = t; k < 4 * N; k += V21_THREADS) {                                    // 1x1 24->4 + BN + ReLU, flattened channel-major
        const int c = k / N, n = k - c * N, l = n / 81, pos = n - l * 81;
        float a = __ldg(P + L.bvc + c);
#pragma unroll 8
        for (int i = 0; i < 24; i++) a = fmaf(__ldg(P + L.wvc + i * 4 + c), X[i * N + n], a);
        VC[l * 340 + c * 81 + pos] = fmaxf(a, 0.f);
    }
    for (int k = t; k < 7 * N; k += V21_THREADS) {                                    // 1x1 24->42 + BN -> logits[r][q][plane]
        const int og = k / N, n = k - og * N, l = n / 81, pos = n - l * 81, o0 = 6 * og;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int i = 0; i < 24; i++) {
            const float x = X[i * N + n]; const float* w = P + L.wpi + i * 42 + o0;
#pragma unroll
            for (int j = 0; j < 6; j++) acc[j] = fmaf(__ldg(w + j), x, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 6; j++) LG[l * A + pos * 42 + o0 + j] = acc[j] + __ldg(P + L.bpi + o0 + j);
    }
    __syncthreads();
    if (t < TB * 64) {   // value Linear(340 -> 64) + ReLU: one thread per (leaf, output)
        const int l = t >> 6, j = t & 63;
        float a = __ldg(P + L.f1b + j);
        for (int i = 0; i < 340; i++) a = fmaf(__ldg(P + L.f1 + i * 64 + j), VC[l * 340 + i], a);
        VH[l * 64 + j] = fmaxf(a, 0.f);
    }
    // masked softmax over 3402 actions: where(valid, logits, -1e8) -> log_softmax -> exp; one warp per leaf
    if (warp < TB) {
        const int l = warp, slot = slot_of[l];
        if (slot >= 0) {
            const uint32_t* mk = masks + (size_t)slot * MW; const float* lg = LG + l * A;
            float mx = -INFINITY;
            for (int k = 0; k < MW; k++) { const int a = lane + 32 * k; if (a < A) mx = fmaxf(mx, (mk[k] >> lane & 1) ? lg[a] : -1e8f); }
            mx = warp_max_f32(mx);
            float sum = 0.f;
            for (int k = 0; k < MW; k++) { const int a = lane + 32 * k; if (a < A) sum += expf(((mk[k] >> lane & 1) ? lg[a] : -1e8f) - mx); }
            sum = warp_sum_f32(sum);
            const float lse = logf(sum);
            for (int k = 0; k < MW; k++) { const int a = lane + 32 * k; if (a < A) pi_out[(size_t)slot * A + a] = expf(((mk[k] >> lane & 1) ? lg[a] : -1e8f) - mx - lse); }
        }
    }
    __syncthreads();
    if (t < TB * 2) {                                                                 // value Linear(64 -> 2), tanh
        const int l = t >> 1, o = t & 1, slot = slot_of[l];
        float a = __ldg(P + L.f2b + o);
        for (int j = 0; j < 64; j++) a = fmaf(__ldg(P + L.f2 + o * 64 + j), VH[l * 64 + j], a);
        if (slot >= 0) v_out[(size_t)slot * 2 + o] = tanhf(a);
    }
}

}  // namespace azg

// net_v89.cuh -- batched forward of SantoriniNNet version 89 (santorini/SantoriniNNet.py:70-84 SimpleResBlock,
// :16-40 SimpleHead, :194-217 layers, :273-279 forward) in eval mode, replacing onnxruntime's InferenceSession.run
// behind GenericNNetWrapper.predict / predict_server (GenericNNetWrapper.py:94-157) for Santorini without gods.
//
//   x[B,5,5,3] -> channels 0,1 (workers, levels) as NCHW -> Conv3x3(2->64)+BN+ReLU -> 5 x { Conv3x3(64->64)+BN+ReLU,
//   Conv3x3(64->64)+BN, +residual, ReLU } -> policy head: Conv1x1(64->2)+BN+ReLU, flatten 50, Linear(50->162), mask,
//   log_softmax, exp | value head: Conv1x1(64->1)+BN+ReLU, flatten 25, Linear(25->64)+ReLU, Linear(64->2), tanh.
//   18.5 MFLOP per leaf, 381 454 parameters.
//
// One CTA (256 threads) evaluates a tile of 6 leaves with both activation planes resident in shared memory as
// [channel][leaf][7x7 zero-padded grid], so the 3x3 taps are plain offsets. Warp w owns output channels 8w..8w+7
// (weights are warp-uniform broadcast loads), lane = (leaf, output row): 8 channels x 5 positions of accumulators,
// 360 FFMA per input channel against 21 activation + 18 weight loads. BatchNorm is folded on the host; the 147 KB of
// weights per convolution stream global -> shared with cp.async in 8-input-channel chunks through the same 3-deep
// ring as the V80 kernel (net_v80.cuh: WPipe). fp32 on CUDA cores for the same parity reason as V80.
#pragma once
#include "common.cuh"
#include "net_v80.cuh"

namespace azg {

constexpr int V89_THREADS = 256;
constexpr int V89_TB = 6;                 // leaves per CTA: 6 x 5 rows = 30 lanes
constexpr int V89_C = 64;                 // trunk width
constexpr int V89_NCONV = 11;             // first layer + 5 blocks x 2
constexpr int V89_GRID = 49;              // 7 x 7 zero-padded positions per leaf
constexpr int V89_A = 162, V89_AP = 164;  // policy outputs, padded to a multiple of 4
constexpr int V89_MAXCHUNK = 96;

struct V89Layout {
    int conv[V89_NCONV], cbias[V89_NCONV];          // conv i: [cin][9 taps][64 out] (BN scale folded), bias [64]
    int pi_w, pi_b, pi_fc, pi_fcb;                  // [64 c][2], [2], [50 k][164], [164]
    int v_w, v_b, v_fc1, v_fc1b, v_fc2, v_fc2b;     // [64], [1], [25 k][64], [64], [2][64], [2]
    int total;
};
struct V89Chunks { int off[V89_MAXCHUNK]; int n[V89_MAXCHUNK]; int count; };

inline V89Layout v89_layout() {
    V89Layout L; int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
    for (int i = 0; i < V89_NCONV; i++) { L.conv[i] = take((i == 0 ? 2 : V89_C) * 9 * V89_C); L.cbias[i] = take(V89_C); }
    L.pi_w = take(V89_C * 2); L.pi_b = take(2); L.pi_fc = take(50 * V89_AP); L.pi_fcb = take(V89_AP);
    L.v_w = take(V89_C); L.v_b = take(1); L.v_fc1 = take(25 * 64); L.v_fc1b = take(64); L.v_fc2 = take(2 * 64); L.v_fc2b = take(2);
    L.total = o; return L;
}
inline V89Chunks v89_chunks(const V89Layout& L) {
    V89Chunks c; c.count = 0;
    auto add = [&](int off, int n) { c.off[c.count] = off; c.n[c.count] = n; c.count++; };
    add(L.conv[0], 2 * 9 * V89_C);
    for (int i = 1; i < V89_NCONV; i++) for (int j = 0; j < V89_C / 8; j++) add(L.conv[i] + j * 8 * 9 * V89_C, 8 * 9 * V89_C);
    return c;
}
inline size_t v89_src_floats() {
    return (size_t)64 * 2 * 9 + 4 * 64 + 10 * ((size_t)64 * 64 * 9 + 4 * 64) + (2 * 64 + 8 + 162 * 50 + 162) + (64 + 4 + 64 * 25 + 64 + 2 * 64 + 2);
}
// Host: fold BN and re-lay the weights. `src` = state_dict tensors in V89_TENSOR_ORDER (nnet.py), `dst` = prepared blob.
inline void v89_prepare(const float* src, const V89Layout& L, float* dst) {
    const float* p = src;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    for (int i = 0; i < L.total; i++) dst[i] = 0.f;
    for (int i = 0; i < V89_NCONV; i++) {
        const int cin = i == 0 ? 2 : V89_C;
        const float *W = take((size_t)V89_C * cin * 9), *g = take(V89_C), *b = take(V89_C), *m = take(V89_C), *v = take(V89_C);
        for (int o = 0; o < V89_C; o++) {
            const float s = g[o] / sqrtf(v[o] + 1e-5f);
            dst[L.cbias[i] + o] = b[o] - m[o] * s;
            for (int c = 0; c < cin; c++) for (int t = 0; t < 9; t++) dst[L.conv[i] + (c * 9 + t) * V89_C + o] = W[(o * cin + c) * 9 + t] * s;
        }
    }
    {   // policy head
        const float *W = take(2 * 64), *g = take(2), *b = take(2), *m = take(2), *v = take(2), *fc = take((size_t)162 * 50), *fb = take(162);
        for (int o = 0; o < 2; o++) { const float s = g[o] / sqrtf(v[o] + 1e-5f); dst[L.pi_b + o] = b[o] - m[o] * s; for (int c = 0; c < 64; c++) dst[L.pi_w + c * 2 + o] = W[o * 64 + c] * s; }
        for (int o = 0; o < 162; o++) { dst[L.pi_fcb + o] = fb[o]; for (int k = 0; k < 50; k++) dst[L.pi_fc + k * V89_AP + o] = fc[o * 50 + k]; }
    }
    {   // value head
        const float *W = take(64), *g = take(1), *b = take(1), *m = take(1), *v = take(1), *f1 = take(64 * 25), *b1 = take(64), *f2 = take(2 * 64), *b2 = take(2);
        const float s = g[0] / sqrtf(v[0] + 1e-5f); dst[L.v_b] = b[0] - m[0] * s;
        for (int c = 0; c < 64; c++) dst[L.v_w + c] = W[c] * s;
        for (int j = 0; j < 64; j++) { dst[L.v_fc1b + j] = b1[j]; for (int k = 0; k < 25; k++) dst[L.v_fc1 + k * 64 + j] = f1[j * 25 + k]; }
        for (int o = 0; o < 2; o++) { dst[L.v_fc2b + o] = b2[o]; for (int j = 0; j < 64; j++) dst[L.v_fc2 + o * 64 + j] = f2[o * 64 + j]; }
    }
}

struct WPipe89 {                            // same protocol as WPipe (net_v80.cuh) for this kernel's thread count / chunk table
    const float* P; float* buf; const V89Chunks* ck; int issued, used;
    __device__ __forceinline__ void issue() {
        if (issued < ck->count) {
            const float4* src = reinterpret_cast<const float4*>(P + ck->off[issued]);
            float* dst = buf + (issued % V80_NBUF) * V80_CH; const int n4 = ck->n[issued] >> 2;
            for (int i = threadIdx.x; i < n4; i += V89_THREADS) cp_async16(dst + 4 * i, src + i);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        issued++;
    }
    __device__ __forceinline__ const float* acquire() {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(V80_NBUF - 2));
        __syncthreads();
        issue();
        return buf + (used++ % V80_NBUF) * V80_CH;
    }
};

// out[o][leaf][y+1][x+1] = relu(bias[o] + sum_{c,ky,kx} W[c][ky*3+kx][o] * in[c][leaf][y+ky][x+kx] (+ res[o][..])) on the padded grids.
template <int CIN, bool RES>
__device__ __forceinline__ void conv3x3(WPipe89& wp, const float* __restrict__ bias, const float* in, float* out, int ldq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool on = lane < V89_TB * 5;
    const int l = on ? lane / 5 : 0, y = on ? lane % 5 : 0;
    float acc[8][5];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int x = 0; x < 5; x++) acc[j][x] = 0.f;
    constexpr int NCH = CIN >= 8 ? CIN / 8 : 1, CPC = CIN >= 8 ? 8 : CIN;
    for (int ch = 0; ch < NCH; ch++) {
        const float* W = wp.acquire();                       // [CPC][9][64]; the barrier also orders the previous layer's writes
#pragma unroll 2
        for (int cl = 0; cl < CPC; cl++) {
            const float* ip = in + (ch * CPC + cl) * ldq + l * V89_GRID + y * 7;
            float r[3][7];
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int k = 0; k < 7; k++) r[ky][k] = ip[ky * 7 + k];
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const float4 wa = *reinterpret_cast<const float4*>(W + (cl * 9 + t) * V89_C + 8 * warp);
                const float4 wb = *reinterpret_cast<const float4*>(W + (cl * 9 + t) * V89_C + 8 * warp + 4);
                const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int x = 0; x < 5; x++) {
                    const float v = r[t / 3][x + t % 3];
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j][x] = fmaf(w[j], v, acc[j][x]);
                }
            }
        }
    }
    if (on) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int o = 8 * warp + j; const float b = __ldg(bias + o);
            float* op = out + o * ldq + l * V89_GRID + (y + 1) * 7 + 1;
#pragma unroll
            for (int x = 0; x < 5; x++) { float v = acc[j][x] + b; if (RES) v += op[x]; op[x] = fmaxf(v, 0.f); }   // RES: `out` holds the block input
        }
    }
}

constexpr size_t v89_smem_bytes() {
    return sizeof(float) * (size_t)(2 * V89_C * V89_TB * V89_GRID + 2 * V89_TB * V89_GRID + V89_TB * (50 + 64 + V89_AP) + 152 + V80_NBUF * V80_CH);
}

// boards: int8[.][75] HWC with `bstride` bytes between boards; masks: MW = 6 words per slot; list/count as in k_v80_forward.
__global__ void __launch_bounds__(V89_THREADS, 1)
k_v89_forward(const float* __restrict__ P, const __grid_constant__ V89Layout L, const __grid_constant__ V89Chunks CK,
              const int* count_ptr, const int* list, const int8_t* boards, int bstride, const uint32_t* masks,
              float* pi_out, float* v_out, int n_max) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TB = V89_TB, LDQ = V89_TB * V89_GRID, C = V89_C, A = V89_A, MW = 6;
    float* X = smem; float* H = X + C * LDQ; float* IN = H + C * LDQ; float* PF = IN + 2 * LDQ;      // PF [TB][50]
    float* VF = PF + TB * 50; float* VH = VF + 152 /* TB*25 padded to 16 B */; float* LG = VH + TB * 64; float* WB = LG + TB * V89_AP;
    static_assert((2 * C * LDQ + 2 * LDQ + TB * 50 + 152 + TB * 64) % 4 == 0 && (TB * V89_AP) % 4 == 0, "16-byte alignment of LG / WB");
    __shared__ int slot_of[TB];
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int tile0 = blockIdx.x * TB;
    if (tile0 >= count) return;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    WPipe89 wp; wp.P = P; wp.buf = WB; wp.ck = &CK; wp.issued = 0; wp.used = 0;
    wp.issue(); wp.issue();
    if (t < TB) { const int j = tile0 + t; slot_of[t] = j < count ? (list ? list[j] : j) : -1; }
    for (int k = t; k < 2 * C * LDQ + 2 * LDQ; k += V89_THREADS) smem[k] = 0.f;        // X, H, IN incl. their zero borders
    __syncthreads();
    for (int k = t; k < TB * 50; k += V89_THREADS) {                                   // input planes: workers, levels
        const int l = k / 50, r = k - l * 50, c = r / 25, pos = r - c * 25, slot = slot_of[l];
        if (slot >= 0) IN[c * LDQ + l * V89_GRID + (pos / 5 + 1) * 7 + pos % 5 + 1] = (float)boards[(size_t)slot * bstride + pos * 3 + c];
    }
    conv3x3<2, false>(wp, P + L.cbias[0], IN, X, LDQ);                                 // first_layer
    for (int b = 0; b < 5; b++) {                                                      // trunk: SimpleResBlock x 5
        conv3x3<C, false>(wp, P + L.cbias[1 + 2 * b], X, H, LDQ);
        conv3x3<C, true>(wp, P + L.cbias[2 + 2 * b], H, X, LDQ);
    }
    __syncthreads();
    // ---- heads: 1x1 convolutions (+BN, ReLU) to 2 + 1 planes, flattened channel-major as torch.flatten(x, 1) does
    for (int k = t; k < TB * 75; k += V89_THREADS) {
        const int l = k / 75, r = k - l * 75, ch = r / 25, pos = r - ch * 25;
        const float* xp = X + l * V89_GRID + (pos / 5 + 1) * 7 + pos % 5 + 1;
        float a = ch < 2 ? __ldg(P + L.pi_b + ch) : __ldg(P + L.v_b);
        for (int c = 0; c < C; c++) a = fmaf(ch < 2 ? __ldg(P + L.pi_w + c * 2 + ch) : __ldg(P + L.v_w + c), xp[c * LDQ], a);
        a = fmaxf(a, 0.f);
        if (ch < 2) PF[l * 50 + r] = a; else VF[l * 25 + pos] = a;
    }
    __syncthreads();
    for (int k = t; k < TB * (V89_AP / 4); k += V89_THREADS) {                         // policy Linear(50 -> 162)
        const int l = k / (V89_AP / 4), og = k - l * (V89_AP / 4);
        float4 a = __ldg(reinterpret_cast<const float4*>(P + L.pi_fcb + 4 * og));
        for (int i = 0; i < 50; i++) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(P + L.pi_fc + i * V89_AP + 4 * og)); const float x = PF[l * 50 + i];
            a.x = fmaf(w.x, x, a.x); a.y = fmaf(w.y, x, a.y); a.z = fmaf(w.z, x, a.z); a.w = fmaf(w.w, x, a.w);
        }
        *reinterpret_cast<float4*>(LG + l * V89_AP + 4 * og) = a;
    }
    for (int k = t; k < TB * 64; k += V89_THREADS) {                                   // value Linear(25 -> 64) + ReLU
        const int l = k >> 6, j = k & 63;
        float a = __ldg(P + L.v_fc1b + j);
        for (int i = 0; i < 25; i++) a = fmaf(__ldg(P + L.v_fc1 + i * 64 + j), VF[l * 25 + i], a);
        VH[l * 64 + j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    // masked softmax: where(valid, logits, -1e8) -> log_softmax -> exp (SantoriniNNet.py:279; GenericNNetWrapper.py:119)
    for (int l = warp; l < TB; l += V89_THREADS / 32) {
        const int slot = slot_of[l];
        if (slot < 0) continue;
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
    if (t < TB * 2) {                                                                  // value Linear(64 -> 2), tanh
        const int l = t >> 1, o = t & 1, slot = slot_of[l];
        float a = __ldg(P + L.v_fc2b + o);
        for (int j = 0; j < 64; j++) a = fmaf(__ldg(P + L.v_fc2 + o * 64 + j), VH[l * 64 + j], a);
        if (slot >= 0) v_out[(size_t)slot * 2 + o] = tanhf(a);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
}

}  // namespace azg

// net_v84.cuh -- AzulNNet version 84, eval-mode forward (azul/AzulNNet.py:84-111 layers, :127-137 forward) behind
// GenericNNetWrapper.predict / predict_server (GenericNNetWrapper.py:94-157).
//
// Same building blocks as SplendorNNet V80 on 23 tokens x 6 features: first_layer Linear(23->23)+BN over the token axis, trunk
// InvertedResidual1d(23->115->23, ReLU, SE avg, residual), policy block (23->115->46, Hardswish, SE avg, no residual) ->
// Linear(276->180)+ReLU -> Linear(180->180) -> masked log_softmax -> exp, value block (23->46->23, Hardswish, SE avg, residual) ->
// Linear(138->2)+ReLU -> Linear(2->2) -> tanh.  0.33 MFLOP per leaf: small enough that one fp32 CUDA-core kernel serves.
//
// One CTA = 8 leaves, 256 threads. Activations live in shared memory as [feature index][leaf] (the 8 leaves of a feature are two
// 128-bit words), so one thread owns one output feature for all 8 leaves: per input it needs ONE weight (K-major images, coalesced
// across the threads' outputs, L1/L2 resident: 118 k parameters) and two 128-bit shared loads for 8 FMAs. Layers with few outputs
// (SE fc1, the value head) split K over thread groups and add the partial sums in a fixed order (bit-reproducible results).
// BatchNorm (eval mode) is folded into the preceding linear on the host (v84_prepare).
#pragma once
#include "common.cuh"
#include "net_v80.cuh"      // act_apply

namespace azg {

constexpr int V84_NV = 23, V84_F = 6, V84_A = 180, V84_TB = 8, V84_THREADS = 256, V84_MW = 6;
struct V84Blk { int in, E, out, Q, act, res; int we, be, dw, sd, td, w1, b1, w2, b2, wp, bp; };
struct V84Layout { int w0, b0; V84Blk blk[3]; int pi2, bpi2, pi4, bpi4, v2, bv2, v4, bv4; int total; };

inline V84Layout v84_layout() {
    V84Layout L; int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
    static const int E_[3] = {115, 115, 46}, OUT_[3] = {23, 46, 23}, Q_[3] = {32, 32, 16}, ACT_[3] = {1, 2, 2};
    L.w0 = take(V84_NV * V84_NV); L.b0 = take(V84_NV);
    for (int k = 0; k < 3; k++) {
        V84Blk& B = L.blk[k]; B.in = V84_NV; B.E = E_[k]; B.out = OUT_[k]; B.Q = Q_[k]; B.act = ACT_[k]; B.res = B.in == B.out;
        B.we = take(B.in * B.E); B.be = take(B.E); B.dw = take(36); B.sd = take(B.E); B.td = take(B.E);
        B.w1 = take(B.E * B.Q); B.b1 = take(B.Q); B.w2 = take(B.Q * B.E); B.b2 = take(B.E); B.wp = take(B.E * B.out); B.bp = take(B.out);
    }
    L.pi2 = take(46 * V84_F * V84_A); L.bpi2 = take(V84_A); L.pi4 = take(V84_A * V84_A); L.bpi4 = take(V84_A);
    L.v2 = take(V84_NV * V84_F * 2); L.bv2 = take(4); L.v4 = take(4); L.bv4 = take(4);
    L.total = o; return L;
}
inline size_t v84_src_floats() {
    size_t n = 23 * 23 + 4 * 23;
    static const int E_[3] = {115, 115, 46}, OUT_[3] = {23, 46, 23}, Q_[3] = {32, 32, 16};
    for (int k = 0; k < 3; k++) n += (size_t)E_[k] * 23 + 4 * E_[k] + 36 + 4 * E_[k] + (size_t)Q_[k] * E_[k] + Q_[k] + (size_t)E_[k] * Q_[k] + E_[k] + (size_t)OUT_[k] * E_[k] + 4 * OUT_[k];
    n += (size_t)V84_A * 276 + V84_A + (size_t)V84_A * V84_A + V84_A + 2 * 138 + 2 + 4 + 2;
    return n;
}
// Host: state_dict order (nnet.py V84_TENSOR_ORDER = the V80 module names) -> K-major, BN-folded device blob.
inline void v84_prepare(const float* src, const V84Layout& L, float* dst) {
    for (int i = 0; i < L.total; i++) dst[i] = 0.f;
    const float* p = src;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    auto fold = [&](const float* W, int out, int in, int wdst, int bdst) {          // Linear(no bias) + BN(eval) -> Wt[i][o] * s_o, b_o
        const float* g = take(out); const float* b = take(out); const float* m = take(out); const float* v = take(out);
        for (int o = 0; o < out; o++) {
            const float s = g[o] / sqrtf(v[o] + 1e-5f);
            for (int i = 0; i < in; i++) dst[wdst + i * out + o] = W[o * in + i] * s;
            dst[bdst + o] = b[o] - m[o] * s;
        }
    };
    { const float* W = take(23 * 23); fold(W, 23, 23, L.w0, L.b0); }
    for (int k = 0; k < 3; k++) {
        const V84Blk& B = L.blk[k];
        { const float* W = take((size_t)B.E * B.in); fold(W, B.E, B.in, B.we, B.be); }
        { const float* W = take(36); for (int i = 0; i < 36; i++) dst[B.dw + i] = W[i];
          const float* g = take(B.E); const float* b = take(B.E); const float* m = take(B.E); const float* v = take(B.E);
          for (int c = 0; c < B.E; c++) { const float s = g[c] / sqrtf(v[c] + 1e-5f); dst[B.sd + c] = s; dst[B.td + c] = b[c] - m[c] * s; } }
        { const float* W = take((size_t)B.Q * B.E); for (int q = 0; q < B.Q; q++) for (int c = 0; c < B.E; c++) dst[B.w1 + c * B.Q + q] = W[q * B.E + c];
          const float* b = take(B.Q); for (int q = 0; q < B.Q; q++) dst[B.b1 + q] = b[q]; }
        { const float* W = take((size_t)B.E * B.Q); for (int c = 0; c < B.E; c++) for (int q = 0; q < B.Q; q++) dst[B.w2 + q * B.E + c] = W[c * B.Q + q];
          const float* b = take(B.E); for (int c = 0; c < B.E; c++) dst[B.b2 + c] = b[c]; }
        { const float* W = take((size_t)B.out * B.E); fold(W, B.out, B.E, B.wp, B.bp); }
    }
    { const float* W = take((size_t)V84_A * 276); for (int o = 0; o < V84_A; o++) for (int i = 0; i < 276; i++) dst[L.pi2 + i * V84_A + o] = W[o * 276 + i];
      const float* b = take(V84_A); for (int o = 0; o < V84_A; o++) dst[L.bpi2 + o] = b[o]; }
    { const float* W = take((size_t)V84_A * V84_A); for (int o = 0; o < V84_A; o++) for (int i = 0; i < V84_A; i++) dst[L.pi4 + i * V84_A + o] = W[o * V84_A + i];
      const float* b = take(V84_A); for (int o = 0; o < V84_A; o++) dst[L.bpi4 + o] = b[o]; }
    { const float* W = take(2 * 138); for (int o = 0; o < 2; o++) for (int i = 0; i < 138; i++) dst[L.v2 + i * 2 + o] = W[o * 138 + i];
      const float* b = take(2); dst[L.bv2] = b[0]; dst[L.bv2 + 1] = b[1]; }
    { const float* W = take(4); for (int i = 0; i < 4; i++) dst[L.v4 + i] = W[i]; const float* b = take(2); dst[L.bv4] = b[0]; dst[L.bv4 + 1] = b[1]; }
}

namespace v84 {
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ld8(const float* p) { F8 r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r; }
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void fma8(float (&acc)[8], float w, const F8& x) {
    acc[0] = fmaf(w, x.a.x, acc[0]); acc[1] = fmaf(w, x.a.y, acc[1]); acc[2] = fmaf(w, x.a.z, acc[2]); acc[3] = fmaf(w, x.a.w, acc[3]);
    acc[4] = fmaf(w, x.b.x, acc[4]); acc[5] = fmaf(w, x.b.y, acc[5]); acc[6] = fmaf(w, x.b.z, acc[6]); acc[7] = fmaf(w, x.b.w, acc[7]);
}
// token-axis Linear (+ folded BN, + activation, + optional residual): Y[(o*F+f)][l] = act(b[o] + sum_i Wt[i*out+o] X[(i*F+f)][l]) (+ R)
__device__ __forceinline__ void token_linear(const float* __restrict__ Wt, const float* __restrict__ b, int out, int in, const float* X, float* Y,
                                             int act, const float* R, int t) {
    for (int idx = t; idx < out * V84_F; idx += V84_THREADS) {
        const int f = idx / out, o = idx - f * out;              // consecutive threads = consecutive outputs: coalesced weight reads
        float acc[8]; const float bb = __ldg(b + o);
#pragma unroll
        for (int l = 0; l < 8; l++) acc[l] = bb;
        for (int i = 0; i < in; i++) fma8(acc, __ldg(Wt + i * out + o), ld8(X + (i * V84_F + f) * V84_TB));
        if (R) { const F8 r = ld8(R + (o * V84_F + f) * V84_TB); const float rr[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
            for (int l = 0; l < 8; l++) acc[l] += rr[l]; }
#pragma unroll
        for (int l = 0; l < 8; l++) acc[l] = act_apply(acc[l], act);
        st8(Y + (o * V84_F + f) * V84_TB, acc);
    }
}
// dense layer over flat features with a K split: PART[(ks*out+o)][l] = sum_{k in slice ks} Wt[k*out+o] X[k][l]
__device__ __forceinline__ void dense_partial(const float* __restrict__ Wt, int out, int K, int KS, const float* X, float* PART, int t) {
    const int per = (K + KS - 1) / KS;
    for (int idx = t; idx < out * KS; idx += V84_THREADS) {
        const int ks = idx / out, o = idx - ks * out;
        float acc[8];
#pragma unroll
        for (int l = 0; l < 8; l++) acc[l] = 0.f;
        const int k1 = min(K, (ks + 1) * per);
        for (int k = ks * per; k < k1; k++) fma8(acc, __ldg(Wt + k * out + o), ld8(X + k * V84_TB));
        st8(PART + (ks * out + o) * V84_TB, acc);
    }
}
// Y[o][l] = epi(b[o] + sum_ks PART[ks][o][l]) in a fixed order; epi: 0 none, 1 relu, 3 hardsigmoid
__device__ __forceinline__ void dense_reduce(const float* PART, const float* __restrict__ b, int out, int KS, float* Y, int epi, int t) {
    for (int idx = t; idx < out * V84_TB; idx += V84_THREADS) {
        const int o = idx >> 3, l = idx & 7;
        float a = __ldg(b + o);
        for (int ks = 0; ks < KS; ks++) a += PART[(ks * out + o) * V84_TB + l];
        Y[idx] = epi == 1 ? fmaxf(a, 0.f) : (epi == 3 ? fminf(fmaxf(a + 3.f, 0.f), 6.f) * (1.f / 6.f) : a);
    }
}
}  // namespace v84

constexpr int V84_SM_X = 0;                                       // [138][8] block input (trunk output is kept here for both heads)
constexpr int V84_SM_T = V84_SM_X + 138 * 8;                      // [138][8] first_layer output / trunk output
constexpr int V84_SM_E = V84_SM_T + 138 * 8;                      // [690][8] expanded activations
constexpr int V84_SM_D = V84_SM_E + 690 * 8;                      // [690][8] depthwise output (gated in place)
constexpr int V84_SM_H = V84_SM_D + 690 * 8;                      // [276][8] head block output
constexpr int V84_SM_SQ = V84_SM_H + 276 * 8;                     // [115][8] squeeze, then gates
constexpr int V84_SM_HID = V84_SM_SQ + 115 * 8;                   // [32][8]
constexpr int V84_SM_PART = V84_SM_HID + 32 * 8;                  // K-split partial sums: up to [8][32][8] / [2][180][8] / [64][2][8]
constexpr int V84_SM_H1 = V84_SM_PART + 2 * 180 * 8;              // [180][8] policy hidden, then logits
constexpr int V84_SM_FLOATS = V84_SM_H1 + 180 * 8;
constexpr size_t v84_smem_bytes() { return (size_t)V84_SM_FLOATS * 4; }

__global__ void __launch_bounds__(V84_THREADS)
k_v84_forward(const float* __restrict__ P, const __grid_constant__ V84Layout L, const int* count_ptr, const int* list, const int8_t* boards, int bstride,
              const uint32_t* masks, float* pi_out, float* v_out, int n_max) {
    using namespace v84;
    extern __shared__ __align__(16) float smf[];
    __shared__ int slot_of[V84_TB];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int tile0 = blockIdx.x * V84_TB;
    if (tile0 >= count) return;
    float* X = smf + V84_SM_X; float* T = smf + V84_SM_T; float* E = smf + V84_SM_E; float* D = smf + V84_SM_D; float* H = smf + V84_SM_H;
    float* SQ = smf + V84_SM_SQ; float* HID = smf + V84_SM_HID; float* PART = smf + V84_SM_PART; float* H1 = smf + V84_SM_H1;
    if (t < V84_TB) { const int j = tile0 + t; slot_of[t] = j < count ? (list ? list[j] : j) : -1; }
    __syncthreads();
    for (int idx = t; idx < 138 * V84_TB; idx += V84_THREADS) {   // X[i][l] = (float)board[l][i]
        const int l = idx / 138, i = idx - l * 138, slot = slot_of[l];
        X[i * V84_TB + l] = slot >= 0 ? (float)boards[(size_t)slot * bstride + i] : 0.f;
    }
    __syncthreads();
    token_linear(P + L.w0, P + L.b0, V84_NV, V84_NV, X, T, 0, nullptr, t);
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const V84Blk& B = L.blk[k];
        const float* IN = k == 0 ? T : X;                         // heads read the trunk output (parked in X after block 0)
        float* OUT = k == 0 ? X : H;
        token_linear(P + B.we, P + B.be, B.E, B.in, IN, E, B.act, nullptr, t);
        __syncthreads();
        for (int idx = t; idx < B.E * V84_F; idx += V84_THREADS) {   // "depthwise": shared Linear(6->6) over the features, BN per channel, act
            const int c = idx / V84_F, g = idx - c * V84_F;
            float acc[8];
#pragma unroll
            for (int l = 0; l < 8; l++) acc[l] = 0.f;
#pragma unroll
            for (int f = 0; f < V84_F; f++) fma8(acc, __ldg(P + B.dw + g * V84_F + f), ld8(E + (c * V84_F + f) * V84_TB));
            const float sd = __ldg(P + B.sd + c), td = __ldg(P + B.td + c);
#pragma unroll
            for (int l = 0; l < 8; l++) acc[l] = act_apply(fmaf(acc[l], sd, td), B.act);
            st8(D + idx * V84_TB, acc);
        }
        __syncthreads();
        for (int idx = t; idx < B.E * V84_TB; idx += V84_THREADS) {  // squeeze: AdaptiveAvgPool1d(1) over the 6 features
            const int c = idx >> 3, l = idx & 7;
            float s = 0.f;
#pragma unroll
            for (int f = 0; f < V84_F; f++) s += D[(c * V84_F + f) * V84_TB + l];
            SQ[idx] = s / (float)V84_F;
        }
        __syncthreads();
        dense_partial(P + B.w1, B.Q, B.E, 8, SQ, PART, t);          // fc1: E -> Q, ReLU
        __syncthreads();
        dense_reduce(PART, P + B.b1, B.Q, 8, HID, 1, t);
        __syncthreads();
        dense_partial(P + B.w2, B.E, B.Q, 2, HID, PART, t);         // fc2: Q -> E, hardsigmoid
        __syncthreads();
        dense_reduce(PART, P + B.b2, B.E, 2, SQ, 3, t);
        __syncthreads();
        for (int idx = t; idx < B.E * V84_F * V84_TB; idx += V84_THREADS) D[idx] *= SQ[((idx >> 3) / V84_F) * V84_TB + (idx & 7)];   // gate
        __syncthreads();
        token_linear(P + B.wp, P + B.bp, B.out, B.E, D, OUT, 0, B.res ? IN : nullptr, t);
        __syncthreads();
        if (k == 1) {   // ---- policy head: Linear(276 -> 180) + ReLU, Linear(180 -> 180), masked log_softmax -> exp
            dense_partial(P + L.pi2, V84_A, 276, 2, H, PART, t);
            __syncthreads();
            dense_reduce(PART, P + L.bpi2, V84_A, 2, H1, 1, t);
            __syncthreads();
            dense_partial(P + L.pi4, V84_A, V84_A, 2, H1, PART, t);
            __syncthreads();
            dense_reduce(PART, P + L.bpi4, V84_A, 2, H1, 0, t);
            __syncthreads();
            {
                const int sl = warp, slot = slot_of[sl];          // 8 warps = 8 leaves
                if (slot >= 0) {
                    float lg[V84_MW]; float mx = -INFINITY;
#pragma unroll
                    for (int kk = 0; kk < V84_MW; kk++) {
                        const int a = lane + 32 * kk;
                        const bool valid = a < V84_A && (masks[(size_t)slot * V84_MW + kk] >> lane & 1);
                        lg[kk] = a < V84_A ? (valid ? H1[a * V84_TB + sl] : -1e8f) : -INFINITY;
                        mx = fmaxf(mx, lg[kk]);
                    }
                    mx = warp_max_f32(mx);
                    float sum = 0.f;
#pragma unroll
                    for (int kk = 0; kk < V84_MW; kk++) sum += expf(lg[kk] - mx);
                    sum = warp_sum_f32(sum);
                    const float lse = logf(sum);
#pragma unroll
                    for (int kk = 0; kk < V84_MW; kk++) { const int a = lane + 32 * kk; if (a < V84_A) pi_out[(size_t)slot * V84_A + a] = expf(lg[kk] - mx - lse); }
                }
            }
            __syncthreads();
        }
    }
    // ---- value head: Linear(138 -> 2) + ReLU, Linear(2 -> 2), tanh (input: the value block's output in H)
    dense_partial(P + L.v2, 2, 138, 46, H, PART, t);
    __syncthreads();
    dense_reduce(PART, P + L.bv2, 2, 46, HID, 1, t);
    __syncthreads();
    if (t < 2 * V84_TB) {
        const int o = t >> 3, l = t & 7, slot = slot_of[l];
        const float a = __ldg(P + L.bv4 + o) + __ldg(P + L.v4 + o * 2) * HID[l] + __ldg(P + L.v4 + o * 2 + 1) * HID[V84_TB + l];
        if (slot >= 0) v_out[(size_t)slot * 2 + o] = tanhf(a);
    }
}

}  // namespace azg

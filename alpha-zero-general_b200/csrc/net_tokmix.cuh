// net_tokmix.cuh -- eval-mode forward of the reference's "token mixer" nets (first_layer Linear over the token axis + three
// InvertedResidual1d blocks + two Linear heads) for any token count / feature width, behind GenericNNetWrapper.predict /
// predict_server (GenericNNetWrapper.py:94-157):
//   AzulNNet version 84      (azul/AzulNNet.py:84-111,127-137): 23 tokens x 6 features, blocks 23->115->23 (ReLU), 23->115->46 and
//                            23->46->23 (Hardswish), SE avg everywhere, 180 actions, 2 players -- 0.41 MFLOP per leaf
//   SplendorNNet version 80  (splendor/SplendorNNet.py:259-280) for 3 and 4 PLAYERS: 71 / 88 tokens x 7 features, blocks nv->3nv->nv,
//                            SE avg in the trunk and max in the heads, 81 actions (the 2-player net, nv = 56, has its own tcgen05
//                            kernel, net_v80_tc.cuh)
// fp32 on the CUDA cores: these are the "next" games of SURVEY 8f, not the headline configuration.
//
// One CTA = 8 leaves, 512 threads. Activations live in shared memory as [feature index][leaf] (the 8 leaves of a feature are two
// 128-bit words); one thread owns TWO output features for all 8 leaves: per input it needs two weights (K-major images, coalesced
// across the threads' outputs, L1/L2 resident) and two 128-bit shared loads for 16 FMAs. Layers with few outputs (SE fc1, the value
// head) split K over thread groups and add the partial sums in a fixed order (bit-reproducible results). BatchNorm (eval mode) is
// folded into the preceding linear on the host (tokmix_prepare).
#pragma once
#include "common.cuh"
#include "net_v80.cuh"      // act_apply

namespace azg {

constexpr int TM_TB = 8, TM_THREADS = 512;   // 16 warps per CTA: the phases are latency-bound (L2 weight loads, barriers), more warps hide more of it
struct TokMixBlk { int in, E, out, Q, act, res, se_max; int we, be, dw, sd, td, w1, b1, w2, b2, wp, bp; };
struct TokMixLayout { int nv, f, a, np; int w0, b0; TokMixBlk blk[3]; int pi2, bpi2, pi4, bpi4, v2, bv2, v4, bv4; int total; };

inline int tm_make_divisible(int v, int divisor) {               // torchvision's _make_divisible (SplendorNNet.py:15-23)
    int nv = std::max(divisor, (v + divisor / 2) / divisor * divisor);
    if (nv < 0.9 * v) nv += divisor;
    return nv;
}
// kind 84: AzulNNet V84; kind 80: SplendorNNet V80 for `nv` tokens (np players)
inline TokMixLayout tokmix_layout(int kind, int nv, int f, int a, int np) {
    TokMixLayout L; int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
    L.nv = nv; L.f = f; L.a = a; L.np = np;
    int E_[3], OUT_[3], ACT_[3] = {1, 2, 2}, MAX_[3];
    if (kind == 84) { E_[0] = 5 * nv; E_[1] = 5 * nv; E_[2] = 2 * nv; OUT_[0] = nv; OUT_[1] = 2 * nv; OUT_[2] = nv; MAX_[0] = MAX_[1] = MAX_[2] = 0; }
    else { for (int k = 0; k < 3; k++) { E_[k] = 3 * nv; OUT_[k] = nv; } MAX_[0] = 0; MAX_[1] = MAX_[2] = 1; }
    L.w0 = take(nv * nv); L.b0 = take(nv);
    for (int k = 0; k < 3; k++) {
        TokMixBlk& B = L.blk[k]; B.in = nv; B.E = E_[k]; B.out = OUT_[k]; B.Q = tm_make_divisible(B.E / 4, 8); B.act = ACT_[k]; B.res = B.in == B.out; B.se_max = MAX_[k];
        B.we = take(B.in * B.E); B.be = take(B.E); B.dw = take(f * f); B.sd = take(B.E); B.td = take(B.E);
        B.w1 = take(B.E * B.Q); B.b1 = take(B.Q); B.w2 = take(B.Q * B.E); B.b2 = take(B.E); B.wp = take(B.E * B.out); B.bp = take(B.out);
    }
    L.pi2 = take(L.blk[1].out * f * a); L.bpi2 = take(a); L.pi4 = take(a * a); L.bpi4 = take(a);
    L.v2 = take(nv * f * np); L.bv2 = take(np); L.v4 = take(np * np); L.bv4 = take(np);
    L.total = o; return L;
}
inline size_t tokmix_src_floats(const TokMixLayout& L) {
    size_t n = (size_t)L.nv * L.nv + 4 * L.nv;
    for (int k = 0; k < 3; k++) { const TokMixBlk& B = L.blk[k]; n += (size_t)B.E * B.in + 4 * B.E + L.f * L.f + 4 * B.E + (size_t)B.Q * B.E + B.Q + (size_t)B.E * B.Q + B.E + (size_t)B.out * B.E + 4 * B.out; }
    const size_t kp = (size_t)L.blk[1].out * L.f, kv = (size_t)L.nv * L.f;
    n += L.a * kp + L.a + (size_t)L.a * L.a + L.a + L.np * kv + L.np + (size_t)L.np * L.np + L.np;
    return n;
}
// Host: state_dict order (nnet.py V80_TENSOR_ORDER: the same module names for both nets) -> K-major, BN-folded device blob.
inline void tokmix_prepare(const float* src, const TokMixLayout& L, float* dst) {
    for (int i = 0; i < L.total; i++) dst[i] = 0.f;
    const float* p = src; const int F = L.f, A = L.a, NP = L.np;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    auto fold = [&](const float* W, int out, int in, int wdst, int bdst) {          // Linear(no bias) + BN(eval) -> Wt[i][o] * s_o, b_o
        const float* g = take(out); const float* b = take(out); const float* m = take(out); const float* v = take(out);
        for (int o = 0; o < out; o++) {
            const float s = g[o] / sqrtf(v[o] + 1e-5f);
            for (int i = 0; i < in; i++) dst[wdst + i * out + o] = W[o * in + i] * s;
            dst[bdst + o] = b[o] - m[o] * s;
        }
    };
    { const float* W = take((size_t)L.nv * L.nv); fold(W, L.nv, L.nv, L.w0, L.b0); }
    for (int k = 0; k < 3; k++) {
        const TokMixBlk& B = L.blk[k];
        { const float* W = take((size_t)B.E * B.in); fold(W, B.E, B.in, B.we, B.be); }
        if (k == 0) {   // first_layer (Linear + BN, no activation) folded into this expand: W'[i][c] = sum_j W0[i][j] We[j][c], b' = be + sum_j b0[j] We[j][c]
            std::vector<double> w2((size_t)L.nv * B.E), b2(B.E);
            for (int c = 0; c < B.E; c++) {
                double a = dst[B.be + c];
                for (int j = 0; j < L.nv; j++) a += (double)dst[L.b0 + j] * (double)dst[B.we + j * B.E + c];
                b2[c] = a;
                for (int i = 0; i < L.nv; i++) {
                    double m = 0;
                    for (int j = 0; j < L.nv; j++) m += (double)dst[L.w0 + i * L.nv + j] * (double)dst[B.we + j * B.E + c];
                    w2[(size_t)i * B.E + c] = m;
                }
            }
            for (int c = 0; c < B.E; c++) dst[B.be + c] = (float)b2[c];
            for (size_t i = 0; i < w2.size(); i++) dst[B.we + i] = (float)w2[i];
        }
        { const float* W = take(F * F); for (int i = 0; i < F * F; i++) dst[B.dw + i] = W[i];
          const float* g = take(B.E); const float* b = take(B.E); const float* m = take(B.E); const float* v = take(B.E);
          for (int c = 0; c < B.E; c++) { const float s = g[c] / sqrtf(v[c] + 1e-5f); dst[B.sd + c] = s; dst[B.td + c] = b[c] - m[c] * s; } }
        { const float* W = take((size_t)B.Q * B.E); for (int q = 0; q < B.Q; q++) for (int c = 0; c < B.E; c++) dst[B.w1 + c * B.Q + q] = W[q * B.E + c];
          const float* b = take(B.Q); for (int q = 0; q < B.Q; q++) dst[B.b1 + q] = b[q]; }
        { const float* W = take((size_t)B.E * B.Q); for (int c = 0; c < B.E; c++) for (int q = 0; q < B.Q; q++) dst[B.w2 + q * B.E + c] = W[c * B.Q + q];
          const float* b = take(B.E); for (int c = 0; c < B.E; c++) dst[B.b2 + c] = b[c]; }
        { const float* W = take((size_t)B.out * B.E); fold(W, B.out, B.E, B.wp, B.bp); }
    }
    const int KP = L.blk[1].out * F, KV = L.nv * F;
    { const float* W = take((size_t)A * KP); for (int o = 0; o < A; o++) for (int i = 0; i < KP; i++) dst[L.pi2 + i * A + o] = W[o * KP + i];
      const float* b = take(A); for (int o = 0; o < A; o++) dst[L.bpi2 + o] = b[o]; }
    { const float* W = take((size_t)A * A); for (int o = 0; o < A; o++) for (int i = 0; i < A; i++) dst[L.pi4 + i * A + o] = W[o * A + i];
      const float* b = take(A); for (int o = 0; o < A; o++) dst[L.bpi4 + o] = b[o]; }
    { const float* W = take((size_t)NP * KV); for (int o = 0; o < NP; o++) for (int i = 0; i < KV; i++) dst[L.v2 + i * NP + o] = W[o * KV + i];
      const float* b = take(NP); for (int o = 0; o < NP; o++) dst[L.bv2 + o] = b[o]; }
    { const float* W = take((size_t)NP * NP); for (int i = 0; i < NP * NP; i++) dst[L.v4 + i] = W[i]; const float* b = take(NP); for (int o = 0; o < NP; o++) dst[L.bv4 + o] = b[o]; }
}

namespace tm {
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
// A thread owns TWO outputs (o, o + 1) of one feature row for all 8 leaves: the two 128-bit activation loads of an input feed 16 FMAs.
// With one output per thread the phase is bound by shared-memory wavefronts (a broadcast 128-bit load still costs one wavefront per
// quarter-warp): 1 wavefront-cycle per FMA instruction. The summation order per output is unchanged.
template <int F>
__device__ __forceinline__ void token_linear(const float* __restrict__ Wt, const float* __restrict__ b, int out, int in, const float* X, float* Y,
                                             int act, const float* R, int t) {
    const int op = (out + 1) >> 1;
    for (int idx = t; idx < op * F; idx += TM_THREADS) {
        const int f = idx / op, o = 2 * (idx - f * op);          // consecutive threads = consecutive output pairs: coalesced weight reads
        const bool two = o + 1 < out;
        float a0[8], a1[8]; const float b0 = __ldg(b + o), b1 = two ? __ldg(b + o + 1) : 0.f;
#pragma unroll
        for (int l = 0; l < 8; l++) { a0[l] = b0; a1[l] = b1; }
        const float* w = Wt + o; const float* x = X + f * TM_TB;
#pragma unroll 4
        for (int i = 0; i < in; i++) {
            const float w0 = __ldg(w + i * out), w1 = __ldg(w + i * out + 1);   // an odd tail reads the next row's first weight (inside the blob); its sums are dropped
            const F8 xv = ld8(x + i * F * TM_TB);
            fma8(a0, w0, xv); fma8(a1, w1, xv);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (h == 1 && !two) break;
            float (&acc)[8] = h ? a1 : a0;
            const int oo = o + h;
            if (R) { const F8 r = ld8(R + (oo * F + f) * TM_TB); const float rr[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
                for (int l = 0; l < 8; l++) acc[l] += rr[l]; }
#pragma unroll
            for (int l = 0; l < 8; l++) acc[l] = act_apply(acc[l], act);
            st8(Y + (oo * F + f) * TM_TB, acc);
        }
    }
}
// dense layer over flat features with a K split: PART[(ks*out+o)][l] = sum_{k in slice ks} Wt[k*out+o] X[k][l]; two outputs per thread
__device__ __forceinline__ void dense_partial(const float* __restrict__ Wt, int out, int K, int KS, const float* X, float* PART, int t) {
    const int per = (K + KS - 1) / KS, op = (out + 1) >> 1;
    for (int idx = t; idx < op * KS; idx += TM_THREADS) {
        const int ks = idx / op, o = 2 * (idx - ks * op);
        const bool two = o + 1 < out;
        float a0[8], a1[8];
#pragma unroll
        for (int l = 0; l < 8; l++) a0[l] = a1[l] = 0.f;
        const int k1 = min(K, (ks + 1) * per);
        const float* w = Wt + o;
#pragma unroll 4
        for (int k = ks * per; k < k1; k++) {
            const float w0 = __ldg(w + k * out), w1 = __ldg(w + k * out + 1);
            const F8 xv = ld8(X + k * TM_TB);
            fma8(a0, w0, xv); fma8(a1, w1, xv);
        }
        st8(PART + (ks * out + o) * TM_TB, a0);
        if (two) st8(PART + (ks * out + o + 1) * TM_TB, a1);
    }
}
// Y[o][l] = epi(b[o] + sum_ks PART[ks][o][l]) in a fixed order; epi: 0 none, 1 relu, 3 hardsigmoid
__device__ __forceinline__ void dense_reduce(const float* PART, const float* __restrict__ b, int out, int KS, float* Y, int epi, int t) {
    for (int idx = t; idx < out * TM_TB; idx += TM_THREADS) {
        const int o = idx >> 3, l = idx & 7;
        float a = __ldg(b + o);
        for (int ks = 0; ks < KS; ks++) a += PART[(ks * out + o) * TM_TB + l];
        Y[idx] = epi == 1 ? fmaxf(a, 0.f) : (epi == 3 ? fminf(fmaxf(a + 3.f, 0.f), 6.f) * (1.f / 6.f) : a);
    }
}
}  // namespace tm

// Shared-memory map (floats) for a net with NV tokens, F features, A actions, expansion <= EMAX, block output <= OMAX, SE hidden <= QMAX.
template <int NV, int F, int A, int NP, int EMAX, int OMAX, int QMAX> struct TokMixSmem {
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int VKS = 32;                                // K split of the value head
    static constexpr int PKS = A <= 96 ? 6 : 4;                   // K split of the two policy Linears (A / 2 output pairs x PKS work items; 5 instead of 4 for Azul measured 45 % SLOWER)
    static constexpr int X = 0, T = X + NV * F * 8, E = T + NV * F * 8, D = E + EMAX * F * 8, H = D + EMAX * F * 8, SQ = H + OMAX * F * 8,
                         HID = SQ + EMAX * 8, PART = HID + cmax(QMAX, NP) * 8,
                         PART_N = cmax(cmax(8 * QMAX, 2 * EMAX), VKS * NP) * 8, TOTAL = PART + PART_N,
                         // the policy head runs between the policy block and the value block, when E and D are dead: its K-split partial sums and
                         // its hidden layer live there (a smaller CTA leaves more of the SM's 256 KB to the L1 that holds the weights)
                         PPART = E, H1 = E + PKS * A * 8;
    static_assert(PKS * A * 8 + A * 8 <= 2 * EMAX * F * 8, "policy head scratch must fit the expanded-activation buffers");
    static constexpr size_t bytes() { return (size_t)TOTAL * 4; }
};

template <int NV, int F, int A, int NP, int EMAX, int OMAX, int QMAX>
__global__ void __launch_bounds__(TM_THREADS)
k_tokmix_forward(const float* __restrict__ P, const __grid_constant__ TokMixLayout L, const int* count_ptr, const int* list, const int8_t* boards, int bstride,
                 const uint32_t* masks, float* pi_out, float* v_out, int n_max) {
    using namespace tm;
    typedef TokMixSmem<NV, F, A, NP, EMAX, OMAX, QMAX> SM;
    constexpr int MW = (A + 31) / 32, S = NV * F;
    extern __shared__ __align__(16) float smf[];
    __shared__ int slot_of[TM_TB];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int tile0 = blockIdx.x * TM_TB;
    if (tile0 >= count) return;
    float* X = smf + SM::X; float* T = smf + SM::T; float* E = smf + SM::E; float* D = smf + SM::D; float* H = smf + SM::H;
    float* SQ = smf + SM::SQ; float* HID = smf + SM::HID; float* PART = smf + SM::PART; float* H1 = smf + SM::H1; float* PPART = smf + SM::PPART;
    if (t < TM_TB) { const int j = tile0 + t; slot_of[t] = j < count ? (list ? list[j] : j) : -1; }
    __syncthreads();
    for (int idx = t; idx < S * TM_TB; idx += TM_THREADS) {       // X[i][l] = (float)board[l][i]
        const int l = idx / S, i = idx - l * S, slot = slot_of[l];
        X[i * TM_TB + l] = slot >= 0 ? (float)boards[(size_t)slot * bstride + i] : 0.f;
    }
    __syncthreads();
    token_linear<F>(P + L.w0, P + L.b0, NV, NV, X, T, 0, nullptr, t);   // T = first_layer(x): only the residual of the trunk block needs it.
    // No barrier: first_layer has no activation and is folded into the trunk block's expand weights on the host (tokmix_prepare), so
    // that expand reads the raw planes X as well and runs in the same phase.
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const TokMixBlk& B = L.blk[k];
        const float* IN = X;                                      // trunk block: the raw planes (folded first_layer); heads: the trunk output (kept in X after block 0)
        const float* RES = k == 0 ? T : X;
        float* OUT = k == 0 ? X : H;
        token_linear<F>(P + B.we, P + B.be, B.E, B.in, IN, E, B.act, nullptr, t);
        __syncthreads();
        constexpr int GP = (F + 1) / 2;
        for (int idx = t; idx < B.E * GP; idx += TM_THREADS) {       // "depthwise": shared Linear(F->F) over the features, BN per channel, act;
            const int c = idx / GP, g = 2 * (idx - c * GP);          // two output features per thread share the activation loads
            const bool two = g + 1 < F;
            float a0[8], a1[8];
#pragma unroll
            for (int l = 0; l < 8; l++) a0[l] = a1[l] = 0.f;
#pragma unroll
            for (int f = 0; f < F; f++) {
                const F8 xv = ld8(E + (c * F + f) * TM_TB);
                fma8(a0, __ldg(P + B.dw + g * F + f), xv); fma8(a1, two ? __ldg(P + B.dw + (g + 1) * F + f) : 0.f, xv);
            }
            const float sd = __ldg(P + B.sd + c), td = __ldg(P + B.td + c);
#pragma unroll
            for (int l = 0; l < 8; l++) { a0[l] = act_apply(fmaf(a0[l], sd, td), B.act); a1[l] = act_apply(fmaf(a1[l], sd, td), B.act); }
            st8(D + (c * F + g) * TM_TB, a0);
            if (two) st8(D + (c * F + g + 1) * TM_TB, a1);
        }
        __syncthreads();
        for (int idx = t; idx < B.E * TM_TB; idx += TM_THREADS) {     // squeeze over the F features: AdaptiveAvgPool1d(1) or AdaptiveMaxPool1d(1)
            const int c = idx >> 3, l = idx & 7;
            float s = B.se_max ? -INFINITY : 0.f;
#pragma unroll
            for (int f = 0; f < F; f++) { const float x = D[(c * F + f) * TM_TB + l]; s = B.se_max ? fmaxf(s, x) : s + x; }
            SQ[idx] = B.se_max ? s : s / (float)F;
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
        for (int idx = t; idx < B.E * F * TM_TB; idx += TM_THREADS) D[idx] *= SQ[((idx >> 3) / F) * TM_TB + (idx & 7)];   // gate
        __syncthreads();
        token_linear<F>(P + B.wp, P + B.bp, B.out, B.E, D, OUT, 0, B.res ? RES : nullptr, t);
        __syncthreads();
        if (k == 1) {   // ---- policy head: Linear(out*F -> A) + ReLU, Linear(A -> A), masked log_softmax -> exp
            dense_partial(P + L.pi2, A, B.out * F, SM::PKS, H, PPART, t);
            __syncthreads();
            dense_reduce(PPART, P + L.bpi2, A, SM::PKS, H1, 1, t);
            __syncthreads();
            dense_partial(P + L.pi4, A, A, SM::PKS, H1, PPART, t);
            __syncthreads();
            dense_reduce(PPART, P + L.bpi4, A, SM::PKS, H1, 0, t);
            __syncthreads();
            {
                const int sl = warp & 7, slot = warp < TM_TB ? slot_of[sl] : -1;   // warps 0-7 = the 8 leaves
                if (slot >= 0) {
                    float lg[MW]; float mx = -INFINITY;
#pragma unroll
                    for (int kk = 0; kk < MW; kk++) {
                        const int a = lane + 32 * kk;
                        const bool valid = a < A && (masks[(size_t)slot * MW + kk] >> lane & 1);
                        lg[kk] = a < A ? (valid ? H1[a * TM_TB + sl] : -1e8f) : -INFINITY;
                        mx = fmaxf(mx, lg[kk]);
                    }
                    mx = warp_max_f32(mx);
                    float sum = 0.f;
#pragma unroll
                    for (int kk = 0; kk < MW; kk++) sum += expf(lg[kk] - mx);
                    sum = warp_sum_f32(sum);
                    const float lse = logf(sum);
#pragma unroll
                    for (int kk = 0; kk < MW; kk++) { const int a = lane + 32 * kk; if (a < A) pi_out[(size_t)slot * A + a] = expf(lg[kk] - mx - lse); }
                }
            }
            __syncthreads();
        }
    }
    // ---- value head: Linear(NV*F -> NP) + ReLU, Linear(NP -> NP), tanh (input: the value block's output in H)
    dense_partial(P + L.v2, NP, S, SM::VKS, H, PART, t);
    __syncthreads();
    dense_reduce(PART, P + L.bv2, NP, SM::VKS, HID, 1, t);
    __syncthreads();
    if (t < NP * TM_TB) {
        const int o = t >> 3, l = t & 7, slot = slot_of[l];
        float a = __ldg(P + L.bv4 + o);
#pragma unroll
        for (int i = 0; i < NP; i++) a = fmaf(__ldg(P + L.v4 + o * NP + i), HID[i * TM_TB + l], a);
        if (slot >= 0) v_out[(size_t)slot * NP + o] = tanhf(a);
    }
}

// The three instances in the library
typedef TokMixSmem<23, 6, 180, 2, 115, 46, 32> TMS_V84;
typedef TokMixSmem<71, 7, 81, 3, 213, 71, 56> TMS_V80_3P;
typedef TokMixSmem<88, 7, 81, 4, 264, 88, 64> TMS_V80_4P;

}  // namespace azg

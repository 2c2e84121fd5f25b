// net_v80.cuh -- batched forward of SplendorNNet version 80 (splendor/SplendorNNet.py:149-204,259-280,
// 397-404,440) in eval mode, replacing onnxruntime's InferenceSession.run behind
// GenericNNetWrapper.predict / predict_server (GenericNNetWrapper.py:94-157).
//
// One CTA evaluates a tile of TB leaves end to end with every activation resident in shared memory
// (input 56x7 "token x feature" grid -> first layer -> trunk inverted-residual block -> policy head
// and value head, each another inverted-residual block + 2 linears -> masked softmax / tanh). fp32 on
// CUDA cores: the layers are 56/168/40-wide contractions with a 7-wide feature axis, the reference
// computes in fp32, and the parity bar is 1e-5 on pi and v. BatchNorm (eval) is folded into the weights
// on the host; weights are stored K-major so a thread's output-channel group is one 128-bit load.
//
// Also here: the deterministic hash-net used by parity tests (oracle/hashnet.py).
#pragma once
#include "common.cuh"

namespace azg {

constexpr int V80_THREADS = 256;
constexpr int V80_Q = 40;                 // SE squeeze width: _make_divisible(168 // 4, 8)

// Offsets (in floats) into the prepared device blob. NV = number of board rows (56 for 2 players).
struct V80Layout {
    int nv, e, np;
    int w0, b0;
    struct Blk { int we, be, wd, sd, td, fc1, b1, fc2, b2, wp, bp; } blk[3];
    int pi2, bpi2, pi4, bpi4, v2, bv2, v4, bv4;
    int total;
    static constexpr int PIP = 84;        // 81 policy outputs padded to a multiple of 4
};
inline V80Layout v80_layout(int nv, int np) {
    V80Layout L; L.nv = nv; L.e = 3 * nv; L.np = np; int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
    L.w0 = take(nv * nv); L.b0 = take(nv);
    for (int k = 0; k < 3; k++) {
        auto& B = L.blk[k];
        B.we = take(nv * L.e); B.be = take(L.e); B.wd = take(49); B.sd = take(L.e); B.td = take(L.e);
        B.fc1 = take(L.e * V80_Q); B.b1 = take(V80_Q); B.fc2 = take(V80_Q * L.e); B.b2 = take(L.e);
        B.wp = take(L.e * nv); B.bp = take(nv);
    }
    L.pi2 = take(nv * 7 * V80Layout::PIP); L.bpi2 = take(V80Layout::PIP);
    L.pi4 = take(81 * V80Layout::PIP); L.bpi4 = take(V80Layout::PIP);
    L.v2 = take(nv * 7 * 4); L.bv2 = take(4); L.v4 = take(16); L.bv4 = take(4);
    L.total = o; return L;
}

// Host: fold BN and transpose. `src` = state_dict tensors in V80_TENSOR_ORDER (see nnet.py), `dst` = prepared blob.
inline void v80_prepare(const float* src, int nv, int np, const V80Layout& L, float* dst) {
    const float* p = src; const int E = L.e, Q = V80_Q, A = 81, PIP = V80Layout::PIP;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    auto fold_token = [&](int out, int in, int w_off, int b_off) {      // Linear(no bias) over tokens + BN(out)
        const float *W = take((size_t)out * in), *g = take(out), *b = take(out), *m = take(out), *v = take(out);
        for (int o = 0; o < out; o++) {
            float s = g[o] / sqrtf(v[o] + 1e-5f);
            dst[b_off + o] = b[o] - m[o] * s;
            for (int i = 0; i < in; i++) dst[w_off + i * out + o] = W[o * in + i] * s;          // K-major
        }
    };
    for (int i = 0; i < L.total; i++) dst[i] = 0.f;
    fold_token(nv, nv, L.w0, L.b0);
    for (int k = 0; k < 3; k++) {
        const auto& B = L.blk[k];
        fold_token(E, nv, B.we, B.be);
        const float *Wd = take(49), *g = take(E), *b = take(E), *m = take(E), *v = take(E);
        for (int i = 0; i < 49; i++) dst[B.wd + i] = Wd[i];
        for (int c = 0; c < E; c++) { float s = g[c] / sqrtf(v[c] + 1e-5f); dst[B.sd + c] = s; dst[B.td + c] = b[c] - m[c] * s; }
        const float *f1 = take((size_t)Q * E), *b1 = take(Q), *f2 = take((size_t)E * Q), *b2 = take(E);
        for (int q = 0; q < Q; q++) { dst[B.b1 + q] = b1[q]; for (int c = 0; c < E; c++) dst[B.fc1 + c * Q + q] = f1[q * E + c]; }
        for (int c = 0; c < E; c++) { dst[B.b2 + c] = b2[c]; for (int q = 0; q < Q; q++) dst[B.fc2 + q * E + c] = f2[c * Q + q]; }
        fold_token(nv, E, B.wp, B.bp);
    }
    const int F = nv * 7;
    const float *w = take((size_t)A * F), *b = take(A);
    for (int o = 0; o < A; o++) { dst[L.bpi2 + o] = b[o]; for (int i = 0; i < F; i++) dst[L.pi2 + i * PIP + o] = w[o * F + i]; }
    w = take((size_t)A * A); b = take(A);
    for (int o = 0; o < A; o++) { dst[L.bpi4 + o] = b[o]; for (int i = 0; i < A; i++) dst[L.pi4 + i * PIP + o] = w[o * A + i]; }
    w = take((size_t)np * F); b = take(np);
    for (int o = 0; o < np; o++) { dst[L.bv2 + o] = b[o]; for (int i = 0; i < F; i++) dst[L.v2 + i * 4 + o] = w[o * F + i]; }
    w = take((size_t)np * np); b = take(np);
    for (int o = 0; o < np; o++) { dst[L.bv4 + o] = b[o]; for (int i = 0; i < np; i++) dst[L.v4 + o * 4 + i] = w[o * np + i]; }
}
inline size_t v80_src_floats(int nv, int np) {
    size_t E = 3 * (size_t)nv, Q = V80_Q, n = 0;
    n += (size_t)nv * nv + 4 * nv;
    n += 3 * (E * nv + 4 * E + 49 + 4 * E + Q * E + Q + E * Q + E + nv * E + 4 * nv);
    n += 81 * (size_t)nv * 7 + 81 + 81 * 81 + 81 + (size_t)np * nv * 7 + np + (size_t)np * np + np;
    return n;
}

__device__ __forceinline__ float act_apply(float x, int act) {
    if (act == 1) return fmaxf(x, 0.f);
    if (act == 2) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);           // hardswish
    return x;
}

// Y[o][s*8+f] = act(sum_k Wt[k][o] * X[k][s*8+f] + bias[o]) (+ R). One task = OCT output channels x one leaf (7 features).
template <int M, int K, int OCT, int TB, int ACT, bool RES>
__device__ __forceinline__ void token_gemm(const float* __restrict__ Wt, const float* __restrict__ bias,
                                           const float* X, float* Y, const float* R) {
    static_assert(M % OCT == 0 && OCT % 4 == 0, "tile");
    constexpr int NG = M / OCT, LD = TB * 8;
    for (int t = threadIdx.x; t < NG * TB; t += V80_THREADS) {
        const int og = t / TB, s = t - og * TB, o0 = og * OCT;
        float acc[OCT][7];
#pragma unroll
        for (int j = 0; j < OCT; j++)
#pragma unroll
            for (int f = 0; f < 7; f++) acc[j][f] = 0.f;
        const float* xp = X + s * 8;
        const float* wp = Wt + o0;
#pragma unroll 2
        for (int k = 0; k < K; k++) {
            const float4 xa = *reinterpret_cast<const float4*>(xp + k * LD);
            const float4 xb = *reinterpret_cast<const float4*>(xp + k * LD + 4);
            const float x[7] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z};
            float wv[OCT];
#pragma unroll
            for (int j = 0; j < OCT; j += 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * M + j));
                wv[j] = q.x; wv[j + 1] = q.y; wv[j + 2] = q.z; wv[j + 3] = q.w;
            }
#pragma unroll
            for (int j = 0; j < OCT; j++)
#pragma unroll
                for (int f = 0; f < 7; f++) acc[j][f] = fmaf(wv[j], x[f], acc[j][f]);
        }
#pragma unroll
        for (int j = 0; j < OCT; j++) {
            const float b = __ldg(bias + o0 + j);
            float* yp = Y + (o0 + j) * LD + s * 8;
#pragma unroll
            for (int f = 0; f < 7; f++) {
                float y = act_apply(acc[j][f] + b, ACT);
                if (RES) y += R[(o0 + j) * LD + s * 8 + f];
                yp[f] = y;
            }
        }
    }
}

// Inverted-residual block (InvertedResidual1d, SplendorNNet.py:189-204): X[NV] -> Y[NV] (+X). E is 3NV-row scratch.
template <int NV, int TB, int ACT, bool SE_MAX>
__device__ __forceinline__ void ir_block(const float* __restrict__ P, const V80Layout::Blk B, const float* X, float* Y,
                                         float* E, float* SQ, float* HID) {
    constexpr int EC = 3 * NV, LD = TB * 8, Q = V80_Q;
    token_gemm<EC, NV, 12, TB, ACT, false>(P + B.we, P + B.be, X, E, nullptr);
    __syncthreads();
    // "depthwise": the same Linear(7->7) on the feature axis of every channel, BN per channel, activation; + squeeze
    for (int t = threadIdx.x; t < EC * TB; t += V80_THREADS) {
        const int c = t / TB, s = t - c * TB;
        float* ep = E + c * LD + s * 8;
        const float4 xa = *reinterpret_cast<const float4*>(ep), xb = *reinterpret_cast<const float4*>(ep + 4);
        const float x[7] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z};
        const float sd = __ldg(P + B.sd + c), td = __ldg(P + B.td + c);
        float pool = SE_MAX ? -INFINITY : 0.f;
#pragma unroll
        for (int g = 0; g < 7; g++) {
            float a = 0.f;
#pragma unroll
            for (int f = 0; f < 7; f++) a = fmaf(__ldg(P + B.wd + g * 7 + f), x[f], a);
            a = act_apply(fmaf(a, sd, td), ACT);
            ep[g] = a;
            pool = SE_MAX ? fmaxf(pool, a) : pool + a;
        }
        SQ[c * TB + s] = SE_MAX ? pool : pool * (1.f / 7.f);
    }
    __syncthreads();
    // squeeze-excitation: fc1 (EC->Q) relu, fc2 (Q->EC) hardsigmoid (SqueezeExcitation1d, SplendorNNet.py:172-187)
    for (int t = threadIdx.x; t < Q * TB; t += V80_THREADS) {
        const int q = t / TB, s = t - q * TB;
        float a = __ldg(P + B.b1 + q);
        for (int c = 0; c < EC; c++) a = fmaf(__ldg(P + B.fc1 + c * Q + q), SQ[c * TB + s], a);
        HID[q * TB + s] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < EC * TB; t += V80_THREADS) {
        const int c = t / TB, s = t - c * TB;
        float a = __ldg(P + B.b2 + c);
#pragma unroll 8
        for (int q = 0; q < Q; q++) a = fmaf(__ldg(P + B.fc2 + q * EC + c), HID[q * TB + s], a);
        const float gate = fminf(fmaxf(a + 3.f, 0.f), 6.f) * (1.f / 6.f);
        float* ep = E + c * LD + s * 8;
#pragma unroll
        for (int f = 0; f < 7; f++) ep[f] *= gate;
    }
    __syncthreads();
    token_gemm<NV, EC, 8, TB, 0, true>(P + B.wp, P + B.bp, E, Y, X);
    __syncthreads();
}

template <int NV, int TB>
constexpr size_t v80_smem_bytes() {
    return sizeof(float) * (size_t)(2 * NV * TB * 8 + 3 * NV * TB * 8 + 3 * NV * TB + V80_Q * TB + 2 * V80Layout::PIP * TB + 8 * TB);
}

// boards: int8, `bstride` bytes between boards; masks: MW words per slot. If `list` is given, tile entry j evaluates
// slot list[j] (the engine's compacted leaf list) and *count_ptr entries exist; otherwise slot j, j < n_max.
template <int NV, int NP, int TB>
__global__ void __launch_bounds__(V80_THREADS, 1)
k_v80_forward(const float* __restrict__ P, V80Layout L, const int* count_ptr, const int* list, const int8_t* boards, int bstride,
              const uint32_t* masks, float* pi_out, float* v_out, int n_max) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = TB * 8, A = 81, PIP = V80Layout::PIP, F = NV * 7;
    float* X0 = smem; float* T = X0 + NV * LD; float* E = T + NV * LD; float* SQ = E + 3 * NV * LD;
    float* HID = SQ + 3 * NV * TB; float* H1 = HID + V80_Q * TB; float* LG = H1 + PIP * TB; float* VH = LG + PIP * TB;
    __shared__ int slot_of[TB];
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int tile0 = blockIdx.x * TB;
    if (tile0 >= count) return;
    if (threadIdx.x < TB) { int j = tile0 + threadIdx.x; slot_of[threadIdx.x] = j < count ? (list ? list[j] : j) : -1; }
    __syncthreads();
    // input: T[i][s*8+f] = (float)board[s][i*7+f]
    for (int t = threadIdx.x; t < TB * NV * 8; t += V80_THREADS) {
        const int i = t / LD, r = t - i * LD, s = r >> 3, f = r & 7;
        const int slot = slot_of[s];
        T[t] = (slot >= 0 && f < 7) ? (float)boards[(size_t)slot * bstride + i * 7 + f] : 0.f;
    }
    __syncthreads();
    token_gemm<NV, NV, 4, TB, 0, false>(P + L.w0, P + L.b0, T, X0, nullptr);                   // first_layer
    __syncthreads();
    ir_block<NV, TB, 1, false>(P, L.blk[0], X0, T, E, SQ, HID);                                 // trunk (ReLU, SE avg)
    // ---- policy head
    ir_block<NV, TB, 2, true>(P, L.blk[1], T, X0, E, SQ, HID);                                  // Hardswish, SE max
    for (int t = threadIdx.x; t < (PIP / 4) * TB; t += V80_THREADS) {                           // Linear(F->81)+ReLU, flatten index = row*7+f
        const int og = t / TB, s = t - og * TB;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < NV; i++) {
            const float* xp = X0 + i * LD + s * 8;
#pragma unroll
            for (int f = 0; f < 7; f++) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(P + L.pi2 + (size_t)(i * 7 + f) * PIP + og * 4));
                const float x = xp[f];
                acc[0] = fmaf(q.x, x, acc[0]); acc[1] = fmaf(q.y, x, acc[1]); acc[2] = fmaf(q.z, x, acc[2]); acc[3] = fmaf(q.w, x, acc[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) H1[(og * 4 + j) * TB + s] = fmaxf(acc[j] + __ldg(P + L.bpi2 + og * 4 + j), 0.f);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (PIP / 4) * TB; t += V80_THREADS) {                           // Linear(81->81)
        const int og = t / TB, s = t - og * TB;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < A; i++) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(P + L.pi4 + (size_t)i * PIP + og * 4));
            const float x = H1[i * TB + s];
            acc[0] = fmaf(q.x, x, acc[0]); acc[1] = fmaf(q.y, x, acc[1]); acc[2] = fmaf(q.z, x, acc[2]); acc[3] = fmaf(q.w, x, acc[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) LG[(og * 4 + j) * TB + s] = acc[j] + __ldg(P + L.bpi4 + og * 4 + j);
    }
    __syncthreads();
    // masked softmax: where(valid, logits, -1e8) -> log_softmax -> exp (SplendorNNet.py:404,440; GenericNNetWrapper.py:119)
    {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int s = w; s < TB; s += V80_THREADS / 32) {
            const int slot = slot_of[s];
            if (slot < 0) continue;
            float l[3]; float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int a = lane + 32 * k;
                const bool valid = a < A && (masks[(size_t)slot * 3 + k] >> lane & 1);
                l[k] = a < A ? (valid ? LG[a * TB + s] : -1e8f) : -INFINITY;
                mx = fmaxf(mx, l[k]);
            }
            mx = warp_max_f32(mx);
            float e[3], sum = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) { e[k] = expf(l[k] - mx); sum += e[k]; }
            sum = warp_sum_f32(sum);
            const float lse = logf(sum);
#pragma unroll
            for (int k = 0; k < 3; k++) { const int a = lane + 32 * k; if (a < A) pi_out[(size_t)slot * A + a] = expf(l[k] - mx - lse); }
        }
    }
    // ---- value head (reads T; X0/E are free again after the policy linears above consumed X0)
    __syncthreads();
    ir_block<NV, TB, 2, true>(P, L.blk[2], T, X0, E, SQ, HID);
    for (int t = threadIdx.x; t < NP * TB; t += V80_THREADS) {                                  // Linear(F->np)+ReLU
        const int o = t / TB, s = t - o * TB;
        float a = __ldg(P + L.bv2 + o);
        for (int i = 0; i < NV; i++)
#pragma unroll
            for (int f = 0; f < 7; f++) a = fmaf(__ldg(P + L.v2 + (i * 7 + f) * 4 + o), X0[i * LD + s * 8 + f], a);
        VH[o * TB + s] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NP * TB; t += V80_THREADS) {                                  // Linear(np->np), tanh
        const int o = t / TB, s = t - o * TB, slot = slot_of[s];
        float a = __ldg(P + L.bv4 + o);
#pragma unroll
        for (int i = 0; i < NP; i++) a = fmaf(__ldg(P + L.v4 + o * 4 + i), VH[i * TB + s], a);
        if (slot >= 0) v_out[(size_t)slot * NP + o] = tanhf(a);
    }
    (void)F;
}

// ---- hash-net (tests only; oracle/hashnet.py): one warp per leaf ------------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
template <int S, int A, int NP, int MW>
__global__ void k_hashnet_forward(const int* count_ptr, const int* list, const int8_t* boards, int bstride, const uint32_t* masks,
                                  float* pi_out, float* v_out, int n_max) {
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= count) return;
    const int slot = list ? list[j] : j;
    uint32_t h = 0x811C9DC5u;
    if (lane == 0) { const int8_t* b = boards + (size_t)slot * bstride; for (int i = 0; i < S; i++) h = (h ^ (uint8_t)b[i]) * 16777619u; }
    h = __shfl_sync(FULL, h, 0);
    int wv[MW]; int W = 0, bw = -1, ba = 0x7FFFFFFF;
#pragma unroll
    for (int k = 0; k < MW; k++) {
        const int a = lane + 32 * k;
        const bool valid = a < A && (masks[(size_t)slot * MW + k] >> lane & 1);
        wv[k] = valid ? 256 + (int)(fmix32(h + (uint32_t)a * 0x9E3779B1u) & 1023u) : 0;
        W += wv[k];
        if (wv[k] > bw) { bw = wv[k]; ba = a; }              // ascending a per lane: first index kept on ties
    }
    W = warp_sum_i32(W);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int ow = __shfl_xor_sync(FULL, bw, o), oa = __shfl_xor_sync(FULL, ba, o);
        if (ow > bw || (ow == bw && oa < ba)) { bw = ow; ba = oa; }
    }
    int kv[MW], ks = 0;
#pragma unroll
    for (int k = 0; k < MW; k++) { kv[k] = (int)(((long long)wv[k] * 4096) / W); ks += kv[k]; }
    ks = warp_sum_i32(ks);
#pragma unroll
    for (int k = 0; k < MW; k++) {
        const int a = lane + 32 * k;
        if (a < A) { int kk = kv[k] + (a == ba ? 4096 - ks : 0); pi_out[(size_t)slot * A + a] = (float)kk / 4096.0f; }
    }
    if (lane == 0) {
        const int jv = (int)(fmix32(h ^ 0xABCDEF01u) % 129u) - 64;
        const float v0 = (float)jv / 64.0f;
        v_out[(size_t)slot * NP] = v0;
        for (int p = 1; p < NP; p++) v_out[(size_t)slot * NP + p] = -v0 / (float)(NP - 1);
    }
}

}  // namespace azg

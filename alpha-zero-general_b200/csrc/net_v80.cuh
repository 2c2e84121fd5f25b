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

constexpr int V80_THREADS = 384;           // 12 warps: one per 14 expanded channels
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
    if (act == 2) return x * __saturatef(fmaf(x, 1.f / 6.f, 0.5f));                    // hardswish = x relu6(x + 3) / 6 as one saturating FMA and one multiply
    return x;
}

// ---- weight streaming: the prepared blob is consumed as a static list of <= V80_CH-float chunks, copied
// global -> shared with cp.async into a ring of V80_NBUF buffers, two chunks ahead of the math -------------------
constexpr int V80_TB = 16;                // leaves per CTA
constexpr int V80_CH = 4704;              // floats per chunk buffer (= 28 rows of a [k][168] matrix = 84 rows of a [k][56] one)
constexpr int V80_NBUF = 3;
constexpr int V80_MAXCHUNK = 40;
struct V80Chunks { int off[V80_MAXCHUNK]; int n[V80_MAXCHUNK]; int count; };
struct V80DW { float w[3][56]; };         // the three blocks' shared Linear(7->7) "depthwise" weights, read from the constant bank

inline V80Chunks v80_chunks(const V80Layout& L) {
    V80Chunks c; c.count = 0;
    auto add = [&](int off, int n) { c.off[c.count] = off; c.n[c.count] = n; c.count++; };
    const int nv = L.nv, e = L.e, PIP = V80Layout::PIP;
    add(L.w0, nv * nv);
    auto block = [&](const V80Layout::Blk& B) {
        add(B.we, 28 * e); add(B.we + 28 * e, (nv - 28) * e);                 // expand  [k=nv][e], 28-row chunks
        add(B.fc1, 84 * V80_Q); add(B.fc1 + 84 * V80_Q, (e - 84) * V80_Q);    // fc1     [c=e][q]
        add(B.fc2, 20 * e); add(B.fc2 + 20 * e, (V80_Q - 20) * e);            // fc2     [q][e]
        add(B.wp, 84 * nv); add(B.wp + 84 * nv, (e - 84) * nv);               // project [k=e][nv]
    };
    block(L.blk[0]); block(L.blk[1]);
    for (int j = 0; j < nv / 8; j++) add(L.pi2 + j * 56 * PIP, 56 * PIP);     // pi2 [k=nv*7][84], 8 tokens per chunk
    add(L.pi4, 41 * PIP); add(L.pi4 + 41 * PIP, 40 * PIP);                    // pi4 [81][84]
    block(L.blk[2]);
    add(L.v2, nv * 7 * 4);                                                    // v2  [k=nv*7][4]
    return c;
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float4* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
struct WPipe {
    const float* P; float* buf; const V80Chunks* ck; int issued, used;
    __device__ __forceinline__ void issue() {                // chunk `issued` -> buffer issued % NBUF; always commits a group
        if (issued < ck->count) {
            const float4* src = reinterpret_cast<const float4*>(P + ck->off[issued]);
            float* dst = buf + (issued % V80_NBUF) * V80_CH; const int n4 = ck->n[issued] >> 2;
            for (int i = threadIdx.x; i < n4; i += V80_THREADS) cp_async16(dst + 4 * i, src + i);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        issued++;
    }
    // Next chunk in list order. The barrier inside also orders the activations written by the previous stage.
    __device__ __forceinline__ const float* acquire() {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(V80_NBUF - 2));
        __syncthreads();
        issue();                                             // refills the buffer whose chunk was consumed before the barrier
        return buf + (used++ % V80_NBUF) * V80_CH;
    }
};

// acc[j][0..3] += sum over the chunk rows kk with (kbase + kk) % KS == ks of W[kk][c0 + j] * X[kbase + kk][4*lane .. 4*lane+3].
// Lane = one quad of activation columns (half a leaf); the NC output channels are uniform across the warp (broadcast loads).
template <int NC, int M, int KS, int LD>
__device__ __forceinline__ void gemm_chunk(float (&acc)[NC][4], const float* W, const float* X, int kbase, int rows, int c0, int ks, int lane) {
    int kk = KS == 1 ? 0 : ((ks - kbase) % KS + KS) % KS;
#pragma unroll 4
    for (; kk < rows; kk += KS) {
        const float4 x = *reinterpret_cast<const float4*>(X + (kbase + kk) * LD + 4 * lane);
        const float* wr = W + kk * M + c0;
        float w[NC];
#pragma unroll
        for (int j = 0; j < NC; j += 2) { const float2 t = *reinterpret_cast<const float2*>(wr + j); w[j] = t.x; w[j + 1] = t.y; }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            acc[j][0] = fmaf(w[j], x.x, acc[j][0]); acc[j][1] = fmaf(w[j], x.y, acc[j][1]);
            acc[j][2] = fmaf(w[j], x.z, acc[j][2]); acc[j][3] = fmaf(w[j], x.w, acc[j][3]);
        }
    }
}

// Y[NV][LD] = W^T X + bias (+ R): 12 warps = 4 channel groups of 14 x 3 interleaved K slices; the slices are summed
// through `scratch` (2*NV*LD floats, may alias X only if X is dead -- it is not: pass a free region).
template <int NV, int K, int LD, bool RES>
__device__ __forceinline__ void token_gemm_nv(WPipe& wp, const float* __restrict__ bias, const float* X, float* Y, const float* R,
                                              float* scratch, int n_chunks, int rows_per_chunk) {
    static_assert(NV == 56, "channel grouping assumes 4 x 14 output channels");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cg = warp & 3, ks = warp >> 2, c0 = 14 * cg;
    float acc[14][4];
#pragma unroll
    for (int j = 0; j < 14; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    for (int ch = 0; ch < n_chunks; ch++) {
        const float* W = wp.acquire();
        const int kbase = ch * rows_per_chunk, rows = min(rows_per_chunk, K - kbase);
        gemm_chunk<14, NV, 3, LD>(acc, W, X, kbase, rows, c0, ks, lane);
    }
    __syncthreads();                                         // every warp is done reading X (scratch may now be written if it aliases)
    if (ks > 0) {
#pragma unroll
        for (int j = 0; j < 14; j++)
            *reinterpret_cast<float4*>(scratch + ((ks - 1) * NV + c0 + j) * LD + 4 * lane) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
    __syncthreads();
    if (ks == 0) {
#pragma unroll
        for (int j = 0; j < 14; j++) {
            const int c = c0 + j; const float b = __ldg(bias + c);
            const float4 p1 = *reinterpret_cast<const float4*>(scratch + c * LD + 4 * lane);
            const float4 p2 = *reinterpret_cast<const float4*>(scratch + (NV + c) * LD + 4 * lane);
            float4 y = make_float4(acc[j][0] + p1.x + p2.x + b, acc[j][1] + p1.y + p2.y + b, acc[j][2] + p1.z + p2.z + b, acc[j][3] + p1.w + p2.w + b);
            if (RES) { const float4 r = *reinterpret_cast<const float4*>(R + c * LD + 4 * lane); y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w; }
            *reinterpret_cast<float4*>(Y + c * LD + 4 * lane) = y;
        }
    }
}

// Inverted-residual block (InvertedResidual1d, SplendorNNet.py:189-204): Y = X + project(SE(depthwise(expand(X)))).
// E: 3NV x LD scratch (expanded activations, then the K-slice partial sums of the projection).
template <int NV, int TB, int ACT, bool SE_MAX>
__device__ __forceinline__ void ir_block(WPipe& wp, const float* __restrict__ P, const V80Layout::Blk B, const float* dw,
                                         const float* X, float* Y, float* E, float* SQ, float* HID) {
    constexpr int EC = 3 * NV, LD = TB * 8, Q = V80_Q;
    static_assert(EC == 14 * (V80_THREADS / 32), "one warp per 14 expanded channels");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    {   // ---- expand (Linear over tokens + BN, act) fused with the "depthwise" Linear(7->7) + BN + act and the SE squeeze
        float acc[14][4];
#pragma unroll
        for (int j = 0; j < 14; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        const int c0 = 14 * warp;
        for (int ch = 0; ch < 2; ch++) { const float* W = wp.acquire(); gemm_chunk<14, EC, 1, LD>(acc, W, X, 28 * ch, 28, c0, 0, lane); }
        const int h = lane & 1, s = lane >> 1;               // lane = (leaf s, feature half h): features 4h .. 4h+3
#pragma unroll
        for (int j = 0; j < 14; j++) {
            const int c = c0 + j;
            const float be = __ldg(P + B.be + c), sd = __ldg(P + B.sd + c), td = __ldg(P + B.td + c);
            float v[4], o[4], e[8];
#pragma unroll
            for (int i = 0; i < 4; i++) { v[i] = act_apply(acc[j][i] + be, ACT); o[i] = __shfl_xor_sync(FULL, v[i], 1); }
#pragma unroll
            for (int i = 0; i < 4; i++) { e[i] = h ? o[i] : v[i]; e[4 + i] = h ? v[i] : o[i]; }
            float d[4]; float pool = SE_MAX ? -INFINITY : 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int g = 4 * h + i;                     // output feature of this lane
                float a = 0.f;
#pragma unroll
                for (int f = 0; f < 7; f++) a = fmaf(h ? dw[(4 + i) * 7 + f] : dw[i * 7 + f], e[f], a);
                a = act_apply(fmaf(a, sd, td), ACT);
                d[i] = g < 7 ? a : 0.f;
                if (g < 7) pool = SE_MAX ? fmaxf(pool, a) : pool + a;
            }
            *reinterpret_cast<float4*>(E + c * LD + 4 * lane) = make_float4(d[0], d[1], d[2], d[3]);
            const float po = __shfl_xor_sync(FULL, pool, 1);
            if (h == 0) SQ[c * TB + s] = SE_MAX ? fmaxf(pool, po) : (pool + po) * (1.f / 7.f);
        }
    }
    {   // ---- squeeze-excitation fc1 (EC -> Q, ReLU): 160 threads = 10 q-quads x 16 leaves (SqueezeExcitation1d, :172-187)
        const int qg = t >> 4, s = t & 15;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < 2; ch++) {
            const float* W = wp.acquire();                   // barrier: SQ complete
            if (t < 160) {
#pragma unroll 4
                for (int kk = 0; kk < 84; kk++) {
                    const float4 w4 = *reinterpret_cast<const float4*>(W + kk * Q + 4 * qg);
                    const float x = SQ[(84 * ch + kk) * TB + s];
                    a[0] = fmaf(w4.x, x, a[0]); a[1] = fmaf(w4.y, x, a[1]); a[2] = fmaf(w4.z, x, a[2]); a[3] = fmaf(w4.w, x, a[3]);
                }
            }
        }
        if (t < 160) {
#pragma unroll
            for (int j = 0; j < 4; j++) HID[(4 * qg + j) * TB + s] = fmaxf(a[j] + __ldg(P + B.b1 + 4 * qg + j), 0.f);
        }
    }
    {   // ---- fc2 (Q -> EC, hardsigmoid) and the channel scaling of E: 42 channel quads x 16 leaves = 672 tasks
        constexpr int NT = (EC / 4) * TB;
        float g[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int ch = 0; ch < 2; ch++) {
            const float* W = wp.acquire();                   // barrier: HID complete
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int task = t + r * V80_THREADS;
                if (task < NT) {
                    const int cq = task >> 4, s = task & 15;
#pragma unroll 4
                    for (int kk = 0; kk < 20; kk++) {
                        const float4 w4 = *reinterpret_cast<const float4*>(W + kk * EC + 4 * cq);
                        const float x = HID[(20 * ch + kk) * TB + s];
                        g[r][0] = fmaf(w4.x, x, g[r][0]); g[r][1] = fmaf(w4.y, x, g[r][1]); g[r][2] = fmaf(w4.z, x, g[r][2]); g[r][3] = fmaf(w4.w, x, g[r][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int task = t + r * V80_THREADS;
            if (task < NT) {
                const int cq = task >> 4, s = task & 15;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = 4 * cq + j;
                    const float gate = fminf(fmaxf(g[r][j] + __ldg(P + B.b2 + c) + 3.f, 0.f), 6.f) * (1.f / 6.f);
                    float4* ep = reinterpret_cast<float4*>(E + c * LD + s * 8);
                    float4 a = ep[0], b = ep[1];
                    a.x *= gate; a.y *= gate; a.z *= gate; a.w *= gate; b.x *= gate; b.y *= gate; b.z *= gate; b.w *= gate;
                    ep[0] = a; ep[1] = b;
                }
            }
        }
    }
    // ---- project (Linear over tokens + BN) + residual; the partial sums of the K slices go through E once it is dead
    token_gemm_nv<NV, EC, LD, true>(wp, P + B.bp, E, Y, X, E, 2, 84);
}

template <int NV, int TB>
constexpr size_t v80_smem_bytes() {
    return sizeof(float) * (size_t)(2 * NV * TB * 8 + 3 * NV * TB * 8 + 3 * NV * TB + V80_Q * TB + 8 * TB + V80_NBUF * V80_CH);
}

// boards: int8, `bstride` bytes between boards; masks: MW words per slot. If `list` is given, tile entry j evaluates
// slot list[j] (the engine's compacted leaf list) and *count_ptr entries exist; otherwise slot j, j < n_max.
template <int NV, int NP, int TB>
__global__ void __launch_bounds__(V80_THREADS, 1)
k_v80_forward(const float* __restrict__ P, const __grid_constant__ V80Layout L, const __grid_constant__ V80Chunks CK, const __grid_constant__ V80DW DW,
              const int* count_ptr, const int* list, const int8_t* boards, int bstride,
              const uint32_t* masks, float* pi_out, float* v_out, int n_max) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = TB * 8, A = 81, PIP = V80Layout::PIP;
    static_assert(TB == 16, "lane <-> (leaf, feature half) mapping assumes 16 leaves per tile");
    float* X0 = smem; float* T = X0 + NV * LD; float* E = T + NV * LD; float* SQ = E + 3 * NV * LD;
    float* HID = SQ + 3 * NV * TB; float* VH = HID + V80_Q * TB; float* WB = VH + 8 * TB;
    float* H1 = E; float* LG = E + PIP * TB;                 // policy-head linears live in E while it is dead
    __shared__ int slot_of[TB];
    const int count = count_ptr ? min(*count_ptr, n_max) : n_max;
    const int tile0 = blockIdx.x * TB;
    if (tile0 >= count) return;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    WPipe wp; wp.P = P; wp.buf = WB; wp.ck = &CK; wp.issued = 0; wp.used = 0;
    wp.issue(); wp.issue();
    if (t < TB) { int j = tile0 + t; slot_of[t] = j < count ? (list ? list[j] : j) : -1; }
    __syncthreads();
    // input: T[i][s*8+f] = (float)board[s][i*7+f]
    for (int k = t; k < TB * NV * 8; k += V80_THREADS) {
        const int i = k / LD, r = k - i * LD, s = r >> 3, f = r & 7;
        const int slot = slot_of[s];
        T[k] = (slot >= 0 && f < 7) ? (float)boards[(size_t)slot * bstride + i * 7 + f] : 0.f;
    }
    token_gemm_nv<NV, NV, LD, false>(wp, P + L.b0, T, X0, nullptr, E, 1, NV);                          // first_layer
    ir_block<NV, TB, 1, false>(wp, P, L.blk[0], DW.w[0], X0, T, E, SQ, HID);                            // trunk (ReLU, SE avg)
    // ---- policy head
    ir_block<NV, TB, 2, true>(wp, P, L.blk[1], DW.w[1], T, X0, E, SQ, HID);                             // Hardswish, SE max
    {
        // Linear(392 -> 81) + ReLU over the flattened (token, feature) grid, then Linear(81 -> 81): 21 output quads x 16 leaves
        const bool on = t < 352;
        const int og = min(2 * (t >> 5) + (t & 1), PIP / 4 - 1), s = (t >> 1) & 15;
        const bool live = on && (2 * (t >> 5) + (t & 1)) < PIP / 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < NV / 8; ch++) {
            const float* W = wp.acquire();                   // first barrier: X0 complete
            if (live) {
#pragma unroll 2
                for (int il = 0; il < 8; il++) {
                    const float* xp = X0 + (8 * ch + il) * LD + s * 8;
                    const float4 xa = *reinterpret_cast<const float4*>(xp), xb = *reinterpret_cast<const float4*>(xp + 4);
                    const float x[7] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z};
#pragma unroll
                    for (int f = 0; f < 7; f++) {
                        const float4 q = *reinterpret_cast<const float4*>(W + (il * 7 + f) * PIP + 4 * og);
                        acc[0] = fmaf(q.x, x[f], acc[0]); acc[1] = fmaf(q.y, x[f], acc[1]); acc[2] = fmaf(q.z, x[f], acc[2]); acc[3] = fmaf(q.w, x[f], acc[3]);
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < 4; j++) H1[(4 * og + j) * TB + s] = fmaxf(acc[j] + __ldg(P + L.bpi2 + 4 * og + j), 0.f);
        }
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        for (int ch = 0; ch < 2; ch++) {
            const float* W = wp.acquire();                   // first barrier: H1 complete
            const int rows = ch == 0 ? 41 : 40;
            if (live) {
#pragma unroll 4
                for (int kk = 0; kk < rows; kk++) {
                    const float4 q = *reinterpret_cast<const float4*>(W + kk * PIP + 4 * og);
                    const float x = H1[(41 * ch + kk) * TB + s];
                    acc[0] = fmaf(q.x, x, acc[0]); acc[1] = fmaf(q.y, x, acc[1]); acc[2] = fmaf(q.z, x, acc[2]); acc[3] = fmaf(q.w, x, acc[3]);
                }
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < 4; j++) LG[(4 * og + j) * TB + s] = acc[j] + __ldg(P + L.bpi4 + 4 * og + j);
        }
    }
    __syncthreads();
    // masked softmax: where(valid, logits, -1e8) -> log_softmax -> exp (SplendorNNet.py:404,440; GenericNNetWrapper.py:119)
    for (int s = warp; s < TB; s += V80_THREADS / 32) {
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
    // ---- value head (reads T; X0 is free again: the policy linears consumed it)
    ir_block<NV, TB, 2, true>(wp, P, L.blk[2], DW.w[2], T, X0, E, SQ, HID);
    {
        if (t < 8 * TB) VH[t] = 0.f;
        const float* W = wp.acquire();                       // barrier: X0 complete, VH zeroed
        const int s = t & 15, kp = t >> 4;                   // 24 token slices x 16 leaves
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int i = kp; i < NV; i += V80_THREADS / 16) {
            const float* xp = X0 + i * LD + s * 8;
#pragma unroll
            for (int f = 0; f < 7; f++) {
                const float4 q = *reinterpret_cast<const float4*>(W + (i * 7 + f) * 4);
                const float x = xp[f];
                a0 = fmaf(q.x, x, a0); a1 = fmaf(q.y, x, a1); a2 = fmaf(q.z, x, a2); a3 = fmaf(q.w, x, a3);
            }
        }
        atomicAdd(VH + 0 * TB + s, a0); atomicAdd(VH + 1 * TB + s, a1);
        if (NP > 2) { atomicAdd(VH + 2 * TB + s, a2); atomicAdd(VH + 3 * TB + s, a3); }
        __syncthreads();
        if (t < NP * TB) {                                   // ReLU, Linear(np -> np), tanh
            const int o = t / TB, sl = t - o * TB, slot = slot_of[sl];
            float a = __ldg(P + L.bv4 + o);
#pragma unroll
            for (int i = 0; i < NP; i++) a = fmaf(__ldg(P + L.v4 + o * 4 + i), fmaxf(VH[i * TB + sl] + __ldg(P + L.bv2 + i), 0.f), a);
            if (slot >= 0) v_out[(size_t)slot * NP + o] = tanhf(a);
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
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

// api.cu -- the C ABI declared in include/azg.h: batched game-step kernels, the net handle and the
// search / self-play engine. Host side is plain C++ (no torch types); device memory is owned by the
// handles, caller buffers may be host or device memory.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/azg.h"
#include "common.cuh"
#include "net_v80.cuh"
#include "net_v80_tc.cuh"
#include "net_v21.cuh"
#include "net_v89.cuh"
#include "net_v89_tc.cuh"
#include "net_tokmix.cuh"
#include "azul.cuh"
#include "abalone.cuh"
#include "santorini.cuh"
#include "selfplay.cuh"
#include "splendor.cuh"
#include "tree.cuh"

using namespace azg;

// ------------------------------------------------------------------ errors -----------------------------
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)
#define CKL() CK(cudaGetLastError())

extern "C" int azg_abi_version(void) { return AZG_ABI_VERSION; }
extern "C" const char* azg_last_error(void) { return g_err.c_str(); }
extern "C" int azg_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

extern "C" int azg_set_device(int device) { CK(cudaSetDevice(device)); return 0; }

static int require_device() {
    if (azg_device_count() <= 0) return fail("no CUDA device: the B200 engine has no CPU fallback");
    return 0;
}

// ------------------------------------------------------------------ host/device staging ----------------
static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
struct Scratch {                       // grow-only device buffer
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, n) != cudaSuccess) return fail("cudaMalloc scratch failed");
        cap = n; return 0;
    }
    ~Scratch() { if (p) cudaFree(p); }
};
// An argument that may live on the host: in() gives a device pointer (copying if needed), out() copies back.
struct Arg {
    Scratch s; const void* src = nullptr; void* dst = nullptr; void* dev = nullptr; size_t bytes = 0; bool staged = false;
    int in(const void* p, size_t n, cudaStream_t st) {
        src = p; dst = nullptr; bytes = n; staged = false; dev = const_cast<void*>(p);
        if (!p || n == 0) { dev = nullptr; return 0; }
        if (is_device_ptr(p)) return 0;
        if (s.ensure(n)) return 1;
        CK(cudaMemcpyAsync(s.p, p, n, cudaMemcpyHostToDevice, st));
        dev = s.p; staged = true; return 0;
    }
    int outbuf(void* p, size_t n) {
        dst = p; src = nullptr; bytes = n; staged = false; dev = p;
        if (!p || n == 0) { dev = nullptr; return 0; }
        if (is_device_ptr(p)) return 0;
        if (s.ensure(n)) return 1;
        dev = s.p; staged = true; return 0;
    }
    int flush(cudaStream_t st) {
        if (staged && dst) CK(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, st));
        return 0;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(dev); }
};
static thread_local Arg tl_arg[12];
static bool any_staged(int n) { for (int i = 0; i < n; i++) if (tl_arg[i].staged) return true; return false; }

// ------------------------------------------------------------------ game registry ------------------------
// GameSwitcher.py:3-35 of the reference maps a game name to its module; here a game id selects the template
// instance. Every host entry point below is a template over the game type G and is reached through DISPATCH.
typedef Splendor<2> SP2;
#define DISPATCH(game_id, np, CALL)                                                                       \
    do {                                                                                                  \
        if ((game_id) == AZG_GAME_SPLENDOR && (np) == 2) { typedef Splendor<2> G; return CALL; }           \
        if ((game_id) == AZG_GAME_SPLENDOR && (np) == 3) { typedef Splendor<3> G; return CALL; }           \
        if ((game_id) == AZG_GAME_SPLENDOR && (np) == 4) { typedef Splendor<4> G; return CALL; }           \
        if ((game_id) == AZG_GAME_SANTORINI && (np) == 2) { typedef Santorini G; return CALL; }            \
        if ((game_id) == AZG_GAME_ABALONE && (np) == 2) { typedef Abalone G; return CALL; }                \
        if ((game_id) == AZG_GAME_AZUL && (np) == 2) { typedef Azul G; return CALL; }                      \
        return fail("unknown game (built: 1 = splendor with 2, 3 or 4 players, 2 = santorini without gods, 3 = abalone, 4 = azul with 2 players)"); \
    } while (0)

template <class G> static int game_info_t(azg_game_info_t* out) {
    out->game_id = G::GAME_ID; out->num_players = G::NP; out->state_rows = G::D0; out->state_cols = G::D1; out->state_depth = G::D2;
    out->state_bytes = G::S; out->action_size = G::A; out->max_symmetries = G::MAX_SYM; out->max_game_len = G::MAX_MOVES;
    return 0;
}
extern "C" int azg_game_info(int game_id, int num_players, azg_game_info_t* out) {
    if (!out) return fail("out is NULL");
    DISPATCH(game_id, num_players, game_info_t<G>(out));
}

// ------------------------------------------------------------------ batched game-step kernels ----------
// One warp per board; the board is staged in shared memory exactly as in the search kernels.
template <class G> __host__ __device__ constexpr int gk_warps() { return G::A > 1024 ? 1 : 4; }     // warps (= boards) per CTA of the game-step kernels
template <class G> __device__ __forceinline__ void load_board(int8_t* sb, const int8_t* src, int lane) {
    for (int i = lane; i < G::SP; i += 32) sb[i] = i < G::S ? src[i] : (int8_t)0;
    __syncwarp();
}
template <class G> __device__ __forceinline__ void store_board(int8_t* dst, const int8_t* sb, int lane) {
    __syncwarp();
    for (int i = lane; i < G::S; i += 32) dst[i] = sb[i];
}

template <class G>
__global__ void k_game_init(int n, const uint64_t* seeds, int8_t* boards) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    if (lane == 0) { Philox rng(seeds[i], 0x1717, 0); G::init_game(sm[w], &rng); }
    store_board<G>(boards + (size_t)i * G::S, sm[w], lane);
}
template <class G>
__global__ void k_game_valid(int n, const int8_t* boards, const int* players, uint8_t* mask) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    load_board<G>(sm[w], boards + (size_t)i * G::S, lane);
    __shared__ uint32_t smw[gk_warps<G>()][G::MASK_WORDS];
    G::valid_mask(sm[w], players ? players[i] : 0, lane, smw[w]);
    for (int a = lane; a < G::A; a += 32) mask[(size_t)i * G::A + a] = (smw[w][a >> 5] >> (a & 31)) & 1;
}
template <class G>
__global__ void k_game_next(int n, const int8_t* boards, const int* players, const int* actions, const long long* seeds,
                            const uint64_t* keys, int8_t* out, int* out_np) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    load_board<G>(sm[w], boards + (size_t)i * G::S, lane);
    if (lane == 0) {
        Philox rng(keys ? keys[i] : (uint64_t)i, 0x2323, 0);
        int np = G::make_move(sm[w], actions[i], players ? players[i] : 0, seeds ? seeds[i] : 0, &rng);
        if (out_np) out_np[i] = np;
    }
    store_board<G>(out + (size_t)i * G::S, sm[w], lane);
}
template <class G>
__global__ void k_game_ended(int n, const int8_t* boards, const int* next_players, float* out) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    load_board<G>(sm[w], boards + (size_t)i * G::S, lane);
    float es[G::NP]; G::ended(sm[w], next_players ? next_players[i] : 0, es, lane);
    if (lane == 0) for (int p = 0; p < G::NP; p++) out[(size_t)i * G::NP + p] = es[p];
}
template <class G>
__global__ void k_game_canonical(int n, const int8_t* boards, const int* players, int8_t* out) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    load_board<G>(sm[w], boards + (size_t)i * G::S, lane);
    const int p = players ? players[i] : 0;
    if (p != 0) G::swap_players(sm[w], p, lane);
    store_board<G>(out + (size_t)i * G::S, sm[w], lane);
}
template <class G>
__global__ void k_game_round_score(int n, const int8_t* boards, int* rounds, int* scores) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    load_board<G>(sm[w], boards + (size_t)i * G::S, lane);
    if (lane == 0 && rounds) rounds[i] = G::round(sm[w]);
    if (scores && lane < G::NP) scores[(size_t)i * G::NP + lane] = G::score(sm[w], lane);
}
template <class G>
__global__ void k_game_symmetries(int n, const int8_t* boards, const float* pi, const uint8_t* mask, int8_t* ob, float* opi,
                                  uint8_t* om, int* ok) {
    __shared__ __align__(16) int8_t sm[gk_warps<G>()][G::SP];
    __shared__ float spi[gk_warps<G>()][G::A];
    __shared__ uint8_t smask[gk_warps<G>()][G::A];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * gk_warps<G>() + w;
    if (i >= n) return;
    load_board<G>(sm[w], boards + (size_t)i * G::S, lane);
    for (int a = lane; a < G::A; a += 32) { spi[w][a] = pi[(size_t)i * G::A + a]; smask[w][a] = mask[(size_t)i * G::A + a]; }
    __syncwarp();
    const int K = G::num_symmetries(sm[w]);
    if (lane == 0) ok[i] = K;
    for (int k = 0; k < G::MAX_SYM; k++) {
        const size_t o = (size_t)i * G::MAX_SYM + k;
        if (k < K) G::symmetry(sm[w], spi[w], smask[w], k, lane, ob + o * G::S, opi + o * G::A, om + o * G::A);
        else {
            for (int j = lane; j < G::S; j += 32) ob[o * G::S + j] = 0;
            for (int a = lane; a < G::A; a += 32) { opi[o * G::A + a] = 0.f; om[o * G::A + a] = 0; }
        }
    }
}

template <class G> static dim3 gk_grid_t(int n) { return dim3((unsigned)((n + gk_warps<G>() - 1) / gk_warps<G>())); }
#define gk_grid(n) gk_grid_t<G>(n)
#define GK_BLOCK (gk_warps<G>() * 32)
#define FINISH(nargs)                                                          \
    do {                                                                       \
        CKL();                                                                 \
        for (int i_ = 0; i_ < (nargs); i_++) if (tl_arg[i_].flush(st)) return 1; \
        if (any_staged(nargs)) CK(cudaStreamSynchronize(st));                  \
        return 0;                                                              \
    } while (0)

template <class G> static int game_init_t(int n, const uint64_t* seeds, int8_t* boards, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(seeds, sizeof(uint64_t) * n, st) || a[1].outbuf(boards, (size_t)n * G::S)) return 1;
    k_game_init<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<uint64_t>(), a[1].as<int8_t>());
    FINISH(2);
}
template <class G> static int game_valid_t(int n, const int8_t* boards, const int32_t* players, uint8_t* mask, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].in(players, sizeof(int) * n, st) || a[2].outbuf(mask, (size_t)n * G::A)) return 1;
    k_game_valid<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<int8_t>(), a[1].as<int>(), a[2].as<uint8_t>());
    FINISH(3);
}
template <class G> static int game_next_t(int n, const int8_t* boards, const int32_t* players, const int32_t* actions, const int64_t* seeds,
                                          const uint64_t* rng_keys, int8_t* out_boards, int32_t* out_next_player, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].in(players, sizeof(int) * n, st) || a[2].in(actions, sizeof(int) * n, st) ||
        a[3].in(seeds, sizeof(int64_t) * n, st) || a[4].in(rng_keys, sizeof(uint64_t) * n, st) ||
        a[5].outbuf(out_boards, (size_t)n * G::S) || a[6].outbuf(out_next_player, sizeof(int) * n)) return 1;
    k_game_next<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<int8_t>(), a[1].as<int>(), a[2].as<int>(), a[3].as<long long>(),
                                                    a[4].as<uint64_t>(), a[5].as<int8_t>(), a[6].as<int>());
    FINISH(7);
}
template <class G> static int game_ended_t(int n, const int8_t* boards, const int32_t* next_players, float* out, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].in(next_players, sizeof(int) * n, st) || a[2].outbuf(out, sizeof(float) * n * G::NP)) return 1;
    k_game_ended<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<int8_t>(), a[1].as<int>(), a[2].as<float>());
    FINISH(3);
}
template <class G> static int game_canonical_t(int n, const int8_t* boards, const int32_t* players, int8_t* out_boards, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].in(players, sizeof(int) * n, st) || a[2].outbuf(out_boards, (size_t)n * G::S)) return 1;
    k_game_canonical<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<int8_t>(), a[1].as<int>(), a[2].as<int8_t>());
    FINISH(3);
}
template <class G> static int game_round_score_t(int n, const int8_t* boards, int32_t* rounds, int32_t* scores, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].outbuf(rounds, sizeof(int) * n) || a[2].outbuf(scores, sizeof(int) * n * G::NP)) return 1;
    k_game_round_score<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<int8_t>(), a[1].as<int>(), a[2].as<int>());
    FINISH(3);
}
template <class G> static int game_symmetries_t(int n, const int8_t* boards, const float* pi, const uint8_t* mask, int8_t* out_boards,
                                                float* out_pi, uint8_t* out_mask, int32_t* out_k, cudaStream_t st) {
    Arg* a = tl_arg; const size_t K = G::MAX_SYM;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].in(pi, sizeof(float) * n * G::A, st) || a[2].in(mask, (size_t)n * G::A, st) ||
        a[3].outbuf(out_boards, n * K * G::S) || a[4].outbuf(out_pi, sizeof(float) * n * K * G::A) ||
        a[5].outbuf(out_mask, n * K * G::A) || a[6].outbuf(out_k, sizeof(int) * n)) return 1;
    k_game_symmetries<G><<<gk_grid(n), GK_BLOCK, 0, st>>>(n, a[0].as<int8_t>(), a[1].as<float>(), a[2].as<uint8_t>(), a[3].as<int8_t>(),
                                                          a[4].as<float>(), a[5].as<uint8_t>(), a[6].as<int>());
    FINISH(7);
}

extern "C" int azg_game_init(int game_id, int np, int n, const uint64_t* seeds, int8_t* boards, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DISPATCH(game_id, np, game_init_t<G>(n, seeds, boards, (cudaStream_t)stream));
}
extern "C" int azg_game_valid(int game_id, int np, int n, const int8_t* boards, const int32_t* players, uint8_t* mask, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DISPATCH(game_id, np, game_valid_t<G>(n, boards, players, mask, (cudaStream_t)stream));
}
extern "C" int azg_game_next(int game_id, int np, int n, const int8_t* boards, const int32_t* players, const int32_t* actions,
                             const int64_t* seeds, const uint64_t* rng_keys, int8_t* out_boards, int32_t* out_next_player, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    if (!actions) return fail("actions is NULL");
    DISPATCH(game_id, np, game_next_t<G>(n, boards, players, actions, seeds, rng_keys, out_boards, out_next_player, (cudaStream_t)stream));
}
extern "C" int azg_game_ended(int game_id, int np, int n, const int8_t* boards, const int32_t* next_players, float* out, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DISPATCH(game_id, np, game_ended_t<G>(n, boards, next_players, out, (cudaStream_t)stream));
}
extern "C" int azg_game_canonical(int game_id, int np, int n, const int8_t* boards, const int32_t* players, int8_t* out_boards, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DISPATCH(game_id, np, game_canonical_t<G>(n, boards, players, out_boards, (cudaStream_t)stream));
}
extern "C" int azg_game_round_score(int game_id, int np, int n, const int8_t* boards, int32_t* rounds, int32_t* scores, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DISPATCH(game_id, np, game_round_score_t<G>(n, boards, rounds, scores, (cudaStream_t)stream));
}
extern "C" int azg_game_symmetries(int game_id, int np, int n, const int8_t* boards, const float* pi, const uint8_t* mask,
                                   int8_t* out_boards, float* out_pi, uint8_t* out_mask, int32_t* out_k, void* stream) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DISPATCH(game_id, np, game_symmetries_t<G>(n, boards, pi, mask, out_boards, out_pi, out_mask, out_k, (cudaStream_t)stream));
}

// ------------------------------------------------------------------ net handle --------------------------
struct azg_net {
    int kind, game_id, np;
    V80Layout L; V80Chunks CK; V80DW DW; float* blob = nullptr;
    V80TCImg TI; float* img = nullptr;   // tensor-core operand images of the V80 token GEMMs (net_v80_tc.cuh)
    long long* prof = nullptr;            // optional phase timestamps of CTA 0 (AZG_V80_PROF=1; azg_net_prof)
    int v80_kernel = 1;                   // 1 = tcgen05 kernel (default), 0 = fp32 CUDA-core kernel (kept for A/B profiling; AZG_V80_KERNEL=fp32)
    V89Layout L89; V89Chunks CK89; V21Layout L21; TokMixLayout LTM; bool tokmix = false;   // tokmix: AzulNNet V84, SplendorNNet V80 for 3 / 4 players
    V89TCImg TI89; V89First F89; float* img89 = nullptr; float* res89 = nullptr; int v89_kernel = 1;   // tcgen05 trunk (default) or the fp32 CUDA-core kernel (AZG_V89_KERNEL=fp32, A/B runs)
    Scratch masks;                        // packed masks for the standalone forward
    unsigned long long launches = 0;
    bool attr_done = false; int n_sm = 0;  // launch attributes of this net's kernel on this net's device (set at the first launch)
};
__global__ void k_pack_masks(int n, int A, int MW, const uint8_t* mask, uint32_t* words) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= n) return;
    for (int k = 0; k < MW; k++) {
        const int a = lane + 32 * k;
        const unsigned w = __ballot_sync(FULL, a < A && mask[(size_t)j * A + a] != 0);
        if (lane == 0) words[(size_t)j * MW + k] = w;
    }
}
// Evaluate slots: list/count on device (engine) or identity (standalone). boards stride `bstride` bytes.
template <class G>
static int net_forward_dev(azg_net* net, const int* count_ptr, const int* list, const int8_t* boards, int bstride,
                           const uint32_t* masks, float* pi, float* v, int n_max, cudaStream_t st) {
    if (n_max <= 0) return 0;
    if (net->game_id != G::GAME_ID || net->np != G::NP) return fail("net was created for another game");
    if (net->kind == AZG_NET_HASH) {
        const int warps_per_block = 4;
        k_hashnet_forward<G::S, G::A, G::NP, G::MASK_WORDS><<<(n_max + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
            count_ptr, list, boards, bstride, masks, pi, v, n_max);
    } else if (net->kind == AZG_NET_SPLENDOR_V80) {
        if constexpr (G::GAME_ID == AZG_GAME_SPLENDOR && G::NP == 2) {
            if (net->v80_kernel == 1) {
                if (!net->attr_done) {                             // function attributes and the SM count belong to the net's device: kept in the handle
                    CK(cudaFuncSetAttribute(k_v80_tc<G::NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
                    int dev = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&net->n_sm, cudaDevAttrMultiProcessorCount, dev));
                    net->attr_done = true;
                }
                const int n_sm = net->n_sm;
                if ((reinterpret_cast<uintptr_t>(boards) & 3) || (bstride & 3)) return fail("V80 forward: boards must be 4-byte aligned");
                const int tiles = (n_max + TC_TB - 1) / TC_TB;
                k_v80_tc<G::NP><<<std::min(tiles, n_sm), TC_THREADS, TC_SMEM, st>>>(
                    net->blob, net->img, net->L, net->TI, net->DW, count_ptr, list, boards, bstride, masks, pi, v, n_max, net->prof);
            } else {
            constexpr size_t smem = v80_smem_bytes<G::ROWS, V80_TB>();
            if (!net->attr_done) { CK(cudaFuncSetAttribute(k_v80_forward<G::ROWS, G::NP, V80_TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); net->attr_done = true; }
            k_v80_forward<G::ROWS, G::NP, V80_TB><<<(n_max + V80_TB - 1) / V80_TB, V80_THREADS, smem, st>>>(
                net->blob, net->L, net->CK, net->DW, count_ptr, list, boards, bstride, masks, pi, v, n_max);
            }
        } else if constexpr (G::GAME_ID == AZG_GAME_SPLENDOR) {      // 3 / 4 players: 71 / 88 tokens, generic fp32 token-mixer kernel
            typedef typename std::conditional<G::NP == 3, TMS_V80_3P, TMS_V80_4P>::type SMX;
            auto kern = k_tokmix_forward<G::ROWS, 7, 81, G::NP, 3 * G::ROWS, G::ROWS, G::NP == 3 ? 56 : 64>;
            if (!net->attr_done) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMX::bytes())); net->attr_done = true; }
            kern<<<(n_max + TM_TB - 1) / TM_TB, TM_THREADS, SMX::bytes(), st>>>(net->blob, net->LTM, count_ptr, list, boards, bstride, masks, pi, v, n_max);
        } else return fail("SplendorNNet V80 only evaluates Splendor boards");
    } else if (net->kind == AZG_NET_SANTORINI_V89) {
        if constexpr (G::GAME_ID == AZG_GAME_SANTORINI) {
            if (net->v89_kernel == 1) {
                if (!net->attr_done) {
                    CK(cudaFuncSetAttribute(k_v89_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, T89_SMEM));
                    int dev = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&net->n_sm, cudaDevAttrMultiProcessorCount, dev));
                    net->attr_done = true;
                }
                if (!net->res89) CK(cudaMalloc(&net->res89, sizeof(float) * (size_t)net->n_sm * T89_RES_FLOATS));   // per-CTA residual scratch (stays in L2)
                const int tiles = (n_max + T89_TB - 1) / T89_TB;
                static const int grid_cap = getenv("AZG_V89_GRID") ? atoi(getenv("AZG_V89_GRID")) : 1 << 30;   // debug: fewer CTAs (is the kernel bound by L2 bandwidth?)
                k_v89_tc<<<std::min(std::min(tiles, net->n_sm), grid_cap), T89_THREADS, T89_SMEM, st>>>(net->blob, net->img89, net->res89, net->L89, net->TI89, net->F89, count_ptr, list, boards, bstride, masks, pi, v, n_max, net->prof);
            } else {
            constexpr size_t smem = v89_smem_bytes();
            if (!net->attr_done) { CK(cudaFuncSetAttribute(k_v89_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); net->attr_done = true; }
            k_v89_forward<<<(n_max + V89_TB - 1) / V89_TB, V89_THREADS, smem, st>>>(net->blob, net->L89, net->CK89, count_ptr, list, boards, bstride, masks, pi, v, n_max);
            }
        } else return fail("SantoriniNNet V89 only evaluates Santorini boards");
    } else if (net->kind == AZG_NET_ABALONE_V21) {
        if constexpr (G::GAME_ID == AZG_GAME_ABALONE) {
            static const size_t smem = v21_smem_bytes();
            if (!net->attr_done) {
                CK(cudaFuncSetAttribute(k_v21_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int dev = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&net->n_sm, cudaDevAttrMultiProcessorCount, dev));
                net->attr_done = true;
            }
            const V21Plan plan = v21_plan(n_max, net->n_sm);
            k_v21_forward<<<plan.n_big + plan.n_small, V21_THREADS, smem, st>>>(net->blob, net->L21, count_ptr, list, boards, bstride, masks, pi, v, n_max, plan.n_big);
        } else return fail("AbaloneNNet V21 only evaluates Abalone boards");
    } else if (net->kind == AZG_NET_AZUL_V84) {
        if constexpr (G::GAME_ID == AZG_GAME_AZUL) {
            auto kern = k_tokmix_forward<23, 6, 180, 2, 115, 46, 32>;
            if (!net->attr_done) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMS_V84::bytes())); net->attr_done = true; }
            kern<<<(n_max + TM_TB - 1) / TM_TB, TM_THREADS, TMS_V84::bytes(), st>>>(net->blob, net->LTM, count_ptr, list, boards, bstride, masks, pi, v, n_max);
        } else return fail("AzulNNet V84 only evaluates Azul boards");
    } else return fail("net kind not built");
    net->launches++;
    CKL();
    return 0;
}
extern "C" int azg_net_load(azg_net* net, const float* weights, size_t n_weights) {
    if (!net) return fail("net is NULL");
    if (net->kind == AZG_NET_HASH) return 0;
    if (net->kind == AZG_NET_ABALONE_V21) {
        const size_t need21 = v21_src_floats();
        if (!weights || n_weights != need21) return fail("V21 weights: expected " + std::to_string(need21) + " floats, got " + std::to_string(n_weights));
        std::vector<float> src(n_weights), dst((size_t)net->L21.total);
        CK(cudaMemcpy(src.data(), weights, n_weights * sizeof(float), cudaMemcpyDefault));
        v21_prepare(src.data(), net->L21, dst.data());
        CK(cudaMemcpy(net->blob, dst.data(), dst.size() * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    }
    if (net->tokmix) {
        const size_t need84 = tokmix_src_floats(net->LTM);
        if (!weights || n_weights != need84) return fail("token-mixer net weights: expected " + std::to_string(need84) + " floats, got " + std::to_string(n_weights));
        std::vector<float> src(n_weights), dst((size_t)net->LTM.total);
        CK(cudaMemcpy(src.data(), weights, n_weights * sizeof(float), cudaMemcpyDefault));
        tokmix_prepare(src.data(), net->LTM, dst.data());
        CK(cudaMemcpy(net->blob, dst.data(), dst.size() * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    }
    if (net->kind == AZG_NET_SANTORINI_V89) {
        const size_t need89 = v89_src_floats();
        if (!weights || n_weights != need89) return fail("V89 weights: expected " + std::to_string(need89) + " floats, got " + std::to_string(n_weights));
        std::vector<float> src(n_weights), dst((size_t)net->L89.total);
        CK(cudaMemcpy(src.data(), weights, n_weights * sizeof(float), cudaMemcpyDefault));
        v89_prepare(src.data(), net->L89, dst.data());
        CK(cudaMemcpy(net->blob, dst.data(), dst.size() * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<float> img((size_t)net->TI89.total);
        v89tc_prepare(dst.data(), net->L89, net->TI89, img.data());
        CK(cudaMemcpy(net->img89, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        for (int i = 0; i < 2 * 9 * 64; i++) net->F89.w[i] = dst[net->L89.conv[0] + i];          // first layer: kernel-parameter copy (constant bank)
        for (int i = 0; i < 64; i++) net->F89.b[i] = dst[net->L89.cbias[0] + i];
        for (int i = 0; i < 64; i++) { net->F89.wh[3 * i] = dst[net->L89.pi_w + 2 * i]; net->F89.wh[3 * i + 1] = dst[net->L89.pi_w + 2 * i + 1]; net->F89.wh[3 * i + 2] = dst[net->L89.v_w + i]; }
        return 0;
    }
    const size_t need = v80_src_floats(SP2::ROWS, net->np);
    if (!weights || n_weights != need) return fail("V80 weights: expected " + std::to_string(need) + " floats, got " + std::to_string(n_weights));
    std::vector<float> src(n_weights), dst((size_t)net->L.total);
    CK(cudaMemcpy(src.data(), weights, n_weights * sizeof(float), cudaMemcpyDefault));
    v80_prepare(src.data(), SP2::ROWS, net->np, net->L, dst.data());
    for (int k = 0; k < 3; k++) for (int i = 0; i < 49; i++) net->DW.w[k][i] = dst[(size_t)net->L.blk[k].wd + i];
    CK(cudaMemcpy(net->blob, dst.data(), dst.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> img((size_t)net->TI.total);
    v80tc_prepare(dst.data(), net->L, net->TI, img.data());
    CK(cudaMemcpy(net->img, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int azg_net_create(int net_kind, int game_id, int np, const float* weights, size_t n_weights, azg_net** out) {
    if (!out) return fail("out is NULL");
    azg_game_info_t gi;
    if (require_device() || azg_game_info(game_id, np, &gi)) return 1;
    if (net_kind != AZG_NET_HASH && net_kind != AZG_NET_SPLENDOR_V80 && net_kind != AZG_NET_SANTORINI_V89 && net_kind != AZG_NET_ABALONE_V21 && net_kind != AZG_NET_AZUL_V84)
        return fail("unknown net kind (built: 0=hash test net, 80=Splendor V80, 89=Santorini V89, 21=Abalone V21, 84=Azul V84)");
    if (net_kind == AZG_NET_AZUL_V84 && game_id != AZG_GAME_AZUL) return fail("AzulNNet V84 only evaluates Azul boards");
    if (net_kind == AZG_NET_ABALONE_V21 && game_id != AZG_GAME_ABALONE) return fail("AbaloneNNet V21 only evaluates Abalone boards");
    if (net_kind == AZG_NET_SPLENDOR_V80 && game_id != AZG_GAME_SPLENDOR) return fail("SplendorNNet V80 only evaluates Splendor boards");
    if (net_kind == AZG_NET_SANTORINI_V89 && game_id != AZG_GAME_SANTORINI) return fail("SantoriniNNet V89 only evaluates Santorini boards");
    azg_net* net = new azg_net(); net->kind = net_kind; net->game_id = game_id; net->np = np;
    if (net_kind == AZG_NET_SPLENDOR_V80 && np != 2) {
        net->tokmix = true; net->LTM = tokmix_layout(80, gi.state_rows, 7, 81, np);
        if (cudaMalloc(&net->blob, sizeof(float) * (size_t)net->LTM.total) != cudaSuccess) { delete net; return fail("cudaMalloc weights failed"); }
        if (azg_net_load(net, weights, n_weights)) { cudaFree(net->blob); delete net; return 1; }
    } else if (net_kind == AZG_NET_SPLENDOR_V80) {
        net->L = v80_layout(SP2::ROWS, np); net->CK = v80_chunks(net->L); memset(&net->DW, 0, sizeof(net->DW));
        net->TI = v80tc_layout();
        const char* kv = getenv("AZG_V80_KERNEL");
        net->v80_kernel = (kv && !strcmp(kv, "fp32")) ? 0 : 1;
        if (getenv("AZG_V80_PROF")) { if (cudaMalloc(&net->prof, 704 * sizeof(long long)) != cudaSuccess) net->prof = nullptr; else cudaMemset(net->prof, 0, 704 * sizeof(long long)); }
        if (cudaMalloc(&net->blob, sizeof(float) * (size_t)net->L.total) != cudaSuccess) { delete net; return fail("cudaMalloc weights failed"); }
        if (cudaMalloc(&net->img, sizeof(float) * (size_t)net->TI.total) != cudaSuccess) { cudaFree(net->blob); delete net; return fail("cudaMalloc weight images failed"); }
        if (azg_net_load(net, weights, n_weights)) { cudaFree(net->blob); cudaFree(net->img); delete net; return 1; }
    } else if (net_kind == AZG_NET_ABALONE_V21) {
        net->L21 = v21_layout();
        if (cudaMalloc(&net->blob, sizeof(float) * (size_t)net->L21.total) != cudaSuccess) { delete net; return fail("cudaMalloc weights failed"); }
        if (azg_net_load(net, weights, n_weights)) { cudaFree(net->blob); delete net; return 1; }
    } else if (net_kind == AZG_NET_AZUL_V84) {
        net->tokmix = true; net->LTM = tokmix_layout(84, 23, 6, 180, 2);
        if (cudaMalloc(&net->blob, sizeof(float) * (size_t)net->LTM.total) != cudaSuccess) { delete net; return fail("cudaMalloc weights failed"); }
        if (azg_net_load(net, weights, n_weights)) { cudaFree(net->blob); delete net; return 1; }
    } else if (net_kind == AZG_NET_SANTORINI_V89) {
        net->L89 = v89_layout(); net->CK89 = v89_chunks(net->L89); net->TI89 = v89tc_layout();
        { const char* kv = getenv("AZG_V89_KERNEL"); net->v89_kernel = (kv && !strcmp(kv, "fp32")) ? 0 : 1; }
        if (getenv("AZG_V89_PROF")) { if (cudaMalloc(&net->prof, 64 * sizeof(long long)) != cudaSuccess) net->prof = nullptr; else cudaMemset(net->prof, 0, 64 * sizeof(long long)); }
        if (cudaMalloc(&net->blob, sizeof(float) * (size_t)net->L89.total) != cudaSuccess) { delete net; return fail("cudaMalloc weights failed"); }
        if (cudaMalloc(&net->img89, sizeof(float) * (size_t)net->TI89.total) != cudaSuccess) { cudaFree(net->blob); delete net; return fail("cudaMalloc weight images failed"); }
        if (azg_net_load(net, weights, n_weights)) { cudaFree(net->blob); cudaFree(net->img89); delete net; return 1; }
    }
    *out = net; return 0;
}
extern "C" int azg_debug_selprof(unsigned long long* out8) {      // debug: see g_selprof (only filled when built with -DAZG_SEL_PROF)
    CK(cudaDeviceSynchronize()); CK(cudaMemcpyFromSymbol(out8, azg::g_selprof, 8 * sizeof(unsigned long long))); CK(cudaMemcpyFromSymbol(out8 + 8, azg::g_selprof2, 8 * sizeof(unsigned long long))); return 0;
}
extern "C" int azg_net_prof_ctas(azg_net* net, long long* out640) {   // debug (V80 only): per CTA {entry, prologue done, exit} in globaltimer ns + SM id
    if (!net || !net->prof || net->kind != AZG_NET_SPLENDOR_V80) return fail("profiling not enabled (AZG_V80_PROF=1)");
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(out640, net->prof + 64, 640 * sizeof(long long), cudaMemcpyDeviceToHost)); return 0;
}
extern "C" int azg_net_prof(azg_net* net, long long* out64) {      // debug: phase timestamps (SM clock) of CTA 0's first tiles
    if (!net || !net->prof) return fail("profiling not enabled (AZG_V80_PROF=1 / AZG_V89_PROF=1)");
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(out64, net->prof, 64 * sizeof(long long), cudaMemcpyDeviceToHost)); return 0;
}
extern "C" int azg_net_destroy(azg_net* net) { if (net) { if (net->blob) cudaFree(net->blob); if (net->img) cudaFree(net->img); if (net->img89) cudaFree(net->img89); if (net->res89) cudaFree(net->res89); delete net; } return 0; }
template <class G> static int net_forward_t(azg_net* net, int n, const int8_t* boards, const uint8_t* mask, float* pi, float* v, cudaStream_t st) {
    Arg* a = tl_arg;
    if (a[0].in(boards, (size_t)n * G::S, st) || a[1].in(mask, (size_t)n * G::A, st) || a[2].outbuf(pi, sizeof(float) * n * G::A) ||
        a[3].outbuf(v, sizeof(float) * n * G::NP)) return 1;
    if (net->masks.ensure(sizeof(uint32_t) * (size_t)n * G::MASK_WORDS)) return 1;
    k_pack_masks<<<(n + 3) / 4, 128, 0, st>>>(n, G::A, G::MASK_WORDS, a[1].as<uint8_t>(), (uint32_t*)net->masks.p);
    if (net_forward_dev<G>(net, nullptr, nullptr, a[0].as<int8_t>(), G::S, (const uint32_t*)net->masks.p, a[2].as<float>(), a[3].as<float>(), n, st)) return 1;
    FINISH(4);
}
extern "C" int azg_net_forward(azg_net* net, int n, const int8_t* boards, const uint8_t* mask, float* pi, float* v, void* stream) {
    if (!net) return fail("net is NULL");
    if (n <= 0) return 0;
    DISPATCH(net->game_id, net->np, net_forward_t<G>(net, n, boards, mask, pi, v, (cudaStream_t)stream));
}

// ------------------------------------------------------------------ engine ------------------------------
template <class G>
__global__ void k_load_roots(Dev<G> d, int n, const int8_t* roots, const uint8_t* full, int sims_full, int sims_fast) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= d.n_games) return;
    if (g < n) {
        for (int i = lane; i < G::SP; i += 32) d.root[(size_t)g * G::SP + i] = i < G::S ? roots[(size_t)g * G::S + i] : (int8_t)0;
        if (lane == 0) {                                          // full[g]: 1 = full search, 0 = fast search (MCTS.py:58-59), 2 = no search for this slot
            const uint8_t fv = full ? full[g] : (uint8_t)1; const bool f = fv == 1;
            d.full[g] = f; d.n_sims[g] = fv >= 2 ? 0 : (f ? sims_full : sims_fast); d.root_node[g] = 0;
        }
    } else if (lane == 0) d.n_sims[g] = 0;
}
static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
enum { PK_SELECT = 0, PK_NET = 1, PK_BACKUP = 2, PK_OTHER = 3 };

struct azg_engine {                      // game-independent face of the engine (the C ABI holds this)
    virtual ~azg_engine() {}
    virtual int reset(int game) = 0;
    virtual int search(int n, const int8_t* roots, const uint8_t* full_search, const double* noise, int32_t* out_counts, int32_t* out_raw,
                       float* out_q, cudaStream_t st) = 0;
    virtual int selfplay(int min_episodes, int max_moves, cudaStream_t st) = 0;
    virtual int selfplay_inject(const azg_selfplay_inject* inj) = 0;
    virtual int selfplay_state(int8_t* boards, int32_t* players, int32_t* plies, int32_t* active) = 0;
    virtual int node(int n, const int32_t* slots, const int8_t* boards, int32_t* found, float* es, uint8_t* vs, float* ps, int32_t* ns, double* qsa,
                     int32_t* nsa, int32_t* round, float* qs, cudaStream_t st) = 0;
    virtual int examples(int cap, int8_t* boards, float* pi, float* z, uint8_t* valids, float* q, int32_t* out_n) = 0;
    virtual int examples_pending(int32_t* out_n) = 0;
    virtual int stats(int64_t* out16) = 0;
    virtual int profile(int enable) = 0;
    virtual int kernel_times(double* out8) = 0;
};

template <class G>
struct EngineT : azg_engine {
    azg_engine_cfg cfg; azg_net* net = nullptr;
    Dev<G> d; SelfPlay<G> sp;
    std::vector<void*> allocs;
    unsigned long long launches = 0;
    int sims_full = 0, sims_fast = 0;
    bool sp_ready = false;
    bool ragged_env = true, rag_started = false;   // ragged self-play (per-slot move boundaries), see selfplay_ragged
    bool profiling = false; int prof_every = 1; bool prof_now = true; std::vector<cudaEvent_t> ev; size_t ev_used = 0; std::vector<int> ev_kind; double prof_ms[4] = {0, 0, 0, 0}; long long prof_n[4] = {0, 0, 0, 0};

    template <class T> int alloc(T** p, size_t n, bool zero = true) {
        void* q = nullptr;
        if (cudaMalloc(&q, sizeof(T) * n) != cudaSuccess) return fail("cudaMalloc failed for " + std::to_string(sizeof(T) * n) + " bytes (reduce n_games / node_cap / edge_cap)");
        if (zero && cudaMemset(q, 0, sizeof(T) * n) != cudaSuccess) return fail("cudaMemset failed");
        allocs.push_back(q); *p = (T*)q; return 0;
    }
    ~EngineT() override {
        for (void* p : allocs) cudaFree(p);
        for (cudaEvent_t x : ev) cudaEventDestroy(x);
    }
    dim3 grid() const { return dim3((unsigned)((d.n_games + sel_warps<G>() - 1) / sel_warps<G>())); }

    int create(const azg_engine_cfg* c, azg_net* n_) {
        cfg = *c; net = n_;
        const int NG = c->n_games;
        if (c->universes < 0 || c->universes > 8) return fail("universes must be in [0, 8]");
        int node_cap = c->node_cap, edge_cap = c->edge_cap;
        const int U0 = std::max(c->universes, 1);
        constexpr int EDGE_FACTOR = G::EDGE_FACTOR;
        if (node_cap <= 0) {
            // default: room for the nodes that survive tree reuse, bounded by 60 % of free HBM
            size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
            const double per_node = 48.0 + G::SP + 8.0 + 4.0 * U0 + EDGE_FACTOR * (17.0 + 4.0 * U0) + 2 * 8.0;
            double fit = 0.6 * (double)free_b / (double)NG / per_node;
            node_cap = (int)std::min<double>(fit, 16.0 * c->numMCTSSims + 1024);
            node_cap = std::max(node_cap, c->numMCTSSims + 64);
        }
        if (edge_cap <= 0) edge_cap = node_cap * EDGE_FACTOR;
        if (edge_cap >= (1 << 24)) return fail("edge_cap must be < 2^24");
        cfg.node_cap = node_cap; cfg.edge_cap = edge_cap;
        d.n_games = NG; d.node_cap = node_cap; d.edge_cap = edge_cap; d.ht_cap = next_pow2(2 * node_cap);
        d.universes = c->universes; d.U = U0; d.forced_playouts = c->forced_playouts; d.dirichlet_noise = c->dirichlet_noise;
        d.cpuct = c->cpuct; d.fpu = c->fpu; d.dir_alpha = c->dirichletAlpha; d.temp2 = c->temperature[2]; d.seed = c->seed;
        d.noise = nullptr; d.noise_gstride = G::A; d.noise_ply = nullptr; d.game_base = c->first_game;
        { const char* rv = getenv("AZG_TREE_REPLAY"); d.replay = (rv && rv[0] == '0') ? 0 : 1; }
        sims_full = c->numMCTSSims; sims_fast = c->ratio_fullMCTS > 0 ? c->numMCTSSims / c->ratio_fullMCTS : c->numMCTSSims;
        int bad = 0;
        bad |= alloc(&d.nodes, (size_t)NG * node_cap, false); bad |= alloc(&d.keys, (size_t)NG * node_cap, false); bad |= alloc(&d.edges, (size_t)NG * edge_cap, false);
        bad |= alloc(&d.acts, (size_t)NG * edge_cap, false); bad |= alloc(&d.ht, (size_t)NG * d.ht_cap);
        bad |= alloc(&d.n_nodes, NG); bad |= alloc(&d.n_edges, NG);
        bad |= alloc(&d.child, (size_t)NG * edge_cap * d.U, false); bad |= alloc(&d.boards, (size_t)NG * node_cap * G::SP, false);
        bad |= alloc(&d.bestlink, (size_t)NG * node_cap * d.U, false);
        bad |= alloc(&d.ord_cnt, 64); bad |= alloc(&d.ord_list, (size_t)64 * NG, false);
        bad |= alloc(&d.remap, (size_t)NG * node_cap, false); bad |= alloc(&d.gcq, (size_t)NG * node_cap, false); bad |= alloc(&d.root_node, NG); bad |= alloc(&d.leaf_link, NG);
        bad |= alloc(&d.root, (size_t)NG * G::SP); bad |= alloc(&d.n_sims, NG); bad |= alloc(&d.full, NG); bad |= alloc(&d.move_ctr, NG);
        bad |= alloc(&d.path, (size_t)NG * d.U * G::MAX_DEPTH); bad |= alloc(&d.path_len, (size_t)NG * d.U); bad |= alloc(&d.leaf_kind, NG);
        bad |= alloc(&d.leaf_key, (size_t)2 * NG); bad |= alloc(&d.leaf_v, (size_t)NG * G::NP); bad |= alloc(&d.leaf_mask, (size_t)NG * G::MASK_WORDS);
        bad |= alloc(&d.leaf_round, NG);
        bad |= alloc(&d.noise_scr, (size_t)NG * G::A, false);
        bad |= alloc(&d.nn_in, (size_t)NG * G::SP); bad |= alloc(&d.nn_pi, (size_t)NG * G::A); bad |= alloc(&d.nn_v, (size_t)NG * G::NP);
        bad |= alloc(&d.nn_list, NG); bad |= alloc(&d.nn_count, 1); bad |= alloc(&d.stats, (size_t)NG * ST_N);
        bad |= alloc(&d.sim_idx, NG); bad |= alloc(&d.turn_list, NG); bad |= alloc(&d.turn_count, 1); d.ragged = 0;
        { const char* rv = getenv("AZG_RAGGED"); ragged_env = !(rv && rv[0] == '0'); }
        if (bad) return bad;
        return 0;
    }
    int reset(int game) override {
        if (game >= d.n_games) return fail("game index out of range");
        k_reset<G><<<game >= 0 ? 8 : 592, 256>>>(d, game);
        launches++;
        CKL(); return 0;
    }

    // ---- optional per-kernel timing: an event before and after each launch, drained by prof_drain() ----
    void prof_mark(int kind, cudaStream_t st) {              // kind >= 0: start of a launch of that kind; -1: end
        if (!profiling || !prof_now) return;
        if (ev_used == ev.size()) { cudaEvent_t x; cudaEventCreate(&x); ev.push_back(x); ev_kind.push_back(0); }
        ev_kind[ev_used] = kind;
        cudaEventRecord(ev[ev_used++], st);
    }
    int prof_drain() {
        if (ev_used == 0) return 0;
        CK(cudaEventSynchronize(ev[ev_used - 1]));
        for (size_t i = 0; i + 1 < ev_used; i++) {
            const int k = ev_kind[i];
            if (k < 0) continue;
            float ms = 0.f; CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            prof_ms[k] += ms; prof_n[k]++;
        }
        ev_used = 0; return 0;
    }
    int profile(int enable) override {                        // enable = N > 1: only every N-th lock-step simulation is timed (the events of the
        if (prof_drain()) return 1;                           // others are not recorded: 6 event records per simulation cost ~4 % of the loop)
        profiling = enable != 0; prof_every = enable > 1 ? enable : 1; prof_now = true;
        if (enable) for (int k = 0; k < 4; k++) { prof_ms[k] = 0; prof_n[k] = 0; }
        return 0;
    }
    int kernel_times(double* out8) override {
        if (prof_drain()) return 1;
        out8[0] = prof_ms[PK_SELECT]; out8[1] = prof_ms[PK_NET]; out8[2] = prof_ms[PK_BACKUP]; out8[3] = prof_ms[PK_OTHER];
        out8[4] = (double)prof_n[PK_SELECT]; out8[5] = (double)prof_n[PK_SELECT]; out8[6] = (double)prof_n[PK_NET]; out8[7] = (double)prof_n[PK_BACKUP];
        return 0;
    }

    // One lock-step simulation for every game: select -> batched leaf evaluation -> expand + backup.
    int step(int s, cudaStream_t st) {
        const int NG = d.n_games;
        prof_now = (s % prof_every) == 0;                     // sampled timing; the per-move kernels (begin / end / gc) are always timed
        struct Unsample { bool& f; ~Unsample() { f = true; } } unsample{prof_now};
        prof_mark(PK_SELECT, st);
        k_select<G><<<(unsigned)((NG + selk_warps<G>() - 1) / selk_warps<G>()), selk_warps<G>() * 32, 0, st>>>(d, s);
        prof_mark(PK_NET, st);
        if (net_forward_dev<G>(net, d.nn_count, d.nn_list, d.nn_in, G::SP, d.leaf_mask, d.nn_pi, d.nn_v, NG, st)) return 1;
        prof_mark(PK_BACKUP, st);
        k_backup<G><<<(unsigned)((NG + bak_warps<G>() - 1) / bak_warps<G>()), bak_warps<G>() * 32, 0, st>>>(d, s);
        prof_mark(-1, st);
        launches += 3;
        return 0;
    }
    int gc(int sims, cudaStream_t st) {
        prof_mark(PK_OTHER, st);
        k_gc<G><<<grid(), sel_warps<G>() * 32, 0, st>>>(d, sims + 2, (sims + 2) * G::MAX_LEGAL, 0);
        prof_mark(-1, st);
        launches++;
        return 0;
    }

    int search(int n, const int8_t* roots, const uint8_t* full_search, const double* noise, int32_t* out_counts, int32_t* out_raw,
               float* out_q, cudaStream_t st) override {
        if (n <= 0 || n > d.n_games) return fail("n must be in [1, n_games]");
        if (!roots || !out_counts) return fail("roots / out_counts is NULL");
        Arg* a = tl_arg; const int NG = d.n_games;
        if (a[0].in(roots, (size_t)n * G::S, st) || a[1].in(full_search, (size_t)n, st) || a[2].in(noise, sizeof(double) * n * G::A, st) ||
            a[3].outbuf(out_counts, sizeof(int) * n * G::A) || a[4].outbuf(out_raw, sizeof(int) * n * G::A) || a[5].outbuf(out_q, sizeof(float) * n * G::NP)) return 1;
        CK(cudaMemsetAsync(d.nn_count, 0, sizeof(int), st));
        k_load_roots<G><<<(NG + 3) / 4, 128, 0, st>>>(d, n, a[0].as<int8_t>(), a[1].as<uint8_t>(), sims_full, sims_fast);
        launches++;
        d.noise = a[2].as<double>();
        // without host knowledge of the flags run the longer budget; finished games idle (k_select early-out)
        int steps = sims_full;
        if (full_search && !is_device_ptr(full_search)) {
            bool any_full = false, any_fast = false;
            for (int i = 0; i < n; i++) { any_full |= full_search[i] == 1; any_fast |= full_search[i] == 0; }
            steps = any_full ? sims_full : (any_fast ? sims_fast : 0);
        }
        gc(steps, st);
        for (int s = 0; s < steps; s++) if (step(s, st)) return 1;
        k_finish<G><<<(n + sel_warps<G>() - 1) / sel_warps<G>(), sel_warps<G>() * 32, 0, st>>>(d, n, a[3].as<int>(), a[4].as<int>(), a[5].as<float>());
        launches++;
        d.noise = nullptr;
        if (profiling) { CKL(); if (prof_drain()) return 1; }
        FINISH(6);
    }

    int node(int n, const int32_t* slots, const int8_t* boards, int32_t* found, float* es, uint8_t* vs, float* ps, int32_t* ns, double* qsa,
             int32_t* nsa, int32_t* round, float* qs, cudaStream_t st) override {
        if (n <= 0) return 0;
        if (!boards || !found) return fail("boards / found is NULL");
        Arg* a = tl_arg; const size_t N = (size_t)n;
        if (a[0].in(slots, sizeof(int) * N, st) || a[1].in(boards, N * G::S, st) || a[2].outbuf(found, sizeof(int) * N) || a[3].outbuf(es, sizeof(float) * N * G::NP) ||
            a[4].outbuf(vs, N * G::A) || a[5].outbuf(ps, sizeof(float) * N * G::A) || a[6].outbuf(ns, sizeof(int) * N) || a[7].outbuf(qsa, sizeof(double) * N * G::A) ||
            a[8].outbuf(nsa, sizeof(int) * N * G::A) || a[9].outbuf(round, sizeof(int) * N) || a[10].outbuf(qs, sizeof(float) * N)) return 1;
        k_node_query<G><<<(n + sel_warps<G>() - 1) / sel_warps<G>(), sel_warps<G>() * 32, 0, st>>>(d, n, a[0].as<int>(), a[1].as<int8_t>(), a[2].as<int>(), a[3].as<float>(),
            a[4].as<uint8_t>(), a[5].as<float>(), a[6].as<int>(), a[7].as<double>(), a[8].as<int>(), a[9].as<int>(), a[10].as<float>());
        launches++;
        FINISH(11);
    }
    int selfplay_state(int8_t* boards, int32_t* players, int32_t* plies, int32_t* active) override {
        if (selfplay_setup()) return 1;
        const int NG = d.n_games;
        CK(cudaDeviceSynchronize());
        if (boards) CK(cudaMemcpy2D(boards, G::S, sp.board, G::SP, G::S, NG, cudaMemcpyDefault));
        if (players) CK(cudaMemcpy(players, sp.player, sizeof(int) * NG, cudaMemcpyDefault));
        if (plies) CK(cudaMemcpy(plies, sp.ply, sizeof(int) * NG, cudaMemcpyDefault));
        if (active) CK(cudaMemcpy(active, sp.active, sizeof(int) * NG, cudaMemcpyDefault));
        return 0;
    }
    int stats(int64_t* out16) override {
        const int NG = d.n_games;
        std::vector<unsigned long long> h((size_t)NG * ST_N);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d.stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int k = 0; k < AZG_N_STATS; k++) out16[k] = 0;
        for (int g = 0; g < NG; g++)
            for (int k = 0; k < ST_N && k < AZG_N_STATS; k++) {
                if (k == ST_MAXNODES) out16[k] = std::max<int64_t>(out16[k], (int64_t)h[(size_t)g * ST_N + k]);
                else out16[k] += (int64_t)h[(size_t)g * ST_N + k];
            }
        out16[12] = (int64_t)(launches + (net ? net->launches : 0));
        out16[14] = d.node_cap;
        if (sp_ready) { unsigned long long c[8]; CK(cudaMemcpy(c, sp.counters, sizeof(c), cudaMemcpyDeviceToHost)); out16[18] = (int64_t)c[2]; }
        return 0;
    }

    // ---- Coach.executeEpisodes (Coach.py:86-148) ----
    int selfplay_setup() {
        if (sp_ready) return 0;
        const int NG = d.n_games; const azg_engine_cfg& c = cfg;
        sp.max_ply = G::MAX_MOVES; sp.ex_cap = NG * G::MAX_MOVES;
        sp.prob_full = c.prob_fullMCTS; sp.t_begin = c.temperature[0]; sp.t_end = c.temperature[1]; sp.half_life = c.tempThreshold;
        int bad = 0; const size_t M = (size_t)NG * sp.max_ply;
        bad |= alloc(&sp.board, (size_t)NG * G::SP); bad |= alloc(&sp.player, NG); bad |= alloc(&sp.ply, NG); bad |= alloc(&sp.active, NG);
        bad |= alloc(&sp.games_started, NG);
        bad |= alloc(&sp.st_board, M * G::S, false); bad |= alloc(&sp.st_pi, M * G::A, false); bad |= alloc(&sp.st_mask, M * G::MASK_WORDS, false);
        bad |= alloc(&sp.st_q, M * G::NP, false); bad |= alloc(&sp.st_player, M, false); bad |= alloc(&sp.st_count, NG);
        bad |= alloc(&sp.ex_board, (size_t)sp.ex_cap * G::S, false); bad |= alloc(&sp.ex_pi, (size_t)sp.ex_cap * G::A, false);
        bad |= alloc(&sp.ex_z, (size_t)sp.ex_cap * G::NP, false); bad |= alloc(&sp.ex_valid, (size_t)sp.ex_cap * G::A, false);
        bad |= alloc(&sp.ex_q, (size_t)sp.ex_cap * G::NP, false); bad |= alloc(&sp.ex_count, 1); bad |= alloc(&sp.counters, 8);
        if (bad) return 1;
        sp_ready = true; return 0;
    }
    // ---- injected randomness (parity tests): see azg_selfplay_inject in azg.h ----
    Scratch inj_buf[5];
    int selfplay_inject(const azg_selfplay_inject* inj) override {
        if (selfplay_setup()) return 1;
        const int NG = d.n_games;
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(sp.active, 0, sizeof(int) * NG)); CK(cudaMemset(sp.games_started, 0, sizeof(unsigned) * NG)); CK(cudaMemset(sp.ply, 0, sizeof(int) * NG));
        CK(cudaMemset(sp.st_count, 0, sizeof(int) * NG)); CK(cudaMemset(sp.player, 0, sizeof(int) * NG));
        sp.inj_P = 0; sp.inj_init = nullptr; sp.inj_u_full = sp.inj_u_move = nullptr; sp.inj_seed = nullptr; sp_noise = nullptr;
        rag_started = false;
        if (!inj) return 0;
        if (inj->n_plies <= 0 || !inj->init_boards || !inj->u_full || !inj->u_move || !inj->chance_seed) return fail("azg_selfplay_inject: n_plies must be positive and init_boards / u_full / u_move / chance_seed non-NULL");
        const size_t P = (size_t)inj->n_plies;
        const void* src[5] = {inj->init_boards, inj->u_full, inj->u_move, inj->chance_seed, inj->noise};
        const size_t bytes[5] = {(size_t)NG * G::S, sizeof(double) * NG * P, sizeof(double) * NG * P, sizeof(long long) * NG * P, sizeof(double) * NG * P * G::A};
        for (int i = 0; i < 5; i++) {
            if (!src[i]) continue;
            if (inj_buf[i].ensure(bytes[i])) return 1;
            CK(cudaMemcpy(inj_buf[i].p, src[i], bytes[i], cudaMemcpyDefault));
        }
        sp.inj_P = (int)P; sp.inj_init = (const int8_t*)inj_buf[0].p; sp.inj_u_full = (const double*)inj_buf[1].p; sp.inj_u_move = (const double*)inj_buf[2].p;
        sp.inj_seed = (const long long*)inj_buf[3].p; sp_noise = inj->noise ? (const double*)inj_buf[4].p : nullptr;
        return 0;
    }
    const double* sp_noise = nullptr;          // injected per-(slot, ply) Dirichlet draws of self-play
    // Ragged self-play: with playout-cap randomisation (MCTS.py:58-59: numMCTSSims simulations with probability prob_fullMCTS, else
    // numMCTSSims / ratio_fullMCTS) a lock-step ply leaves the fast-search slots idle for most of its launches. Here every slot
    // carries its own simulation index; a slot whose budget is spent makes its move and starts the next search at the next launch
    // (k_sp_turn), so every launch works on all n_games trees. Same games as the lock-step schedule (every RNG stream is keyed by
    // slot, game and ply), only the order in which slots reach their plies differs. max_moves counts plies per slot on average.
    bool use_ragged() const { return ragged_env && !sp.inj_P && cfg.prob_fullMCTS > 0.0 && cfg.prob_fullMCTS < 1.0 && sims_fast < sims_full; }
    int selfplay_ragged(int min_episodes, int max_moves, cudaStream_t st) {
        const int NG = d.n_games;
        unsigned long long start[8], now[8];
        CK(cudaMemcpyAsync(start, sp.counters, sizeof(start), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        d.ragged = 1;
        struct Restore { Dev<G>& d; ~Restore() { d.ragged = 0; } } restore{d};
        if (!rag_started) {                                       // every slot starts its first search together
            prof_mark(PK_OTHER, st);
            k_sp_begin<G><<<grid(), sel_warps<G>() * 32, 0, st>>>(d, sp, sims_full, sims_fast);
            prof_mark(-1, st);
            launches++;
            gc(sims_full, st);
            rag_started = true;
        }
        const int chunk = std::max(16, std::min(sims_fast, 128));
        const unsigned turn_grid = (unsigned)std::min((NG + sel_warps<G>() - 1) / sel_warps<G>(), 296);
        for (int ln = 0;;) {
            for (int k = 0; k < chunk; k++, ln++) {
                if (step(ln, st)) return 1;
                prof_mark(PK_OTHER, st);
                k_sp_turn<G><<<turn_grid, sel_warps<G>() * 32, 0, st>>>(d, sp, sims_full, sims_fast);
                prof_mark(-1, st);
                launches++;
            }
            CKL();
            if (profiling && prof_drain()) return 1;
            int ring = 0;
            CK(cudaMemcpyAsync(now, sp.counters, sizeof(now), cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(&ring, sp.ex_count, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (min_episodes > 0 && (long long)(now[0] - start[0]) >= min_episodes) break;
            if (max_moves > 0 && (long long)(now[3] - start[3]) >= (long long)max_moves * NG) break;
            if (ring > sp.ex_cap / 2) break;
            if (max_moves <= 0 && min_episodes <= 0) break;
        }
        return 0;
    }
    int selfplay(int min_episodes, int max_moves, cudaStream_t st) override {
        if (selfplay_setup()) return 1;
        if (use_ragged()) return selfplay_ragged(min_episodes, max_moves, st);
        d.noise = sp_noise; d.noise_gstride = sp_noise ? (size_t)sp.inj_P * G::A : (size_t)G::A; d.noise_ply = sp_noise ? sp.ply : nullptr;
        struct Restore { Dev<G>& d; ~Restore() { d.noise = nullptr; d.noise_gstride = G::A; d.noise_ply = nullptr; } } restore{d};
        unsigned long long start[8], now[8];
        CK(cudaMemcpyAsync(start, sp.counters, sizeof(start), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        for (int mv = 0; max_moves <= 0 || mv < max_moves; mv++) {
            prof_mark(PK_OTHER, st);
            k_sp_begin<G><<<grid(), sel_warps<G>() * 32, 0, st>>>(d, sp, sims_full, sims_fast);
            prof_mark(-1, st);
            launches++;
            const int steps = (cfg.prob_fullMCTS > 0.0) ? sims_full : sims_fast;
            gc(steps, st);
            for (int s = 0; s < steps; s++) if (step(s, st)) return 1;
            prof_mark(PK_OTHER, st);
            k_sp_end<G><<<grid(), sel_warps<G>() * 32, 0, st>>>(d, sp);
            prof_mark(-1, st);
            launches++;
            CKL();
            if (profiling && prof_drain()) return 1;
            if (min_episodes > 0) {
                int ring = 0;
                CK(cudaMemcpyAsync(now, sp.counters, sizeof(now), cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(&ring, sp.ex_count, sizeof(int), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if ((long long)(now[0] - start[0]) >= min_episodes) break;
                if (ring > sp.ex_cap / 2) break;                  // the caller drains the ring (azg_engine_examples) and calls again
            }
            if (max_moves <= 0 && min_episodes <= 0) break;
        }
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    int examples_pending(int32_t* out_n) override {
        *out_n = 0;
        if (!sp_ready) return 0;
        CK(cudaDeviceSynchronize());
        int count = 0; CK(cudaMemcpy(&count, sp.ex_count, sizeof(int), cudaMemcpyDeviceToHost));
        *out_n = std::min(count, sp.ex_cap); return 0;
    }
    int examples(int cap, int8_t* boards, float* pi, float* z, uint8_t* valids, float* q, int32_t* out_n) override {
        *out_n = 0;
        if (!sp_ready) return 0;
        CK(cudaDeviceSynchronize());
        int count = 0; CK(cudaMemcpy(&count, sp.ex_count, sizeof(int), cudaMemcpyDeviceToHost));
        count = std::min(count, sp.ex_cap);
        const int m = std::min(count, cap);
        if (m > 0) {
            if (!boards || !pi || !z || !valids || !q) return fail("NULL output buffer");
            CK(cudaMemcpy(boards, sp.ex_board, (size_t)m * G::S, cudaMemcpyDefault));
            CK(cudaMemcpy(pi, sp.ex_pi, sizeof(float) * (size_t)m * G::A, cudaMemcpyDefault));
            CK(cudaMemcpy(z, sp.ex_z, sizeof(float) * (size_t)m * G::NP, cudaMemcpyDefault));
            CK(cudaMemcpy(valids, sp.ex_valid, (size_t)m * G::A, cudaMemcpyDefault));
            CK(cudaMemcpy(q, sp.ex_q, sizeof(float) * (size_t)m * G::NP, cudaMemcpyDefault));
        }
        const int rest = count - m;
        if (rest > 0) {                                   // keep what did not fit: slide it to the front through a temporary
            Scratch tmp;
            auto slide = [&](void* base, size_t elt) -> int {
                if (tmp.ensure((size_t)rest * elt)) return 1;
                if (cudaMemcpy(tmp.p, (char*)base + (size_t)m * elt, (size_t)rest * elt, cudaMemcpyDeviceToDevice) != cudaSuccess) return fail("slide copy failed");
                if (cudaMemcpy(base, tmp.p, (size_t)rest * elt, cudaMemcpyDeviceToDevice) != cudaSuccess) return fail("slide copy failed");
                return 0;
            };
            if (slide(sp.ex_board, G::S) || slide(sp.ex_pi, sizeof(float) * G::A) || slide(sp.ex_z, sizeof(float) * G::NP) ||
                slide(sp.ex_valid, G::A) || slide(sp.ex_q, sizeof(float) * G::NP)) return 1;
        }
        CK(cudaMemcpy(sp.ex_count, &rest, sizeof(int), cudaMemcpyHostToDevice));
        *out_n = m; return 0;
    }
};

template <class G> static int engine_create_t(const azg_engine_cfg* cfg, azg_net* net, azg_engine** out) {
    EngineT<G>* e = new EngineT<G>();
    if (e->create(cfg, net)) { delete e; return 1; }
    *out = e; return 0;
}
extern "C" int azg_engine_create(const azg_engine_cfg* cfg, azg_net* net, azg_engine** out) {
    if (!cfg || !net || !out) return fail("NULL argument");
    if (require_device()) return 1;
    if (cfg->n_games <= 0 || cfg->numMCTSSims <= 0) return fail("n_games and numMCTSSims must be positive");
    if (net->game_id != cfg->game_id || net->np != cfg->num_players) return fail("net was created for another game");
    DISPATCH(cfg->game_id, cfg->num_players, engine_create_t<G>(cfg, net, out));
}
extern "C" int azg_engine_destroy(azg_engine* e) { delete e; return 0; }
extern "C" int azg_engine_reset(azg_engine* e, int game) { if (!e) return fail("engine is NULL"); return e->reset(game); }
extern "C" int azg_engine_profile(azg_engine* e, int enable) { if (!e) return fail("engine is NULL"); return e->profile(enable); }
extern "C" int azg_engine_kernel_times(azg_engine* e, double* out8) { if (!e || !out8) return fail("NULL argument"); return e->kernel_times(out8); }
extern "C" int azg_engine_search(azg_engine* e, int n, const int8_t* roots, const uint8_t* full_search, const double* noise,
                                 int32_t* out_counts, int32_t* out_raw, float* out_q, void* stream) {
    if (!e) return fail("engine is NULL");
    return e->search(n, roots, full_search, noise, out_counts, out_raw, out_q, (cudaStream_t)stream);
}
extern "C" int azg_engine_stats(azg_engine* e, int64_t* out16) { if (!e || !out16) return fail("NULL argument"); return e->stats(out16); }
extern "C" int azg_engine_selfplay(azg_engine* e, int min_episodes, int max_moves, void* stream) {
    if (!e) return fail("engine is NULL");
    return e->selfplay(min_episodes, max_moves, (cudaStream_t)stream);
}
extern "C" int azg_engine_node(azg_engine* e, int n, const int32_t* slots, const int8_t* boards, int32_t* found, float* es, uint8_t* vs, float* ps,
                               int32_t* ns, double* qsa, int32_t* nsa, int32_t* round, float* qs, void* stream) {
    if (!e) return fail("engine is NULL");
    return e->node(n, slots, boards, found, es, vs, ps, ns, qsa, nsa, round, qs, (cudaStream_t)stream);
}
extern "C" int azg_engine_selfplay_state(azg_engine* e, int8_t* boards, int32_t* players, int32_t* plies, int32_t* active) {
    if (!e) return fail("engine is NULL");
    return e->selfplay_state(boards, players, plies, active);
}
extern "C" int azg_engine_selfplay_inject(azg_engine* e, const azg_selfplay_inject* inj) {
    if (!e) return fail("engine is NULL");
    return e->selfplay_inject(inj);
}
extern "C" int azg_engine_examples_pending(azg_engine* e, int32_t* out_n) {
    if (!e || !out_n) return fail("NULL argument");
    return e->examples_pending(out_n);
}
extern "C" int azg_engine_examples(azg_engine* e, int cap, int8_t* boards, float* pi, float* z, uint8_t* valids, float* q, int32_t* out_n) {
    if (!e || !out_n) return fail("NULL argument");
    return e->examples(cap, boards, pi, z, valids, q, out_n);
}

// tree.cuh -- per-game search trees in HBM and the MCTS kernels (select / expand+backup / finish / gc).
//
// Replaces MCTS.search and its njit helpers (MCTS.py:105-261) for n_games independent trees advanced in
// lock-step, one simulation per game per step, ONE WARP PER GAME:
//   k_select  : root -> leaf walk. Per level: 32 B node header -> coalesced read of the node's legal-edge
//               array (16 B/edge) together with the per-(edge, universe) child links -> PUCT argmax by warp
//               shuffle (first index wins ties, MCTS.py:227-228) -> follow the child link. The reference
//               recomputes every child state on every traversal and finds it again through a dict of board
//               bytes (MCTS.py:125,233-248); here the FIRST traversal of an (edge, universe) does exactly
//               that (parent board from the node's board slot -> make_move + swap_players in shared memory
//               -> 128-bit board hash -> warp-wide probe of the game's open-addressing table, so
//               transpositions resolve to the same node as in the reference) and stores the resulting node
//               index in the link; later traversals are a pure pointer walk.
//   (net)     : batched leaf evaluation of the compacted leaf list (GenericNNetWrapper.predict_server).
//   k_backup  : writes the new node (normalised priors, Q=-42 sentinel, N=0), inserts it in the table,
//               then backs the value up the recorded path, lanes parallel over levels (MCTS.py:176-181).
// Arithmetic follows the reference operation by operation (see oracle/azg_oracle.c header): f64 Q and
// PUCT with explicit round-to-nearest intrinsics so nvcc cannot contract or reassociate them.
//
// HBM layout per game g (arena sizes are engine parameters):
//   nodes [node_cap] NodeHdr 32 B : cpuct*sqrt(Ns), Ns, Qs, edge_off, n_legal, round, kind, best (cached PUCT choice)
//   bestlink[node_cap][U] u32     : child link of the node's best edge per universe (what a non-root visit follows)
//   keys  [node_cap] NodeKey 16 B : 128-bit board hash
//   edges [edge_cap] Edge    16 B : {Q f64, P f32, N i32} for LEGAL actions only, ascending action index
//   acts  [edge_cap] act_t        : action id of each edge
//   child [edge_cap][U] u32       : (next_player << 28) | (child node index + 1), 0 = not resolved yet; U = universes
//   boards[node_cap][SP] i8       : canonical board of every expanded node (source of first-traversal make_move)
//   ht    [ht_cap]   u64          : (tag32 << 32) | (node index + 1), 0 = empty, linear probing
#pragma once
#include "common.cuh"
#include <type_traits>

namespace azg {

struct __align__(16) Edge { double q; float p; int n; };
struct __align__(16) NodeHdr {
    double c1;                                                 // cpuct*sqrt(Ns): PUCT constant, refreshed by the backup
    int ns; float qs;
    uint32_t edge_off; uint16_t n_legal; uint8_t round; uint8_t kind;
    uint16_t best; uint16_t prog; uint32_t rsv1;               // prog: G::progress of the node's state (tree GC); best: edge (index within the node) the next non-root visit takes; refreshed
};                                                             //       by the backup whenever the node's statistics change
struct __align__(16) NodeKey { uint64_t lo, hi; };             // 128-bit board hash of the node (hash-table verification, GC)
static_assert(sizeof(Edge) == 16 && sizeof(NodeHdr) == 32 && sizeof(NodeKey) == 16, "layout");
struct __align__(16) PathEnt { uint32_t node; uint32_t edge_np; uint32_t edge_off; uint32_t n_legal; };   // edge_np: edge index (24 bits) | next_player << 24;
                                                                // edge_off / n_legal of `node` ride along so that the backup can prefetch its edge block at once

enum { NODE_EXPANDED = 0, NODE_TERMINAL = 1 };
enum { LEAF_NONE = 0, LEAF_EXPAND = 1, LEAF_NEW_TERMINAL = 2, LEAF_OLD_TERMINAL = 3 };
enum { ST_SIMS = 0, ST_VISITS, ST_EXPANSIONS, ST_NNEVALS, ST_TERMINAL, ST_OVERFLOW, ST_GC, ST_MAXNODES, ST_SUMLEGAL,
       ST_MOVES, ST_EPISODES, ST_EXAMPLES, ST_GC_SWEEP = 13, ST_SELLEGAL = 15, ST_ROOTLEGAL = 16, ST_REFLEGAL = 17, ST_GC_TRIM = 19, ST_N = 20 };

#ifndef AZG_SEL_PROF
#define AZG_SEL_PROF 0
#endif
constexpr double kNanQ = -42.0;                                // MCTS.py:11
__device__ unsigned long long g_selprof[8];
__device__ unsigned long long g_selprof2[8];                    // debug (AZG_SEL_PROF): cycles per phase of k_select summed over warps
__constant__ long long kMagicSeeds[8] = {31416, 1, 14142, 42, 27183, 2, 16180, 7};   // MCTS.py:14

template <class G>
struct Dev {
    // configuration
    int n_games, node_cap, edge_cap, ht_cap;
    int universes, U, forced_playouts, dirichlet_noise;         // U = max(universes, 1): child links per edge
    int replay;                                                // path replay in k_select (1; AZG_TREE_REPLAY=0 turns it off for A/B tests: results are identical)
    double cpuct, fpu, dir_alpha, temp2;
    uint64_t seed;
    // trees
    NodeHdr* nodes; NodeKey* keys; Edge* edges; typename G::act_t* acts; uint64_t* ht; int* n_nodes; int* n_edges;
    uint32_t* child; int8_t* boards; int* remap; int* gcq;     // remap, gcq: [G][node_cap] scratch of the tree GC
    uint32_t* bestlink;                                        // [G][node_cap][U] child link of every node's cached best edge
    int* ord_cnt; int* ord_list;                               // longest-first work order: [2][32] bucket counts, [2][32][G] games by path depth / 4
    int* root_node;                                            // [G] root node index + 1 once known for this search, else 0
    uint32_t* leaf_link;                                       // [G] child-link slot (index into child, +1) the new leaf hangs on; 0 = root
    // search control (per game)
    int8_t* root;              // [G][SP] canonical root boards
    int* n_sims;               // sims requested for the current search
    uint8_t* full;             // full-search flag (MCTS.py:58)
    const double* noise;       // injected Dirichlet draws or nullptr: game g reads noise[g * noise_gstride + (noise_ply ? (noise_ply[g] - 1) * A : 0) + k]
    size_t noise_gstride;      // elements per game (A for azg_engine_search; P * A for injected self-play)
    const int* noise_ply;      // self-play: current ply of every slot (1-based), selects the ply's row; nullptr for azg_engine_search
    uint64_t game_base;        // global id of slot 0 (azg_engine_cfg.first_game): every RNG stream is keyed by game_base + g, so the games a
                               // slot plays do not depend on how the slots are sharded over ranks
    double* noise_scr;         // [G][A] f64 scratch of the root re-noising in k_select
    unsigned* move_ctr;        // searches done in this slot (RNG counter)
    // ragged self-play (selfplay.cuh k_sp_turn): every slot carries its own simulation index, so a slot whose budget is spent starts
    // its next move at the next launch instead of idling until the slowest budget of the lock-step is done (MCTS.py:58-59 budgets)
    int ragged;                // 1: k_select / k_backup take the simulation index of slot g from sim_idx[g] instead of the launch number
    int* sim_idx;              // [G] simulations done in the current search
    int* turn_list; int* turn_count;   // slots whose search finished in this launch (appended by k_backup, consumed by k_sp_turn)
    // per-simulation scratch
    PathEnt* path; int* path_len; int* leaf_kind; uint64_t* leaf_key; float* leaf_v; uint32_t* leaf_mask; int* leaf_round;
    int8_t* nn_in; float* nn_pi; float* nn_v; int* nn_list; int* nn_count;
    unsigned long long* stats; // [G][ST_N]

    __device__ __forceinline__ NodeHdr* g_nodes(int g) const { return nodes + (size_t)g * node_cap; }
    __device__ __forceinline__ NodeKey* g_keys(int g) const { return keys + (size_t)g * node_cap; }
    __device__ __forceinline__ Edge* g_edges(int g) const { return edges + (size_t)g * edge_cap; }
    __device__ __forceinline__ typename G::act_t* g_acts(int g) const { return acts + (size_t)g * edge_cap; }
    __device__ __forceinline__ uint64_t* g_ht(int g) const { return ht + (size_t)g * ht_cap; }
    __device__ __forceinline__ uint32_t* g_child(int g) const { return child + (size_t)g * edge_cap * U; }
    __device__ __forceinline__ int8_t* g_boards(int g) const { return boards + (size_t)g * node_cap * G::SP; }
    __device__ __forceinline__ uint32_t* g_best(int g) const { return bestlink + (size_t)g * node_cap * U; }
    // the last path walked in universe `uni` of game g (kept per universe: k_select replays its still-valid prefix in parallel)
    __device__ __forceinline__ PathEnt* g_path(int g, int uni, int max_depth) const { return path + ((size_t)g * U + uni) * max_depth; }
    __device__ __forceinline__ int* g_path_len(int g, int uni) const { return path_len + (size_t)g * U + uni; }
};
constexpr uint32_t LINK_IDX = 0x0FFFFFFFu;                      // child link: low 28 bits = node index + 1, high 4 = next player

// ---- board hash (WARP): sum over words of two independent 64-bit mixes of (position, word) ----------
template <class G>
__device__ __forceinline__ void board_hash(const int8_t* b, int lane, uint64_t& lo, uint64_t& hi) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(b);
    uint64_t a = 0, c = 0;
#pragma unroll
    for (int i = lane; i < G::SP / 4; i += 32) {
        uint64_t x = ((uint64_t)(i + 1) << 32) | w[i];
        a += mix64(x ^ 0x2545F4914F6CDD1DULL);
        c += mix64(x * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL);
    }
    lo = warp_sum_u64(a); hi = warp_sum_u64(c);
}

// ---- hash table (WARP): 32 slots probed per round trip ---------------------------------------------------
__device__ __forceinline__ int ht_find(const uint64_t* ht, int cap, const NodeKey* keys, uint64_t klo, uint64_t khi, int lane) {
    const uint32_t tag = (uint32_t)(khi >> 32);
    const uint32_t start = (uint32_t)klo & (uint32_t)(cap - 1);
    for (int probe = 0; probe < cap; probe += 32) {
        uint64_t e = ht[(start + probe + lane) & (uint32_t)(cap - 1)];
        unsigned em = __ballot_sync(FULL, e == 0);
        unsigned mm = __ballot_sync(FULL, e != 0 && (uint32_t)(e >> 32) == tag);
        if (em) mm &= (1u << (__ffs(em) - 1)) - 1u;            // only entries before the first empty slot
        while (mm) {
            int l = __ffs(mm) - 1; mm &= mm - 1;
            int idx = (int)(uint32_t)__shfl_sync(FULL, e, l) - 1;
            const NodeKey k = keys[idx];
            if (k.lo == klo && k.hi == khi) return idx;
        }
        if (em) return -1;
    }
    return -1;
}
__device__ __forceinline__ void ht_insert(uint64_t* ht, int cap, uint64_t klo, uint64_t khi, int idx, int lane) {
    const uint32_t start = (uint32_t)klo & (uint32_t)(cap - 1);
    const uint64_t ent = ((uint64_t)(uint32_t)(khi >> 32) << 32) | (uint32_t)(idx + 1);
    for (int probe = 0; probe < cap; probe += 32) {
        uint32_t slot = (start + probe + lane) & (uint32_t)(cap - 1);
        unsigned em = __ballot_sync(FULL, ht[slot] == 0);
        if (em) { if (lane == __ffs(em) - 1) ht[slot] = ent; __syncwarp(); return; }
    }
}

// ---- float32 sum in the order of the reference's vectorised np.sum (oracle/azg_oracle.c:sum_f32_avx2) -----
__device__ __forceinline__ float warp_sum_avx2order(const float* x, int n, int lane) {
    float s = 0.f; int i = 0;
    if (n >= 32) {
        int nb = n / 32; float acc = 0.f;
        for (int b = 0; b < nb; b++) acc = __fadd_rn(acc, x[32 * b + lane]);
        float p = __fadd_rn(__shfl_down_sync(FULL, acc, 8), acc);
        float t = __fadd_rn(p, __shfl_down_sync(FULL, p, 16));
        float u = __fadd_rn(__shfl_down_sync(FULL, t, 4), t);
        float w = __fadd_rn(u, __shfl_down_sync(FULL, u, 2));
        s = __shfl_sync(FULL, __fadd_rn(w, __shfl_down_sync(FULL, w, 1)), 0);
        i = 32 * nb;
    }
    if ((n & 28) != 0 && (n & ~3) > i) {
        float q = lane == 0 ? s : 0.f;
        for (int k = i; k < (n & ~3); k += 4) if (lane < 4) q = __fadd_rn(q, x[k + lane]);
        float w = __fadd_rn(q, __shfl_down_sync(FULL, q, 2));
        s = __shfl_sync(FULL, __fadd_rn(w, __shfl_down_sync(FULL, w, 1)), 0);
        i = n & ~3;
    }
    for (; i < n; i++) s = __fadd_rn(s, x[i]);
    return s;
}

// ---- root prior noise (WARP): softmax(Ps, T) -> 0.75 P + 0.25 Dir -> (caller normalises) ----------------
// MCTS.py:147-149,156-160,187-197,255-261. `p` is the dense prior (0 at illegal actions) in shared memory,
// `dscr` an A-sized f64 scratch. Injected noise if d.noise, else drawn from the Philox stream.
template <class G>
__device__ void root_noise(const Dev<G>& d, int g, float* p, double* dscr, const uint32_t* mask /*shared, MASK_WORDS*/, int lane) {
    constexpr int A = G::A;
    if (d.temp2 != 1.0) {
        double inv_t = 1.0 / d.temp2;
        for (int a = lane; a < A; a += 32) dscr[a] = pow((double)p[a], inv_t);
        __syncwarp();
        double s = 0;
        if (lane == 0) { for (int a = 0; a < A; a++) s = __dadd_rn(s, dscr[a]); s = __ddiv_rn(1.0, s); }
        s = __shfl_sync(FULL, s, 0);
        for (int a = lane; a < A; a += 32) p[a] = (float)__dmul_rn(dscr[a], s);
        __syncwarp();
    }
    int L = 0;
    for (int k = 0; k < G::MASK_WORDS; k++) L += __popc(mask[k]);
    // dscr[k] <- k-th Dirichlet component
    if (d.noise) {
        const double* nz = d.noise + (size_t)g * d.noise_gstride + (d.noise_ply ? (size_t)(d.noise_ply[g] - 1) * A : 0);
        for (int k = lane; k < L; k += 32) dscr[k] = nz[k];
    } else {
        double alpha = d.dir_alpha > 0 ? d.dir_alpha : 10.0 / (double)L, part = 0;
        for (int k = lane; k < L; k += 32) {
            Philox r(d.seed, ((d.game_base + (uint64_t)g) << 8) | 1u, ((uint64_t)d.move_ctr[g] << 16) | (unsigned)k);
            double x = r.gamma(alpha); dscr[k] = x; part += x;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        for (int k = lane; k < L; k += 32) dscr[k] = dscr[k] / part;
    }
    __syncwarp();
    int before = 0;
    for (int k = 0; k < G::MASK_WORDS; k++) {
        int a = lane + 32 * k;
        if (a < A && (mask[k] >> lane & 1)) {
            int rank = before + __popc(mask[k] & ((1u << lane) - 1u));
            float t1 = __fmul_rn(0.75f, p[a]);
            p[a] = (float)__dadd_rn((double)t1, __dmul_rn(0.25, dscr[rank]));
        }
        before += __popc(mask[k]);
    }
    __syncwarp();
}

// ---- PUCT (WARP): pick_highest_UCB, MCTS.py:210-230; returns the edge index and, in `link`, the child link
// of that edge for universe `uni` (fetched together with the edges so that following it costs no extra round trip).
// c1 = cpuct*sqrt(Ns) and c0 = cpuct*sqrt(Ns+1e-8) come from the node header (the backup refreshes them).
__device__ __forceinline__ uint64_t f64_order_key(double x) {   // unsigned order == double order (no NaNs here)
    const long long b = __double_as_longlong(x);
    return (uint64_t)b ^ ((uint64_t)(b >> 63) | 0x8000000000000000ULL);
}
__device__ __forceinline__ int puct_select(const Edge* e, const uint32_t* child, int U, int uni, int L, double c1, double c0, float qs,
                                           double fpu, bool forced, int n_iter, int lane, uint32_t& link) {
    const double fpu_init = fpu > 0 ? __dsub_rn((double)qs, fpu) : fpu;
    const double kn = __dmul_rn((double)n_iter, 0.5);
    double best = -INFINITY; int best_i = 0x7FFFFFFF, forced_i = 0x7FFFFFFF; uint32_t best_c = 0;
    for (int i = lane; i < L; i += 32) {
        const Edge ed = e[i];
        const uint32_t c = child[(size_t)i * U + uni];
        const double pd = (double)ed.p;
        if (forced) {
            long long th = __double2ll_rz(__dsqrt_rn(__dmul_rn(kn, pd)));
            if ((long long)ed.n < th && i < forced_i) forced_i = i;
        }
        const double ucb = ed.q != kNanQ ? __dadd_rn(ed.q, __ddiv_rn(__dmul_rn(c1, pd), (double)(ed.n + 1)))
                                         : __fma_rn(c0, pd, fpu_init);
        if (ucb > best) { best = ucb; best_i = i; best_c = c; }
    }
    if (forced) {
        forced_i = __reduce_min_sync(FULL, forced_i);
        if (forced_i != 0x7FFFFFFF) { link = child[(size_t)forced_i * U + uni]; return forced_i; }
    }
    // warp argmax with first-index tie-break (MCTS.py:227-228): three REDUX steps on the order-preserving key
    const uint64_t key = f64_order_key(best);
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mh = __reduce_max_sync(FULL, hi);
    const bool ch = hi == mh && best_i != 0x7FFFFFFF;
    const uint32_t ml = __reduce_max_sync(FULL, ch ? lo : 0u);
    const int win = __reduce_min_sync(FULL, (ch && lo == ml) ? best_i : 0x7FFFFFFF);
    link = __shfl_sync(FULL, best_c, win & 31);                  // the winner is the local best of lane (win % 32)
    return win;
}

// L2 prefetch of [p, p + bytes), one 32-byte sector per instruction
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    const char* a = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)31);
    const char* e = reinterpret_cast<const char*>(p) + bytes;
    for (; a < e; a += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
}
// cpuct*sqrt(Ns + 1e-8): pick_highest_UCB's constant for unvisited edges (MCTS.py:226), computed where it is needed
__device__ __forceinline__ double puct_c0(double cpuct, int ns) { return __dmul_rn(cpuct, __dsqrt_rn(__dadd_rn((double)ns, 1e-8))); }
__device__ __forceinline__ uint32_t f32_order_key(float x) { const uint32_t b = __float_as_uint(x); return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u); }

// pick_highest_UCB (MCTS.py:210-230, no forced playouts: every non-root call, MCTS.py:175) for the NEXT visit of a node whose
// statistics just changed. Same exact f64 arithmetic and first-index tie-break as puct_select, but behind an f32 pre-filter with a
// rigorous bound: u32 = f32 evaluation, |u32 - u| <= tol = (|q| + |explore|) * 2^-18 (five f32 roundings + a 2-ulp division are
// < 2^-21 relative); an edge can only be the exact argmax, or tie with it, if u32 + tol >= max_j (u32_j - tol_j). Usually one edge
// survives and no f64 division is executed at all.
__device__ __forceinline__ float puct_bounds(const Edge& ed, float c1f, float c0f, float fif, float& ub) {
    const bool vis = ed.q != kNanQ;
    const float base = vis ? (float)ed.q : fif;
    const float tt = vis ? __fdividef(c1f * ed.p, (float)(ed.n + 1)) : c0f * ed.p;
    const float tol = (fabsf(base) + fabsf(tt)) * 0x1p-18f + 1e-30f;
    ub = base + tt + tol;
    return base + tt - tol;
}
__device__ __forceinline__ double puct_exact(const Edge& ed, double c1, double c0, double fpu_init) {
    const double pd = (double)ed.p;
    return ed.q != kNanQ ? __dadd_rn(ed.q, __ddiv_rn(__dmul_rn(c1, pd), (double)(ed.n + 1))) : __fma_rn(c0, pd, fpu_init);
}
__device__ __forceinline__ int best_edge(const Edge* e, int L, double c1, double cpuct, int ns, float qs, double fpu, int lane) {
    const double fpu_init = fpu > 0 ? __dsub_rn((double)qs, fpu) : fpu;
    const float c1f = (float)c1, c0f = (float)cpuct * sqrtf((float)ns + 1e-8f), fif = (float)fpu_init;
    // up to two edges per lane stay in registers (every Splendor node: L <= 64); longer edge lists are streamed
    const bool h0 = lane < L, h1 = lane + 32 < L;
    Edge e0, e1; e0.q = kNanQ; e0.p = 0.f; e0.n = 0; e1 = e0;
    if (h0) e0 = e[lane];
    if (h1) e1 = e[lane + 32];
    float ub0 = -INFINITY, ub1 = -INFINITY, lb = -INFINITY;
    if (h0) lb = puct_bounds(e0, c1f, c0f, fif, ub0);
    if (h1) lb = fmaxf(lb, puct_bounds(e1, c1f, c0f, fif, ub1));
    for (int i = lane + 64; i < L; i += 32) { float ub; lb = fmaxf(lb, puct_bounds(e[i], c1f, c0f, fif, ub)); }
    const uint32_t mk = __reduce_max_sync(FULL, f32_order_key(lb));
    const unsigned m0 = __ballot_sync(FULL, h0 && f32_order_key(ub0) >= mk), m1 = __ballot_sync(FULL, h1 && f32_order_key(ub1) >= mk);
    if (L <= 64 && __popc(m0) + __popc(m1) == 1) return m0 ? __ffs(m0) - 1 : 31 + __ffs(m1);
    // several candidates (exact ties among equal priors, near ties) or a long edge list: exact evaluation of the candidates only
    const double c0 = puct_c0(cpuct, ns);
    double best = -INFINITY; int best_i = 0x7FFFFFFF;
    if (m0 >> lane & 1) { best = puct_exact(e0, c1, c0, fpu_init); best_i = lane; }
    if (m1 >> lane & 1) { const double u = puct_exact(e1, c1, c0, fpu_init); if (u > best) { best = u; best_i = lane + 32; } }
    for (int i = lane + 64; i < L; i += 32) {
        const Edge ed = e[i]; float ub; puct_bounds(ed, c1f, c0f, fif, ub);
        if (f32_order_key(ub) >= mk) { const double u = puct_exact(ed, c1, c0, fpu_init); if (u > best) { best = u; best_i = i; } }
    }
    const uint64_t key = f64_order_key(best);
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mh = __reduce_max_sync(FULL, best_i != 0x7FFFFFFF ? hi : 0u);
    const bool ch = hi == mh && best_i != 0x7FFFFFFF;
    const uint32_t ml = __reduce_max_sync(FULL, ch ? lo : 0u);
    return __reduce_min_sync(FULL, (ch && lo == ml) ? best_i : 0x7FFFFFFF);
}

// The same choice computed by ONE thread over its own node (the backup refreshes every level of a path at once, one lane per
// level: 12 independent edge streams per warp instead of one). Single pass: the edge with the largest lower bound, and the two
// largest upper bounds; if no other edge's upper bound reaches that lower bound the choice is certain, else the candidates are
// evaluated exactly.
__device__ __forceinline__ int best_edge_lane(const Edge* e, int L, double c1, double cpuct, int ns, float qs, double fpu) {
    const double fpu_init = fpu > 0 ? __dsub_rn((double)qs, fpu) : fpu;
    const float c1f = (float)c1, c0f = (float)cpuct * sqrtf((float)ns + 1e-8f), fif = (float)fpu_init;
    float lbm = -INFINITY, u1 = -INFINITY, u2 = -INFINITY; int im = 0, i1 = -1;
#pragma unroll 4
    for (int i = 0; i < L; i++) {
        const Edge ed = e[i];
        float ub; const float lb = puct_bounds(ed, c1f, c0f, fif, ub);
        if (lb > lbm) { lbm = lb; im = i; }
        if (ub > u1) { u2 = u1; u1 = ub; i1 = i; } else if (ub > u2) u2 = ub;
    }
    if (i1 == im && u2 < lbm) return im;
    const double c0 = puct_c0(cpuct, ns);
    double best = -INFINITY; int bi = im;
    for (int i = 0; i < L; i++) {
        const Edge ed = e[i];
        float ub; puct_bounds(ed, c1f, c0f, fif, ub);
        if (ub >= lbm) { const double u = puct_exact(ed, c1, c0, fpu_init); if (u > best) { best = u; bi = i; } }
    }
    return bi;
}

// Longest-first scheduling. Walk lengths differ by an order of magnitude between games and the kernels end with the slowest
// warp, so work item b of simulation `step` is not game b but the b-th DEEPEST game of the previous simulation (k_backup files
// every game it processed under bucket depth/4; CTAs are dispatched in index order). Simulation 0 uses the identity. Returns -1
// if there is no b-th item (games that have finished their search are no longer listed).
template <class G> struct WorkOrder {                              // per-warp view of the bucket counts: loaded and scanned once per launch
    int c, inc;
    __device__ __forceinline__ void load(const Dev<G>& d, int step, int lane) {
        c = inc = 0;
        if (step == 0) return;
        c = d.ord_cnt[((step - 1) & 1) * 32 + (31 - lane)];      // lane l <-> bucket 31 - l: deepest first
        inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    }
    __device__ __forceinline__ int game(const Dev<G>& d, int b, int step, int lane) const {
        if (step == 0) return b < d.n_games ? b : -1;
        const unsigned m = __ballot_sync(FULL, b < inc);
        if (!m) return -1;
        const int l = __ffs(m) - 1;
        const int exc = __shfl_sync(FULL, inc - c, l);
        return d.ord_list[(size_t)(((step - 1) & 1) * 32 + (31 - l)) * d.n_games + (b - exc)];
    }
};

// k_select runs ONE warp (= one game) per CTA so that an SM slot is recycled as soon as its game's walk ends
// (walk lengths differ a lot between games); 64 registers/thread -> 32 resident CTAs = 32 games per SM.
#ifndef AZG_SELK_WARPS
#define AZG_SELK_WARPS 1
#endif
#ifndef AZG_SEL_MIN_BLOCKS
#define AZG_SEL_MIN_BLOCKS (32 / AZG_SELK_WARPS)
#endif

template <class G> struct WarpSmem {
    __align__(16) int8_t board[G::SP];
    __align__(16) float f[(G::A + 31) / 32 * 32];
    __align__(16) double d[(G::A + 31) / 32 * 32];
    uint32_t mask[G::MASK_WORDS];                                // legal-action bitmask of the state being expanded / re-noised
    uint64_t key[2]; int np;                                     // key / next player of the state in `board` (select kernel)
};
// k_select only needs the board being materialised, its mask and key: the A-sized prior / Dirichlet scratch of the (once per search)
// root re-noising lives in global memory (nn_pi is idle then, plus d.noise_scr), so the kernel keeps 4 warps per CTA and full
// occupancy even with Abalone's 3402 actions.
template <class G> struct SelSmem {
    __align__(16) int8_t board[G::SP];
    uint32_t mask[G::MASK_WORDS];
    uint64_t key[2]; int np;
};
// Warps per CTA of the one-warp-per-game kernels: 4, or 1 when the per-warp scratch is large (Abalone: 3402 actions).
template <class G> __host__ __device__ constexpr int sel_warps() { return sizeof(WarpSmem<G>) * 4 <= 40 * 1024 ? 4 : 1; }
// Warps per CTA of k_select (AZG_SELK_WARPS where the per-warp scratch allows it)
// k_backup: games with a large action space (Abalone) keep the dense priors in global memory (nn_pi is consumed by this kernel only)
// instead of a 41 KB per-warp staging buffer, so the kernel runs 4 warps per CTA at full occupancy for every game.
template <class G> __host__ __device__ constexpr bool big_actions() { return sizeof(WarpSmem<G>) * 4 > 40 * 1024; }
template <class G> __host__ __device__ constexpr int bak_warps() { return 4; }
template <class G> __host__ __device__ constexpr int selk_warps() { return sizeof(SelSmem<G>) * AZG_SELK_WARPS <= 40 * 1024 ? AZG_SELK_WARPS : 1; }

template <class G> __device__ __forceinline__ void warp_load_board(int8_t* sb, const int8_t* src, int lane) {
    if (lane < G::SP / 16) reinterpret_cast<uint4*>(sb)[lane] = reinterpret_cast<const uint4*>(src)[lane];
    if (G::SP / 16 > 32) for (int i = lane + 32; i < G::SP / 16; i += 32) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(src)[i];
    __syncwarp();
}
template <class G> __device__ __forceinline__ void warp_store_board(int8_t* dst, const int8_t* sb, int lane) {
    for (int i = lane; i < G::SP / 16; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(sb)[i];
}

// Re-noise an already expanded root (MCTS.py:156-160): the stored P of the root's edges gets a fresh Dirichlet mix.
template <class G>
__device__ __noinline__ void renoise_root(const Dev<G>& d, int g, uint32_t edge_off, int n_legal, Edge* edges, const typename G::act_t* acts,
                                          SelSmem<G>& ws, int lane) {
    struct { uint32_t edge_off; int n_legal; } h = {edge_off, n_legal};
    float* pf = d.nn_pi + (size_t)g * G::A; uint32_t* m = ws.mask;   // nn_pi[g] is idle until the net runs: dense prior scratch
    for (int k = lane; k < G::MASK_WORDS; k += 32) m[k] = 0;
    for (int a = lane; a < G::A; a += 32) pf[a] = 0.f;
    __syncwarp();
    for (int i = lane; i < h.n_legal; i += 32) { int a = acts[h.edge_off + i]; pf[a] = edges[h.edge_off + i].p; atomicOr(&m[a >> 5], 1u << (a & 31)); }
    __syncwarp();
    root_noise<G>(d, g, pf, d.noise_scr + (size_t)g * G::A, m, lane);
    float s = warp_sum_avx2order(pf, G::A, lane);
    float inv = __fdiv_rn(1.0f, s);
    for (int i = lane; i < h.n_legal; i += 32) { int a = acts[h.edge_off + i]; edges[h.edge_off + i].p = __fmul_rn(pf[a], inv); }
    __syncwarp();
}

// First traversal of an (edge, universe): child state = make_move(parent board) + swap_players (MCTS.py:233-248),
// then the reference's dict lookup (MCTS.py:125-126) as a hash-table probe. The board stays in shared memory, its
// key in ws.key; returns (found node index, or -1) and the next player in ws.np. Kept out of line: once per simulation.
template <class G>
__device__ __noinline__ int materialise_child(const Dev<G>& d, int g, SelSmem<G>& ws, int parent, int action, long long seed, int lane) {
    int8_t* sb = ws.board;
    warp_load_board<G>(sb, d.g_boards(g) + (size_t)parent * G::SP, lane);
    int np = 0;
    if (lane == 0) { np = G::make_move(sb, action, 0, seed, nullptr); ws.np = np; }   // seed != 0 in search: deterministic chance
    __syncwarp();
    np = ws.np;
    if (np != 0) G::swap_players(sb, np, lane);
    uint64_t klo, khi;
    board_hash<G>(sb, lane, klo, khi);
    if (lane == 0) { ws.key[0] = klo; ws.key[1] = khi; }
    __syncwarp();
    return ht_find(d.g_ht(g), d.ht_cap, d.g_keys(g), klo, khi, lane);
}

// Root of this search (MCTS.py:125-126 for the top-level call): board from d.root, key in ws.key; returns its node or -1.
template <class G>
__device__ __noinline__ int locate_root(const Dev<G>& d, int g, SelSmem<G>& ws, int lane) {
    warp_load_board<G>(ws.board, d.root + (size_t)g * G::SP, lane);
    uint64_t klo, khi;
    board_hash<G>(ws.board, lane, klo, khi);
    if (lane == 0) { ws.key[0] = klo; ws.key[1] = khi; }
    __syncwarp();
    const int idx = ht_find(d.g_ht(g), d.ht_cap, d.g_keys(g), klo, khi, lane);
    if (idx >= 0 && lane == 0) d.root_node[g] = idx + 1;
    return idx;
}

// The state in ws.board has never been seen (MCTS.py:130-154): terminal test, else legal mask + hand-over to the net.
template <class G>
__device__ __noinline__ int new_leaf(const Dev<G>& d, int g, SelSmem<G>& ws, uint32_t link_slot, int lane) {
    const int8_t* sb = ws.board;
    float es[G::NP];
    int kind;
    if (G::ended(sb, 0, es, lane)) {                            // MCTS.py:131: getGameEnded(canonicalBoard, 0)
        kind = LEAF_NEW_TERMINAL;
        if (lane == 0) for (int p = 0; p < G::NP; p++) d.leaf_v[(size_t)g * G::NP + p] = es[p];
    } else {
        kind = LEAF_EXPAND;
        int pos = 0;
        if (lane == 0) pos = atomicAdd(d.nn_count, 1);            // one counter for all games: issued first, its round trip hides behind the legal-move test
        G::valid_mask(sb, 0, lane, ws.mask);
        for (int k = lane; k < G::MASK_WORDS; k += 32) d.leaf_mask[(size_t)g * G::MASK_WORDS + k] = ws.mask[k];
        warp_store_board<G>(d.nn_in + (size_t)g * G::SP, sb, lane);
        if (lane == 0) d.nn_list[pos] = g;
    }
    if (lane == 0) { d.leaf_key[2 * (size_t)g] = ws.key[0]; d.leaf_key[2 * (size_t)g + 1] = ws.key[1]; d.leaf_round[g] = (G::round(sb) & 0xFF) | (G::progress(sb) << 8); d.leaf_link[g] = link_slot; }
    return kind;
}

// ============================================================ select ==================================
template <class G>
__device__ __forceinline__ void select_game(const Dev<G>& d, const int g, const int step, SelSmem<G>* sm, const int w, const int lane) {
    if (step >= d.n_sims[g]) { if (lane == 0) d.leaf_kind[g] = LEAF_NONE; return; }
    const bool full = d.full ? d.full[g] != 0 : true;
    const bool forced_root = full && d.forced_playouts;
    const bool noise_now = step == 0 && full && d.dirichlet_noise;
    const int uni = d.universes > 0 ? step % d.universes : 0;
    const NodeHdr* nodes = d.g_nodes(g); Edge* edges = d.g_edges(g); uint32_t* child = d.g_child(g); uint32_t* bestlink = d.g_best(g);
    PathEnt* path = d.g_path(g, uni, G::MAX_DEPTH);
    int depth = 0, kind = LEAF_NONE, sum_legal = 0, root_legal = 0;
    uint32_t link_slot = 0;                                      // child-link slot (+1) a new leaf hangs on; 0 = it is the root
    bool at_new = false;                                         // ws.board holds a state that is not in the tree yet
#if AZG_SEL_PROF == 1
    long long tp0 = clock64(), tp_root = 0, tp_mat = 0, tp_leaf = 0; int n_mat = 0;
#endif
    int idx = d.root_node[g] - 1;
    // ---- path replay, part 1: the previous walk of this universe. Fetch the header and best link of every node it recorded
    //      (levels 1..32 here, one lane each, all loads in flight together) before the dependent walk starts.
    int plen = d.replay ? *d.g_path_len(g, uni) : 0;
    if (plen <= 4) plen = 0;                                     // shallow walks: the parallel fetch costs more than the two or three round trips it saves
#if AZG_SEL_PROF == 1
    if (lane == 0) atomicAdd(&g_selprof2[6], (unsigned long long)plen);
#endif
    PathEnt rp; rp.node = 0; rp.edge_np = 0; rp.edge_off = 0; rp.n_legal = 0;
    NodeHdr rh; rh.kind = NODE_TERMINAL; rh.best = 0; rh.edge_off = 0; rh.n_legal = 0;
    uint32_t rl = 0;
    if (lane + 1 < plen) { rp = path[lane + 1]; rh = nodes[rp.node]; rl = bestlink[(size_t)rp.node * d.U + uni]; }
    if (idx < 0) { idx = locate_root<G>(d, g, sm[w], lane); at_new = idx < 0; }
    while (!at_new) {
        const NodeHdr h = nodes[idx];                            // one round trip per level: 32 B header + the 4 B link of its cached best edge
        uint32_t link = bestlink[(size_t)idx * d.U + uni];
        if (h.kind == NODE_TERMINAL) {                           // MCTS.py:136-138
            kind = LEAF_OLD_TERMINAL;
            if (lane == 0) { const float* es = reinterpret_cast<const float*>(edges + h.edge_off); for (int p = 0; p < G::NP; p++) d.leaf_v[(size_t)g * G::NP + p] = es[p]; }
            break;
        }
        int e = h.best;
        if (depth == 0) {                                        // the root is scanned in full: forced playouts, fresh Dirichlet noise, tree reuse
            if (noise_now) renoise_root<G>(d, g, h.edge_off, (int)h.n_legal, edges, d.g_acts(g), sm[w], lane);
            e = puct_select(edges + h.edge_off, child + (size_t)h.edge_off * d.U, d.U, uni, h.n_legal, h.c1, puct_c0(d.cpuct, h.ns), h.qs, d.fpu,
                            forced_root, step, lane, link);
            root_legal = h.n_legal;
#if AZG_SEL_PROF == 1
            tp_root = clock64() - tp0;
#endif
        }
        sum_legal += h.n_legal;
        const uint32_t eidx = h.edge_off + (uint32_t)e;
        if (link != 0) {                                         // resolved before: pure pointer walk
            if (lane == 0) { PathEnt pe; pe.node = (uint32_t)idx; pe.edge_np = eidx | ((link >> 28) << 24); pe.edge_off = h.edge_off; pe.n_legal = h.n_legal; path[depth] = pe; }
            idx = (int)(link & LINK_IDX) - 1;
            if (++depth >= G::MAX_DEPTH) break;
            if (depth == 1) {
                // ---- path replay, part 2: level L = lane + 1 of the previous path is confirmed if the link INTO it (the root's choice for
                //      lane 0, the fetched best link of level L - 1 otherwise) leads to the recorded node and that node is expanded. The
                //      confirmed prefix is exactly what the dependent walk would have visited (same nodes, same cached choices), found in
                //      two memory round trips instead of one per level. Deep chains (the slowest walks) mostly replay in full.
                uint32_t in_link = link;
                int base = 0;                                    // levels base+1 .. base+32 are in the lanes
                for (;;) {
                    const uint32_t prev = __shfl_up_sync(FULL, rl, 1);
                    const uint32_t inl = lane == 0 ? in_link : prev;
                    const bool ok = base + lane + 1 < plen && (int)(inl & LINK_IDX) - 1 == (int)rp.node && rh.kind == NODE_EXPANDED && base + lane + 1 < G::MAX_DEPTH - 1;
                    const unsigned okm = __ballot_sync(FULL, ok);
                    const int k = okm == FULL ? 32 : __ffs(~okm) - 1;           // confirmed levels base+1 .. base+k
                    if (k == 0) break;
                    // levels whose outgoing edge is confirmed too (all but the last confirmed one) are recorded and counted
                    if (lane < k - 1) {
                        PathEnt pe; pe.node = rp.node; pe.edge_np = (rh.edge_off + rh.best) | ((rl >> 28) << 24); pe.edge_off = rh.edge_off; pe.n_legal = rh.n_legal;
                        path[base + lane + 1] = pe;
                    }
                    sum_legal += warp_sum_i32(lane < k - 1 ? (int)rh.n_legal : 0);
                    idx = (int)__shfl_sync(FULL, rp.node, k - 1); depth = base + k;
#if AZG_SEL_PROF == 1
                    if (lane == 0) atomicAdd(&g_selprof2[5], (unsigned long long)k);
#endif
                    if (k < 32 || base + 33 >= plen) break;
                    // the whole chunk was confirmed and the old path goes on: the last lane's node becomes a recorded level as well
                    // (its outgoing link is the incoming link of the next chunk's first level)
                    in_link = __shfl_sync(FULL, rl, 31);
                    if (lane == 31) { PathEnt pe; pe.node = rp.node; pe.edge_np = (rh.edge_off + rh.best) | ((rl >> 28) << 24); pe.edge_off = rh.edge_off; pe.n_legal = rh.n_legal; path[base + 32] = pe; }
                    const int n31 = (int)__shfl_sync(FULL, (uint32_t)rh.n_legal, 31);
                    base += 32;
                    rp.node = 0; rh.kind = NODE_TERMINAL; rl = 0;
                    if (base + lane + 1 < plen) { rp = path[base + lane + 1]; rh = nodes[rp.node]; rl = bestlink[(size_t)rp.node * d.U + uni]; }
                    // tentatively step onto the first level of the new chunk: if it is not confirmed the walk resumes at the old chunk's last node
                    const uint32_t p0 = __shfl_sync(FULL, rp.node, 0);
                    const int k0 = __shfl_sync(FULL, (int)rh.kind, 0);
                    if (!((int)(in_link & LINK_IDX) - 1 == (int)p0 && k0 == NODE_EXPANDED && base + 1 < plen && base + 1 < G::MAX_DEPTH - 1)) { base -= 32; break; }
                    sum_legal += n31;
                }
            }
            continue;
        }
        const long long seed = d.universes > 0 ? kMagicSeeds[uni] : -1;
#if AZG_SEL_PROF == 1
        long long tm0 = clock64();
#endif
        const int action = d.g_acts(g)[eidx];
        const bool all_uni = d.U > 1 && !G::is_chance_move(action);   // a deterministic move has the same child in every universe: resolve all links at once
        const int found = materialise_child<G>(d, g, sm[w], idx, action, seed, lane);
#if AZG_SEL_PROF == 1
        tp_mat += clock64() - tm0; n_mat++;
#endif
        const uint32_t np = (uint32_t)sm[w].np;
        if (lane == 0) { PathEnt pe; pe.node = (uint32_t)idx; pe.edge_np = eidx | (np << 24); pe.edge_off = h.edge_off; pe.n_legal = h.n_legal; path[depth] = pe; }
        const size_t slot = (size_t)eidx * d.U + uni;
        if (found >= 0) {                                        // transposition, or the same child under another universe
            const uint32_t lv = (uint32_t)(found + 1) | (np << 28);
            if (all_uni) { if (lane < d.U) { child[(size_t)eidx * d.U + lane] = lv; if (depth > 0) bestlink[(size_t)idx * d.U + lane] = lv; } }
            else if (lane == 0) { child[slot] = lv; if (depth > 0) bestlink[(size_t)idx * d.U + uni] = lv; }
            depth++;
            idx = found;
            if (depth >= G::MAX_DEPTH) break;
            continue;
        }
        depth++;
        link_slot = ((uint32_t)slot + 1u) | (all_uni ? 0x80000000u : 0u); at_new = true;   // bit 31: the new node hangs on the edge in every universe
    }
#if AZG_SEL_PROF == 1
    long long tl0 = clock64();
#endif
    if (at_new) kind = new_leaf<G>(d, g, sm[w], link_slot, lane);
#if AZG_SEL_PROF == 1
    tp_leaf = clock64() - tl0;
    if (lane == 0) { const unsigned long long tw = (unsigned long long)(clock64() - tp0); atomicMax(&g_selprof[7], (tw << 20) | ((unsigned long long)depth << 8) | (unsigned long long)n_mat); if (tw > 100000) atomicAdd(&g_selprof2[0], 1ULL); if (tw > 200000) atomicAdd(&g_selprof2[1], 1ULL); if (tw > 50000) atomicAdd(&g_selprof2[2], 1ULL); if (depth > 32) atomicAdd(&g_selprof2[3], 1ULL); if (depth > 64) atomicAdd(&g_selprof2[4], 1ULL);
        atomicAdd(&g_selprof[0], (unsigned long long)(clock64() - tp0)); atomicAdd(&g_selprof[1], (unsigned long long)tp_root); atomicAdd(&g_selprof[2], (unsigned long long)tp_mat);
        atomicAdd(&g_selprof[3], (unsigned long long)tp_leaf); atomicAdd(&g_selprof[4], 1ULL); atomicAdd(&g_selprof[5], (unsigned long long)n_mat); atomicAdd(&g_selprof[6], (unsigned long long)depth); }
#endif
    if (lane == 0) {                                              // counters: reductions without a return value (no load to wait for before the warp can retire)
        *d.g_path_len(g, uni) = depth; d.leaf_kind[g] = kind;
        atomicAdd(&d.stats[(size_t)g * ST_N + ST_SELLEGAL], (unsigned long long)sum_legal); atomicAdd(&d.stats[(size_t)g * ST_N + ST_ROOTLEGAL], (unsigned long long)root_legal);
    }
}

// One CTA per work item (persistent warps pulling tickets from a global counter were measured: no gain for k_select, a loss for k_backup).
template <class G>
__global__ void __launch_bounds__(selk_warps<G>() * 32, selk_warps<G>() == 1 ? 32 : AZG_SEL_MIN_BLOCKS) k_select(const __grid_constant__ Dev<G> d, int step) {
    __shared__ SelSmem<G> sm[selk_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < 32) d.ord_cnt[(step & 1) * 32 + threadIdx.x] = 0;     // the set this simulation's backup fills
    if (d.ragged && blockIdx.x == 0 && threadIdx.x == 0) *d.turn_count = 0;                       // consumed by k_sp_turn before this launch
    WorkOrder<G> wo; wo.load(d, step, lane);
    const int g = wo.game(d, blockIdx.x * selk_warps<G>() + w, step, lane);   // CTAs are dispatched in index order: deepest games first
    if (g >= 0) select_game<G>(d, g, d.ragged ? d.sim_idx[g] : step, sm, w, lane);
}

// Hang a new node on the child-link slot recorded by k_select (slot + 1; bit 31 = deterministic move: all U universes of the edge).
__device__ __forceinline__ void link_new_node(uint32_t* child, uint32_t ls, int U, uint32_t value) {
    const uint32_t slot = (ls & 0x7FFFFFFFu) - 1u;
    if (ls >> 31) { const uint32_t base = slot - slot % (uint32_t)U; for (int u = 0; u < U; u++) child[base + u] = value; }
    else child[slot] = value;
}

// ============================================================ expand + backup =========================
template <class G, class SM>
__device__ __forceinline__ void backup_game(const Dev<G>& d, const int g, const int step, const int launch, SM* sm, const int w, const int lane) {
#if AZG_SEL_PROF == 2
    const long long bp0 = clock64(); long long bp1 = 0, bp2 = 0;
#endif
    const int kind = d.leaf_kind[g];
    if (kind == LEAF_NONE) return;
    constexpr int NP = G::NP, A = G::A, MW = G::MASK_WORDS;
    const int uni_b = d.universes > 0 ? step % d.universes : 0;
    const int depth = *d.g_path_len(g, uni_b);
    if (lane == 0) {                                             // file this game for the next simulation's work order
        const int bk = (launch & 1) * 32 + min(depth >> 2, 31);
        d.ord_list[(size_t)bk * d.n_games + atomicAdd(&d.ord_cnt[bk], 1)] = g;
    }
    NodeHdr* nodes = d.g_nodes(g); Edge* edges = d.g_edges(g); typename G::act_t* acts = d.g_acts(g);
    unsigned long long* st = d.stats + (size_t)g * ST_N;
    if (lane >= 1 && lane < depth) {                             // levels 1..31 of the path: pull the edge blocks and links that the refresh
        const PathEnt pp = d.g_path(g, uni_b, G::MAX_DEPTH)[lane];    // below scans towards L2 now, while the expansion runs
        if (pp.n_legal) {
            l2_prefetch(edges + pp.edge_off, pp.n_legal * 16u);
        }
    }
    float v[NP];
    if (kind == LEAF_EXPAND) {
        float* pf; double* dscr;
        if constexpr (big_actions<G>()) { pf = d.nn_pi + (size_t)g * A; dscr = d.noise_scr + (size_t)g * A; }
        else { pf = sm[w].f; dscr = sm[w].d; for (int a = lane; a < A; a += 32) pf[a] = d.nn_pi[(size_t)g * A + a]; }
        uint32_t* m = sm[w].mask; int L = 0;
        for (int k = lane; k < MW; k += 32) { const uint32_t x = d.leaf_mask[(size_t)g * MW + k]; m[k] = x; L += __popc(x); }
        L = warp_sum_i32(L);
#pragma unroll
        for (int p = 0; p < NP; p++) v[p] = d.nn_v[(size_t)g * NP + p];
        __syncwarp();
        const bool full = d.full ? d.full[g] != 0 : true;
        if (depth == 0 && step == 0 && full && d.dirichlet_noise) root_noise<G>(d, g, pf, dscr, m, lane);   // MCTS.py:147-149
        const float s = warp_sum_avx2order(pf, A, lane);                                                    // normalise, MCTS.py:150
        const float inv = __fdiv_rn(1.0f, s);
        const int ni = d.n_nodes[g], eo = d.n_edges[g];
        if (ni >= d.node_cap || eo + L > d.edge_cap) {
            if (lane == 0) atomicAdd(&st[ST_OVERFLOW], 1ULL);    // arena full: value is still backed up, node not stored
        } else {
            int before = 0;
            for (int k = 0; k < MW; k++) {
                int a = lane + 32 * k;
                if (a < A && (m[k] >> lane & 1)) {
                    int rank = before + __popc(m[k] & ((1u << lane) - 1u));
                    Edge ed; ed.q = kNanQ; ed.p = __fmul_rn(pf[a], inv); ed.n = 0;
                    edges[eo + rank] = ed; acts[eo + rank] = (typename G::act_t)a;
                }
                before += __popc(m[k]);
            }
            uint32_t* child = d.g_child(g);
            for (int k = lane; k < L * d.U; k += 32) child[(size_t)eo * d.U + k] = 0;                          // links unresolved
            {                                                        // board slot of the node (source of its children's states)
                const uint4* src = reinterpret_cast<const uint4*>(d.nn_in + (size_t)g * G::SP);
                uint4* dst = reinterpret_cast<uint4*>(d.g_boards(g) + (size_t)ni * G::SP);
                for (int i = lane; i < G::SP / 16; i += 32) dst[i] = src[i];
            }
            const uint64_t klo = d.leaf_key[2 * (size_t)g], khi = d.leaf_key[2 * (size_t)g + 1];
            const uint32_t ls = d.leaf_link[g];
            __syncwarp();                                        // the new edges are visible to the whole warp
            const int nb = best_edge(edges + eo, L, 0.0, d.cpuct, 0, v[0], d.fpu, lane);      // first visit's choice (all edges unvisited)
            if (lane < d.U) d.g_best(g)[(size_t)ni * d.U + lane] = 0;
            if (lane == 0) {
                if (ls) link_new_node(child, ls, d.U, (uint32_t)(ni + 1) | ((d.g_path(g, uni_b, G::MAX_DEPTH)[depth - 1].edge_np >> 24) << 28));
                else d.root_node[g] = ni + 1;
                NodeKey nk; nk.lo = klo; nk.hi = khi; d.g_keys(g)[ni] = nk;
                NodeHdr h; h.c1 = 0.0; h.ns = 0; h.qs = v[0]; h.edge_off = (uint32_t)eo; h.n_legal = (uint16_t)L;
                h.round = (uint8_t)d.leaf_round[g]; h.prog = (uint16_t)(d.leaf_round[g] >> 8); h.kind = NODE_EXPANDED; h.best = (uint16_t)nb; h.rsv1 = 0;
                nodes[ni] = h; d.n_nodes[g] = ni + 1; d.n_edges[g] = eo + L;
                atomicAdd(&st[ST_EXPANSIONS], 1ULL); atomicAdd(&st[ST_NNEVALS], 1ULL); atomicAdd(&st[ST_SUMLEGAL], (unsigned long long)L);
                atomicMax(&st[ST_MAXNODES], (unsigned long long)(ni + 1));
            }
            ht_insert(d.g_ht(g), d.ht_cap, klo, khi, ni, lane);
        }
    } else {
#pragma unroll
        for (int p = 0; p < NP; p++) v[p] = d.leaf_v[(size_t)g * NP + p];
        if (kind == LEAF_NEW_TERMINAL) {                         // MCTS.py:130-135: terminal states are stored too
            const int ni = d.n_nodes[g], eo = d.n_edges[g];
            if (ni >= d.node_cap || eo + 1 > d.edge_cap) { if (lane == 0) atomicAdd(&st[ST_OVERFLOW], 1ULL); }
            else {
                const uint64_t klo = d.leaf_key[2 * (size_t)g], khi = d.leaf_key[2 * (size_t)g + 1];
                const uint32_t ls = d.leaf_link[g];
                if (lane == 0) {
                    if (ls) link_new_node(d.g_child(g), ls, d.U, (uint32_t)(ni + 1) | ((d.g_path(g, uni_b, G::MAX_DEPTH)[depth - 1].edge_np >> 24) << 28));
                    else d.root_node[g] = ni + 1;
                    float* es = reinterpret_cast<float*>(edges + eo);
                    for (int p = 0; p < 4; p++) es[p] = p < NP ? v[p] : 0.f;
                    NodeKey nk; nk.lo = klo; nk.hi = khi; d.g_keys(g)[ni] = nk;
                    NodeHdr h; h.c1 = 0.0; h.ns = 0; h.qs = 0.f; h.edge_off = (uint32_t)eo; h.n_legal = 0;
                    h.round = (uint8_t)d.leaf_round[g]; h.prog = (uint16_t)(d.leaf_round[g] >> 8); h.kind = NODE_TERMINAL; h.best = 0; h.rsv1 = 0;
                    nodes[ni] = h; d.n_nodes[g] = ni + 1; d.n_edges[g] = eo + 1;
                }
                ht_insert(d.g_ht(g), d.ht_cap, klo, khi, ni, lane);
            }
        }
        if (lane == 0) atomicAdd(&st[ST_TERMINAL], 1ULL);
    }
#if AZG_SEL_PROF == 2
    bp1 = clock64();
#endif
    // ---- backup (MCTS.py:176-181), lanes parallel over levels; a path never visits a node twice (the round
    //      counter in the key increases with every move) so the updates are independent.
    const PathEnt* path = d.g_path(g, uni_b, G::MAX_DEPTH);
    uint32_t* child = d.g_child(g); uint32_t* bestlink = d.g_best(g);
    int carry = 0, ref_legal = 0;                                // rotation accumulated from deeper chunks
    for (int base = ((depth - 1) / 32) * 32; base >= 0 && depth > 0; base -= 32) {
        const int lvl = base + lane;
        PathEnt pe; pe.node = 0; pe.edge_np = 0;
        if (lvl < depth) pe = path[lvl];
        int npv = lvl < depth ? (int)(pe.edge_np >> 24) : 0;
        int suf = npv;                                           // inclusive suffix sum over lanes (levels >= lvl in this chunk)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_down_sync(FULL, suf, o); if (lane + o < 32) suf += t; }
        const int rot = (suf + carry) % NP;                      // v_lvl = roll(v_leaf, rot)
        NodeHdr hn; hn.c1 = 0.0; hn.ns = 0; hn.qs = 0.f; hn.edge_off = 0; hn.n_legal = 0; hn.kind = NODE_TERMINAL;
        if (lvl < depth) {
            const float v0 = v[(NP - rot) % NP];
            Edge* ed = edges + (pe.edge_np & 0xFFFFFFu);
            NodeHdr* nh = nodes + pe.node;
            Edge x = *ed; hn = *nh;
            x.q = __ddiv_rn(__dadd_rn(__dmul_rn((double)x.n, x.q), (double)v0), (double)(x.n + 1));
            x.n += 1;
            hn.qs = __fdiv_rn(__fadd_rn(__fmul_rn((float)(hn.ns + 1), hn.qs), v0), (float)(hn.ns + 2));
            hn.ns += 1;
            hn.c1 = __dmul_rn(d.cpuct, __dsqrt_rn((double)hn.ns));                  // pick_highest_UCB's sqrt(Ns) term, MCTS.py:224-226
            *ed = x; *reinterpret_cast<uint4*>(nh) = *reinterpret_cast<const uint4*>(&hn);     // {c1, ns, qs}: first 16 bytes of the header
            if (base > 0) {                                      // deep paths: levels >= 32 were not covered by the prefetch at the top
                if (pe.n_legal) l2_prefetch(edges + pe.edge_off, pe.n_legal * 16u);
            }
        }
        carry = (carry + __shfl_sync(FULL, suf, 0)) % NP;
#if AZG_SEL_PROF == 2
        __syncwarp(); bp2 += clock64() - bp1;
#endif
        // ---- refresh the cached PUCT choice of every updated non-root node (what the next visit will follow). The root is always
        //      scanned in full by k_select and is skipped here. Two forms with the same (exact) result:
        //      * short paths (the common case in the opening: a handful of levels): the WARP takes the nodes one after the other, all
        //        lanes stream one node's edges coalesced (best_edge), one memory round trip per node;
        //      * long paths: one LANE per level, each lane streams the edge list of its own node (best_edge_lane): up to 32 independent
        //        edge streams per warp instead of one per round trip.
        if (depth <= 7 && base == 0) {
            __syncwarp();                                        // the updated edges / headers of all levels are visible to the whole warp
            for (int l = 1; l < depth; l++) {
                const uint32_t nd = __shfl_sync(FULL, pe.node, l), eoff = __shfl_sync(FULL, pe.edge_off, l), nl = __shfl_sync(FULL, pe.n_legal, l);
                const double c1 = __shfl_sync(FULL, hn.c1, l); const int ns = __shfl_sync(FULL, hn.ns, l); const float qs = __shfl_sync(FULL, hn.qs, l);
                const int knd = __shfl_sync(FULL, (int)hn.kind, l);
                if (knd != NODE_EXPANDED || nl == 0) continue;
                const int nb = best_edge(edges + eoff, (int)nl, c1, d.cpuct, ns, qs, d.fpu, lane);
                if (lane == 0) nodes[nd].best = (uint16_t)nb;
                if (lane < d.U) bestlink[(size_t)nd * d.U + lane] = child[(size_t)(eoff + nb) * d.U + lane];
                if (lane == 0) ref_legal += (int)nl;
            }
        } else if (lvl > 0 && lvl < depth && hn.kind == NODE_EXPANDED && pe.n_legal > 0) {
            const int nb = best_edge_lane(edges + pe.edge_off, (int)pe.n_legal, hn.c1, d.cpuct, hn.ns, hn.qs, d.fpu);
            nodes[pe.node].best = (uint16_t)nb;
            for (int u = 0; u < d.U; u++) bestlink[(size_t)pe.node * d.U + u] = child[(size_t)(pe.edge_off + nb) * d.U + u];
            ref_legal += (int)pe.n_legal;
        }
    }
    ref_legal = warp_sum_i32(ref_legal);
#if AZG_SEL_PROF == 2
    if (lane == 0) { const long long e = clock64(); atomicAdd(&g_selprof[7], 1ULL); atomicAdd(&g_selprof[0], (unsigned long long)(e - bp0)); atomicAdd(&g_selprof[1], (unsigned long long)(bp1 - bp0));
        atomicAdd(&g_selprof[2], (unsigned long long)bp2); atomicAdd(&g_selprof[3], (unsigned long long)(e - bp1 - bp2)); }
#endif
    if (lane == 0) {
        atomicAdd(&st[ST_REFLEGAL], (unsigned long long)ref_legal); atomicAdd(&st[ST_SIMS], 1ULL); atomicAdd(&st[ST_VISITS], (unsigned long long)depth);
        if (d.ragged) {                                          // this slot's next simulation; budget spent => its move is made by k_sp_turn
            d.sim_idx[g] = step + 1;
            if (step + 1 >= d.n_sims[g]) d.turn_list[atomicAdd(d.turn_count, 1)] = g;
        }
    }
}

template <class G>
__global__ void __launch_bounds__(bak_warps<G>() * 32, 32 / bak_warps<G>()) k_backup(const __grid_constant__ Dev<G> d, int step) {
    typedef typename std::conditional<big_actions<G>(), SelSmem<G>, WarpSmem<G>>::type SM;
    __shared__ SM sm[bak_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *d.nn_count = 0;   // leaf list consumed by the net; reset for the next step
    WorkOrder<G> wo; wo.load(d, step, lane);
    const int g = wo.game(d, blockIdx.x * bak_warps<G>() + w, step, lane);   // CTAs are dispatched in index order: deepest games first
    if (g >= 0) backup_game<G, SM>(d, g, d.ragged ? d.sim_idx[g] : step, step, sm, w, lane);
}

// ============================================================ finish (getActionProb tail) ==============
// MCTS.py:67-80: root counts, q vector, forced-playout policy-target pruning.
template <class G>
__global__ void __launch_bounds__(sel_warps<G>() * 32) k_finish(Dev<G> d, int n, int* out_counts, int* out_raw, float* out_q) {
    __shared__ WarpSmem<G> sm[sel_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = blockIdx.x * sel_warps<G>() + w;
    if (g >= n) return;
    constexpr int A = G::A, NP = G::NP;
    int8_t* sb = sm[w].board;
    warp_load_board<G>(sb, d.root + (size_t)g * G::SP, lane);     // (boards may be longer than 32 x 16 bytes: Splendor with 4 players)
    uint64_t klo, khi; board_hash<G>(sb, lane, klo, khi);
    const NodeHdr* nodes = d.g_nodes(g); const Edge* edges = d.g_edges(g); const typename G::act_t* acts = d.g_acts(g);
    const int idx = ht_find(d.g_ht(g), d.ht_cap, d.g_keys(g), klo, khi, lane);
    int* cnt = reinterpret_cast<int*>(sm[w].f);
    for (int a = lane; a < A; a += 32) { cnt[a] = 0; if (out_raw) out_raw[(size_t)g * A + a] = 0; }
    __syncwarp();
    if (idx < 0 || nodes[idx].kind == NODE_TERMINAL) {            // no search ran / terminal root: all-zero counts
        for (int a = lane; a < A; a += 32) out_counts[(size_t)g * A + a] = 0;
        if (out_q && lane < NP) out_q[(size_t)g * NP + lane] = 0.f;
        if (lane == 0) d.move_ctr[g]++;
        return;
    }
    const NodeHdr h = nodes[idx];
    const bool full = d.full ? d.full[g] != 0 : true;
    const bool forced = full && d.forced_playouts;
    const int nsims = d.n_sims[g];
    int best = 0;
    for (int i = lane; i < h.n_legal; i += 32) best = max(best, edges[h.edge_off + i].n);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
    for (int i = lane; i < h.n_legal; i += 32) {
        const Edge ed = edges[h.edge_off + i]; const int a = acts[h.edge_off + i];
        int c = ed.n;
        if (out_raw) out_raw[(size_t)g * A + a] = c;
        if (forced) {
            if (c != best) { float t = __fmul_rn(__fmul_rn(0.5f, ed.p), (float)nsims); c -= (int)__double2ll_rz(__dsqrt_rn((double)t)); }
            c = c > 1 ? c : 0;
        }
        cnt[a] = c;
    }
    __syncwarp();
    for (int a = lane; a < A; a += 32) out_counts[(size_t)g * A + a] = cnt[a];
    if (out_q && lane < NP) out_q[(size_t)g * NP + lane] = lane == 0 ? h.qs : -h.qs / (float)(NP - 1);
    if (lane == 0) d.move_ctr[g]++;
}

// ============================================================ node read-out (MCTS.nodes_data) ==========
// nodes_data[stringRepresentation(board)] of the reference (MCTS.py:37-39: (Es, Vs, Ps, Ns, Qsa, Nsa, r, Qs)) for query i in the tree
// of slot slots[i] (or i): dense A-wide rows rebuilt from the compact legal-edge arrays. found: 0 = not in the tree, 1 = expanded
// node, 2 = terminal node (only Es / round are meaningful).
template <class G>
__global__ void __launch_bounds__(sel_warps<G>() * 32) k_node_query(Dev<G> d, int n, const int* slots, const int8_t* boards, int* found, float* es,
                                                                     uint8_t* vs, float* ps, int* ns, double* qsa, int* nsa, int* rnd, float* qs) {
    __shared__ __align__(16) int8_t sm[sel_warps<G>()][G::SP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * sel_warps<G>() + w;
    if (i >= n) return;
    const int g = slots ? slots[i] : i;
    constexpr int A = G::A, NP = G::NP;
    int8_t* sb = sm[w];
    for (int k = lane; k < G::SP; k += 32) sb[k] = k < G::S ? boards[(size_t)i * G::S + k] : (int8_t)0;
    __syncwarp();
    for (int a = lane; a < A; a += 32) { const size_t o = (size_t)i * A + a; if (vs) vs[o] = 0; if (ps) ps[o] = 0.f; if (qsa) qsa[o] = kNanQ; if (nsa) nsa[o] = 0; }
    if (lane < NP && es) es[(size_t)i * NP + lane] = 0.f;
    if (lane == 0) { found[i] = 0; if (ns) ns[i] = 0; if (rnd) rnd[i] = 0; if (qs) qs[i] = 0.f; }
    if (g < 0 || g >= d.n_games) return;
    uint64_t klo, khi; board_hash<G>(sb, lane, klo, khi);
    const int idx = ht_find(d.g_ht(g), d.ht_cap, d.g_keys(g), klo, khi, lane);
    if (idx < 0) return;
    __syncwarp();
    const NodeHdr h = d.g_nodes(g)[idx]; const Edge* edges = d.g_edges(g); const typename G::act_t* acts = d.g_acts(g);
    if (lane == 0) { found[i] = h.kind == NODE_TERMINAL ? 2 : 1; if (rnd) rnd[i] = h.round; }
    if (h.kind == NODE_TERMINAL) {
        const float* t = reinterpret_cast<const float*>(edges + h.edge_off);
        if (lane < NP && es) es[(size_t)i * NP + lane] = t[lane];
        return;
    }
    if (lane == 0) { if (ns) ns[i] = h.ns; if (qs) qs[i] = h.qs; }
    for (int k = lane; k < h.n_legal; k += 32) {
        const Edge ed = edges[h.edge_off + k]; const size_t o = (size_t)i * A + acts[h.edge_off + k];
        if (vs) vs[o] = 1; if (ps) ps[o] = ed.p; if (qsa) qsa[o] = ed.q; if (nsa) nsa[o] = ed.n;
    }
}

// ============================================================ tree GC ==================================
// Runs only when the arena could not hold another `need_nodes` / `need_edges`. Two tiers:
//   tier 1 (exact): drop every node whose progress is <= the new root's progress and that is not the root itself.
//          G::progress(board) is part of the key and grows with EVERY move (the round counter for Splendor /
//          Santorini / Abalone; round * 21 + tiles taken off the table this round for Azul, whose round counter
//          only moves once per round), so such a node can never be looked up again. This is the reference's cleaning (MCTS.py:86-91, nodes with round < r-5) made tight; both are
//          semantic no-ops.
//   tier 2 (memory pressure only, counted in ST_GC_SWEEP): if tier 1 would not free enough, keep only what is
//          reachable from the new root through resolved child links (breadth-first mark). The reference has
//          unbounded memory and would keep the sub-trees of the moves that were not played; a later transposition
//          into one of them finds fresh statistics here instead of the old ones. Parity tests size the arena so
//          that tier 2 never runs.
// Compacts nodes + boards + edges + actions + child links in place (ascending, destination <= source), rewrites
// the links through an old->new index map and rebuilds the hash table.
template <class G>
__device__ void gc_game(const Dev<G>& d, const int g, int8_t* sb /* warp-private board scratch, SP bytes */, int need_nodes, int need_edges, int force, const int lane) {
    const int nn = d.n_nodes[g], ne = d.n_edges[g];
    if (!force && nn + need_nodes <= d.node_cap && ne + need_edges <= d.edge_cap) return;
    warp_load_board<G>(sb, d.root + (size_t)g * G::SP, lane);
    uint64_t klo, khi; board_hash<G>(sb, lane, klo, khi);
    const int r = G::progress(sb), U = d.U;
    NodeHdr* nodes = d.g_nodes(g); Edge* edges = d.g_edges(g); typename G::act_t* acts = d.g_acts(g); uint64_t* ht = d.g_ht(g);
    uint32_t* child = d.g_child(g); int8_t* boards = d.g_boards(g); int* remap = d.remap + (size_t)g * d.node_cap;
    // ---- would tier 1 free enough?
    int kn1 = 0, ke1 = 0;
    NodeKey* keys = d.g_keys(g);
    for (int i = lane; i < nn; i += 32) {
        const NodeHdr h = nodes[i]; const NodeKey k = keys[i];
        if ((int)h.prog > r || (k.lo == klo && k.hi == khi)) { kn1++; ke1 += h.kind == NODE_TERMINAL ? 1 : (int)h.n_legal; }
    }
    kn1 = warp_sum_i32(kn1); ke1 = warp_sum_i32(ke1);
    const bool sweep = force != 1 && (force == 2 || kn1 + need_nodes > d.node_cap || ke1 + need_edges > d.edge_cap);
    if (sweep) {                                                 // ---- tier 2: mark what the new root reaches
        int* q = d.gcq + (size_t)g * d.node_cap;
        for (int i = lane; i < nn; i += 32) remap[i] = 0;
        const int root = ht_find(ht, d.ht_cap, keys, klo, khi, lane);
        __syncwarp();
        int head = 0, tail = 0;
        if (root >= 0) { if (lane == 0) { remap[root] = -1; q[0] = root; } tail = 1; }
        __syncwarp();
        while (head < tail) {
            const int cnt = min(32, tail - head);
            int off = 0, len = 0;
            if (lane < cnt) { const NodeHdr h = nodes[q[head + lane]]; off = (int)h.edge_off; len = h.kind == NODE_TERMINAL ? 0 : (int)h.n_legal; }
            for (int j = 0; j < cnt; j++) {
                const int oj = __shfl_sync(FULL, off, j), lj = __shfl_sync(FULL, len, j) * U;
                for (int k = 0; k < lj; k += 32) {
                    int ci = -1;
                    if (k + lane < lj) { const uint32_t c = child[(size_t)oj * U + k + lane]; if (c) ci = (int)(c & LINK_IDX) - 1; }
                    const bool won = ci >= 0 && atomicCAS(&remap[ci], 0, -1) == 0;
                    const unsigned wm = __ballot_sync(FULL, won);
                    if (won) q[tail + __popc(wm & ((1u << lane) - 1u))] = ci;
                    tail += __popc(wm);
                }
            }
            head += cnt;
            __syncwarp();
        }
        // ---- tier 3: even the tree the new root reaches leaves no room for this search (long principal variations accumulate
        // thousands of reused nodes). Keep its first nodes in breadth-first order -- q[] is in that order: the nodes closest to the
        // root -- up to what fits beside the coming search, drop the deeper ones. Their parents' links are cleared below, so a later
        // visit re-expands them like any new state (the edge statistics of the kept nodes stay). Counted in ST_GC_TRIM.
        {
            const int budget_n = max(1, d.node_cap - need_nodes), budget_e = d.edge_cap - need_edges;
            int keep_n = 0, acc_e = 0;
            for (int base = 0; base < tail; base += 32) {
                const int i = base + lane; int len = 0;
                if (i < tail) { const NodeHdr h = nodes[q[i]]; len = h.kind == NODE_TERMINAL ? 1 : (int)h.n_legal; }
                int pre = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
                const bool ok = i < tail && i < budget_n && acc_e + pre <= budget_e;      // both conditions are prefix-monotone
                const int cnt = __popc(__ballot_sync(FULL, ok));
                keep_n = base + cnt; acc_e += __shfl_sync(FULL, pre, 31);
                if (cnt < 32) break;
            }
            keep_n = max(keep_n, min(tail, 1));                  // the root itself always stays
            if (keep_n < tail) {
                for (int i = keep_n + lane; i < tail; i += 32) remap[q[i]] = 0;
                if (lane == 0) d.stats[(size_t)g * ST_N + ST_GC_TRIM]++;
                __syncwarp();
            }
        }
    }
    int wn = 0, we = 0;                                          // write cursors
    for (int base = 0; base < nn; base += 32) {
        const int i = base + lane;
        NodeHdr h; h.kind = 0; h.n_legal = 0; h.edge_off = 0; h.round = 0; h.c1 = 0; h.ns = 0; h.qs = 0; h.best = 0; h.prog = 0; h.rsv1 = 0;
        NodeKey nk; nk.lo = nk.hi = 0;
        bool keep = false;
        if (i < nn) { h = nodes[i]; nk = keys[i]; keep = sweep ? remap[i] == -1 : ((int)h.prog > r || (nk.lo == klo && nk.hi == khi)); }
        const int len = keep ? (h.kind == NODE_TERMINAL ? 1 : (int)h.n_legal) : 0;
        const unsigned km = __ballot_sync(FULL, keep);
        int pre = len;                                           // inclusive prefix sum of edge counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
        const int new_off = we + pre - len, old_off = (int)h.edge_off;
        const int new_idx = wn + __popc(km & ((1u << lane) - 1u));
        __syncwarp();
        if (i < nn) remap[i] = keep ? new_idx + 1 : 0;
        if (keep) { h.edge_off = (uint32_t)new_off; nodes[new_idx] = h; keys[new_idx] = nk; }
        // move board, edge, action and link blocks of this chunk, node by node in ascending order (dest <= src)
        for (unsigned mm = km; mm; mm &= mm - 1) {
            const int l = __ffs(mm) - 1;
            const int so = __shfl_sync(FULL, old_off, l), dn = __shfl_sync(FULL, new_off, l), ln = __shfl_sync(FULL, len, l);
            const int si = base + l, di = __shfl_sync(FULL, new_idx, l);
            if (si != di) {
                uint4 t = make_uint4(0, 0, 0, 0);
                const uint4* src = reinterpret_cast<const uint4*>(boards + (size_t)si * G::SP);
                uint4* dst = reinterpret_cast<uint4*>(boards + (size_t)di * G::SP);
                for (int k = 0; k < G::SP / 16; k += 32) {
                    const bool in = k + lane < G::SP / 16;
                    if (in) t = src[k + lane];
                    __syncwarp();
                    if (in) dst[k + lane] = t;
                    __syncwarp();
                }
            }
            if (so != dn) {
                for (int k = 0; k < ln; k += 32) {
                    Edge tmp; typename G::act_t ta = 0; const bool in = k + lane < ln;
                    if (in) { tmp = edges[so + k + lane]; ta = acts[so + k + lane]; }
                    __syncwarp();
                    if (in) { edges[dn + k + lane] = tmp; acts[dn + k + lane] = ta; }
                    __syncwarp();
                }
                for (int k = 0; k < ln * U; k += 32) {
                    uint32_t c = 0; const bool in = k + lane < ln * U;
                    if (in) c = child[(size_t)so * U + k + lane];
                    __syncwarp();
                    if (in) child[(size_t)dn * U + k + lane] = c;
                    __syncwarp();
                }
            }
        }
        wn += __popc(km); we += __shfl_sync(FULL, pre, 31);
    }
    __threadfence_block(); __syncwarp();
    // links: old node index -> new node index (children of kept nodes are kept: their round is larger still)
    for (size_t k = lane; k < (size_t)we * U; k += 32) {
        const uint32_t c = child[k];
        if (c) { const int m = remap[(int)(c & LINK_IDX) - 1]; child[k] = m ? ((uint32_t)m | (c & ~LINK_IDX)) : 0u; }
    }
    {   // cached best-edge links follow the remapped child links
        uint32_t* bestlink = d.g_best(g);
        for (int i = lane; i < wn; i += 32) {
            const NodeHdr h = nodes[i];
            for (int u = 0; u < U; u++) bestlink[(size_t)i * U + u] = h.kind == NODE_TERMINAL ? 0u : child[((size_t)h.edge_off + h.best) * U + u];
        }
    }
    for (int s = lane; s < d.ht_cap; s += 32) ht[s] = 0;
    __threadfence_block(); __syncwarp();
    for (int i = lane; i < wn; i += 32) {                         // lane-parallel re-insert
        const NodeKey k = keys[i];
        const uint64_t ent = ((uint64_t)(uint32_t)(k.hi >> 32) << 32) | (uint32_t)(i + 1);
        uint32_t slot = (uint32_t)k.lo & (uint32_t)(d.ht_cap - 1);
        while (atomicCAS(reinterpret_cast<unsigned long long*>(ht + slot), 0ULL, (unsigned long long)ent) != 0ULL)
            slot = (slot + 1) & (uint32_t)(d.ht_cap - 1);
    }
    if (lane == 0) { d.n_nodes[g] = wn; d.n_edges[g] = we; d.root_node[g] = 0; d.stats[(size_t)g * ST_N + ST_GC]++; if (sweep) d.stats[(size_t)g * ST_N + ST_GC_SWEEP]++; }
}
template <class G>
__global__ void __launch_bounds__(sel_warps<G>() * 32) k_gc(Dev<G> d, int need_nodes, int need_edges, int force) {
    __shared__ WarpSmem<G> sm[sel_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = blockIdx.x * sel_warps<G>() + w;
    if (g >= d.n_games) return;
    gc_game<G>(d, g, sm[w].board, need_nodes, need_edges, force, lane);
}

// Reset trees (MCTS.reset_all_search_trees, MCTS.py:199-203): one slot (game >= 0) or all.
template <class G>
__global__ void k_reset(Dev<G> d, int game) {
    const int g0 = game >= 0 ? game : 0, g1 = game >= 0 ? game + 1 : d.n_games;
    const size_t per = (size_t)d.ht_cap, total = (size_t)(g1 - g0) * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        d.ht[(size_t)g0 * per + i] = 0;
    for (int g = g0 + blockIdx.x * blockDim.x + threadIdx.x; g < g1; g += gridDim.x * blockDim.x) { d.n_nodes[g] = 0; d.n_edges[g] = 0; d.root_node[g] = 0; }
}

}  // namespace azg

// selfplay.cuh -- Coach.executeEpisode / executeEpisodes_batch (Coach.py:37-103) for n_games slots in
// lock-step on the device. One warp per slot.
//   k_sp_begin : (re)start finished slots (Game.getInitBoard + fresh MCTS, Coach.py:94-96), canonicalise the
//                board for the mover (Coach.py:61) and flip the playout-cap coin (MCTS.py:58-59).
//   ... engine_step x n_sims (tree.cuh) ...
//   k_sp_end   : visit-count policy with policy-target pruning (MCTS.py:67-80), example record on full-search
//                plies (Coach.py:65-69), temperature-sampled move (Coach.py:63,266-292), real move with a TRUE
//                random chance draw (random_seed=0, Coach.py:71), end-of-game check and z assignment
//                (Coach.py:73-82), hand-over of the finished game's examples to the output ring.
// Examples are stored un-augmented; symmetries (Coach.py:67) are applied by azg_game_symmetries on read-out.
#pragma once
#include "common.cuh"
#include "tree.cuh"

namespace azg {

template <class G>
struct SelfPlay {
    // per slot
    int8_t* board = nullptr;        // [n][SP] absolute-frame board
    int* player = nullptr;          // mover
    int* ply = nullptr;             // episodeStep (Coach.py:57-60)
    int* active = nullptr;          // 0 => slot needs a new game
    unsigned* games_started = nullptr;
    // per-slot staging of the running game's examples
    int8_t* st_board = nullptr; float* st_pi = nullptr; uint32_t* st_mask = nullptr; float* st_q = nullptr; uint8_t* st_player = nullptr; int* st_count = nullptr;
    // output ring (finished games only)
    int8_t* ex_board = nullptr; float* ex_pi = nullptr; float* ex_z = nullptr; uint8_t* ex_valid = nullptr; float* ex_q = nullptr;
    int* ex_count = nullptr; int ex_cap = 0;
    unsigned long long* counters = nullptr;   // [0] episodes finished [1] examples emitted [2] examples dropped (ring full) [3] moves
    // schedule
    double prob_full = 1.0, t_begin = 1.0, t_end = 0.1, half_life = 10.0;
    int max_ply = 0;
    // injected randomness (azg_engine_selfplay_inject; parity tests replay an episode recorded from the reference): per slot and ply
    // p = episodeStep - 1 < inj_P. With injection every slot plays exactly ONE game from inj_init and then idles.
    const int8_t* inj_init = nullptr;       // [n][S] initial boards
    const double* inj_u_full = nullptr;     // [n][P] playout-cap coin (MCTS.py:58)
    const double* inj_u_move = nullptr;     // [n][P] uniform of random_pick's np.random.choice (Coach.py:289-292)
    const long long* inj_seed = nullptr;    // [n][P] random_seed of the real move (non-zero: deterministic chance draw)
    int inj_P = 0;
};

// visit-count policy at the root of slot g (shared by k_finish-style read-out and self-play)
template <class G>
__device__ bool root_policy(const Dev<G>& d, int g, const int8_t* sb, int lane, int* cnt, uint32_t* mask /*shared, MASK_WORDS*/, float& qs) {
    constexpr int A = G::A;
    uint64_t klo, khi; board_hash<G>(sb, lane, klo, khi);
    const NodeHdr* nodes = d.g_nodes(g); const Edge* edges = d.g_edges(g); const typename G::act_t* acts = d.g_acts(g);
    const int idx = ht_find(d.g_ht(g), d.ht_cap, d.g_keys(g), klo, khi, lane);
    for (int a = lane; a < A; a += 32) cnt[a] = 0;
    for (int k = lane; k < G::MASK_WORDS; k += 32) mask[k] = 0;
    __syncwarp();
    if (idx < 0 || nodes[idx].kind == NODE_TERMINAL) return false;
    const NodeHdr h = nodes[idx];
    const bool forced = d.full[g] && d.forced_playouts;
    const int nsims = d.n_sims[g];
    int best = 0;
    for (int i = lane; i < h.n_legal; i += 32) best = max(best, edges[h.edge_off + i].n);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
    for (int i = lane; i < h.n_legal; i += 32) {
        const Edge ed = edges[h.edge_off + i]; const int a = acts[h.edge_off + i];
        int c = ed.n;
        if (forced) {
            if (c != best) { float t = __fmul_rn(__fmul_rn(0.5f, ed.p), (float)nsims); c -= (int)__double2ll_rz(__dsqrt_rn((double)t)); }
            c = c > 1 ? c : 0;
        }
        cnt[a] = c;
        atomicOr(&mask[a >> 5], 1u << (a & 31));
    }
    qs = h.qs;
    __syncwarp();
    return true;
}

// Start of a move for slot g (WARP): new game if the slot is free (Coach.py:94-96), canonical root (Coach.py:61), playout-cap coin
// (MCTS.py:58-59). `sb` = warp-private board scratch.
template <class G>
__device__ void sp_begin_game(const Dev<G>& d, const SelfPlay<G>& sp, const int g, int8_t* sb, const int lane, int sims_full, int sims_fast) {
    int8_t* board = sp.board + (size_t)g * G::SP;
    const uint64_t gid = d.game_base + (uint64_t)g;                // global slot id: keys every RNG stream
    if (!sp.active[g]) {                                          // new game in this slot: fresh board, fresh tree
        if (sp.inj_P) {                                           // injected episode: one game per slot, then the slot idles
            if (sp.games_started[g] != 0) { if (lane == 0) { d.n_sims[g] = 0; d.full[g] = 0; } return; }
            for (int i = lane; i < G::SP; i += 32) sb[i] = i < G::S ? sp.inj_init[(size_t)g * G::S + i] : (int8_t)0;
        } else if (lane == 0) { Philox rng(d.seed, (gid << 8) | 2u, (uint64_t)sp.games_started[g]); G::init_game(sb, &rng); }
        __syncwarp();
        for (int i = lane; i < G::SP; i += 32) board[i] = sb[i];
        uint64_t* ht = d.g_ht(g);
        for (int s = lane; s < d.ht_cap; s += 32) ht[s] = 0;
        if (lane == 0) { d.n_nodes[g] = 0; d.n_edges[g] = 0; sp.player[g] = 0; sp.ply[g] = 0; sp.active[g] = 1; sp.games_started[g]++; sp.st_count[g] = 0; }
        __syncwarp();
    } else {
        for (int i = lane; i < G::SP; i += 32) sb[i] = board[i];
        __syncwarp();
    }
    const int player = sp.player[g];
    if (player != 0) G::swap_players(sb, player, lane);           // getCanonicalForm, Coach.py:61
    for (int i = lane; i < G::SP; i += 32) d.root[(size_t)g * G::SP + i] = sb[i];
    if (lane == 0) {
        const int ply = sp.ply[g] + 1; sp.ply[g] = ply;
        Philox rng(d.seed, (gid << 8) | 3u, ((uint64_t)sp.games_started[g] << 16) | (unsigned)ply);
        const double coin = sp.inj_P ? sp.inj_u_full[(size_t)g * sp.inj_P + min(ply, sp.inj_P) - 1] : rng.uniform();
        const bool full = coin < sp.prob_full;                    // MCTS.py:58
        d.full[g] = full; d.n_sims[g] = full ? sims_full : sims_fast; d.root_node[g] = 0;
        d.move_ctr[g] = (sp.games_started[g] << 8) | (unsigned)ply;   // a fresh Dirichlet draw for every search (MCTS.py:187-197): counter of root_noise's stream
    }
}
template <class G>
__global__ void __launch_bounds__(sel_warps<G>() * 32) k_sp_begin(Dev<G> d, SelfPlay<G> sp, int sims_full, int sims_fast) {
    __shared__ WarpSmem<G> sm[sel_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = blockIdx.x * sel_warps<G>() + w;
    if (blockIdx.x == 0 && threadIdx.x == 0) *d.nn_count = 0;
    if (g >= d.n_games) return;
    sp_begin_game<G>(d, sp, g, sm[w].board, lane, sims_full, sims_fast);
    if (d.ragged && lane == 0) d.sim_idx[g] = 0;
}

// End of a move for slot g (WARP): visit-count policy, example record, sampled move, real move, end-of-game hand-over (see the header).
template <class G>
__device__ void sp_end_game(const Dev<G>& d, const SelfPlay<G>& sp, const int g, WarpSmem<G>& ws, const int lane) {
    if (!sp.active[g]) return;                                    // idle slot (injected episodes: its one game is over)
    constexpr int A = G::A, NP = G::NP, MW = G::MASK_WORDS;
    const uint64_t gid = d.game_base + (uint64_t)g;
    int8_t* sb = ws.board;
    for (int i = lane; i < G::SP; i += 32) sb[i] = d.root[(size_t)g * G::SP + i];
    __syncwarp();
    int* cnt = reinterpret_cast<int*>(ws.f); double* pw = ws.d;
    uint32_t* mask = ws.mask; float qs = 0.f;
    const bool found = root_policy<G>(d, g, sb, lane, cnt, mask, qs);
    const int ply = sp.ply[g], player = sp.player[g];
    long long total = 0;
    for (int a = lane; a < A; a += 32) total += cnt[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
    Philox rng(d.seed, (gid << 8) | 4u, ((uint64_t)sp.games_started[g] << 16) | (unsigned)ply);
    const size_t inj_i = sp.inj_P ? (size_t)g * sp.inj_P + min(ply, sp.inj_P) - 1 : 0;
    int action = -1;
    if (found && total > 0) {
        const bool full = d.full[g] != 0;
        const int sc = sp.st_count[g];
        if (full && sc < sp.max_ply) {                            // record (canonicalBoard, pi, player, valids, q), Coach.py:65-69
            const size_t o = (size_t)g * sp.max_ply + sc;
            for (int i = lane; i < G::S; i += 32) sp.st_board[o * G::S + i] = sb[i];
            for (int a = lane; a < A; a += 32) sp.st_pi[o * A + a] = (float)((double)cnt[a] / (double)total);
            for (int k = lane; k < MW; k += 32) sp.st_mask[o * MW + k] = mask[k];
            if (lane < NP) sp.st_q[o * NP + lane] = lane == 0 ? qs : -qs / (float)(NP - 1);
            if (lane == 0) { sp.st_player[o] = (uint8_t)player; sp.st_count[g] = sc + 1; }
        }
        // temperature schedule, Coach.py:266-271, then random_pick, Coach.py:278-292
        double T;
        if (sp.half_life < 0) T = ply > -sp.half_life ? sp.t_end : sp.t_begin;
        else T = sp.t_end + (sp.t_begin - sp.t_end) * pow(0.5, (double)ply / sp.half_life);
        if (T == 0.0) {
            int best = 0;
            for (int a = lane; a < A; a += 32) best = max(best, cnt[a]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
            for (int a = lane; a < A; a += 32) pw[a] = cnt[a] == best ? 1.0 : 0.0;
        } else {
            for (int a = lane; a < A; a += 32) pw[a] = cnt[a] > 0 ? pow((double)cnt[a] / (double)total, 1.0 / T) : 0.0;
        }
        __syncwarp();
        if (lane == 0) {                                          // np.random.choice(p) (Coach.py:289-292): cdf = cumsum(p / sum p) / cdf[-1], searchsorted right
            double s = 0; for (int a = 0; a < A; a++) s += pw[a];
            double acc = 0; for (int a = 0; a < A; a++) { acc += pw[a] / s; pw[a] = acc; }
            const double u = sp.inj_P ? sp.inj_u_move[inj_i] : rng.uniform(), last = pw[A - 1];
            int pick = 0; for (int a = 0; a < A; a++) if (pw[a] / last <= u) pick = a + 1;
            action = min(pick, A - 1);
        }
    } else if (lane == 0) {                                       // arena overflow kept the root out of the tree: uniform legal move
        int nlegal = 0; for (int a = 0; a < A; a++) nlegal += G::action_valid(sb, a, 0);
        int k = (int)(rng.uniformf() * (float)nlegal); if (k >= nlegal) k = nlegal - 1;
        for (int a = 0; a < A; a++) if (G::action_valid(sb, a, 0)) { if (k == 0) { action = a; break; } k--; }
        d.stats[(size_t)g * ST_N + ST_OVERFLOW]++;
    }
    action = __shfl_sync(FULL, action, 0);
    // real move on the absolute board with a true random chance draw (random_seed = 0), Coach.py:71
    int8_t* board = sp.board + (size_t)g * G::SP;
    for (int i = lane; i < G::SP; i += 32) sb[i] = board[i];
    __syncwarp();
    int np = 0;
    if (lane == 0) np = G::make_move(sb, action, player, sp.inj_P ? sp.inj_seed[inj_i] : 0, &rng);
    __syncwarp();
    np = __shfl_sync(FULL, np, 0);
    for (int i = lane; i < G::SP; i += 32) board[i] = sb[i];
    float r[NP];
    const bool over = G::ended(sb, np, r, lane);                  // Coach.py:73: getGameEnded(board, curPlayer)
    if (lane == 0) { sp.player[g] = np; atomicAdd(&sp.counters[3], 1ULL); d.stats[(size_t)g * ST_N + ST_MOVES]++; }
    if (over) {
        const int n_ex = sp.st_count[g];
        int base = 0;
        if (lane == 0) base = atomicAdd(sp.ex_count, n_ex);
        base = __shfl_sync(FULL, base, 0);
        for (int e = 0; e < n_ex; e++) {
            const size_t o = (size_t)g * sp.max_ply + e; const int dst = base + e;
            if (dst >= sp.ex_cap) { if (lane == 0) atomicAdd(&sp.counters[2], 1ULL); continue; }
            const int pl = sp.st_player[o];
            for (int i = lane; i < G::S; i += 32) sp.ex_board[(size_t)dst * G::S + i] = sp.st_board[o * G::S + i];
            for (int a = lane; a < A; a += 32) {
                sp.ex_pi[(size_t)dst * A + a] = sp.st_pi[o * A + a];
                sp.ex_valid[(size_t)dst * A + a] = (sp.st_mask[o * MW + (a >> 5)] >> (a & 31)) & 1;
            }
            if (lane < NP) {
                sp.ex_z[(size_t)dst * NP + lane] = r[(lane + pl) % NP];          // np.roll(r, -player), Coach.py:79
                sp.ex_q[(size_t)dst * NP + lane] = sp.st_q[o * NP + lane];
            }
        }
        if (lane == 0) {
            sp.active[g] = 0; atomicAdd(&sp.counters[0], 1ULL); atomicAdd(&sp.counters[1], (unsigned long long)n_ex);
            d.stats[(size_t)g * ST_N + ST_EPISODES]++; d.stats[(size_t)g * ST_N + ST_EXAMPLES] += (unsigned)n_ex;
        }
    }
}
template <class G>
__global__ void __launch_bounds__(sel_warps<G>() * 32) k_sp_end(Dev<G> d, SelfPlay<G> sp) {
    __shared__ WarpSmem<G> sm[sel_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = blockIdx.x * sel_warps<G>() + w;
    if (g >= d.n_games) return;
    sp_end_game<G>(d, sp, g, sm[w], lane);
}

// Ragged self-play: the slots whose search budget was spent in the launch that just ended (k_backup listed them) make their move
// and start the next one -- end of move, begin of move (new game if the old one is over), tree GC check -- while every other slot
// simply goes on with its next simulation. One launch per lock-step simulation, a few warps of work on average.
template <class G>
__global__ void __launch_bounds__(sel_warps<G>() * 32) k_sp_turn(Dev<G> d, SelfPlay<G> sp, int sims_full, int sims_fast) {
    __shared__ WarpSmem<G> sm[sel_warps<G>()];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = *d.turn_count;
    for (int i = blockIdx.x * sel_warps<G>() + w; i < n; i += gridDim.x * sel_warps<G>()) {
        const int g = d.turn_list[i];
        sp_end_game<G>(d, sp, g, sm[w], lane);
        __syncwarp();
        sp_begin_game<G>(d, sp, g, sm[w].board, lane, sims_full, sims_fast);
        __syncwarp();
        const int ns = d.n_sims[g];
        gc_game<G>(d, g, sm[w].board, ns + 2, (ns + 2) * G::MAX_LEGAL, 0, lane);
        __syncwarp();
        if (lane == 0) d.sim_idx[g] = 0;
    }
}

}  // namespace azg

// azul.cuh -- Azul for 2 players (azul/AzulLogicNumba.py) as __device__ code against the game-plugin interface of
// splendor.cuh / santorini.cuh / abalone.cuh.
//
// The CPU oracle's Azul restatement (oracle/azg_oracle.c: azo_azul_*) is pinned bit-exactly against the reference
// (tests/test_oracle_azul.py); this file follows it function by function and passes the same goldens through the C ABI
// (tests/test_gpu_azul.py). The LANE functions also compile for the host (tests/host/azul_plugin_emul.cpp, AZG_HOST_EMUL).
//
// Board = the reference's int8[23][6] state, byte-compatible: row 0 scores (P0, P1, round), 1 bag, 2 discards, 3 centre (+ first-player
// token in column 5), 4-8 factories, 9-10 pattern-line colours (-1 empty; column 5 = holds the token), 11-12 tiles per pattern line
// (column 5 = floor), 13-17 / 18-22 walls (:6-24). Action = 30 source + 6 colour + line; source 0 = centre, line 5 = floor (:27-48).
// WARP functions are called by all 32 lanes, LANE functions by one lane (caller brackets with __syncwarp()).
#pragma once
#ifndef AZG_HOST_EMUL                       // tests/host/azul_plugin_emul.cpp compiles the LANE functions for the host
#include "common.cuh"
#endif

namespace azg {

__constant__ int8_t kAzulFloorPenalty[8] = {0, 1, 2, 4, 6, 8, 11, 14};     // score_round :176 discard_mapping

struct Azul {
    static constexpr int GAME_ID = 4;
    static constexpr int NP = 2;
    static constexpr int D0 = 23, D1 = 6, D2 = 1;
    static constexpr int S = 138;                              // observation_size() = (23, 6)
    static constexpr int SP = 144;                             // padded to 16 B
    static constexpr int A = 180;                              // action_size()
    static constexpr int MASK_WORDS = 6;
    static constexpr int MAX_LEGAL = 128;                      // probe on the reference: mean 24, max 114
    static constexpr int EDGE_FACTOR = 32;
    static constexpr int MAX_MOVES = 127;                      // the round counter is an int8; games recorded from the reference end after 6-9 rounds (<= 70 plies)
    static constexpr int MAX_DEPTH = MAX_MOVES + 4;
    static constexpr int MAX_SYM = 120;                        // all orders of the five factories
    typedef uint8_t act_t;

    static __device__ __forceinline__ int8_t& at(int8_t* b, int r, int c) { return b[6 * r + c]; }
    static __device__ __forceinline__ int at(const int8_t* b, int r, int c) { return b[6 * r + c]; }

    // The round may end on any move (it does when the last tiles leave the table), and a new round draws 20 tiles: every move is
    // treated as a potential chance move, i.e. children are resolved per universe.
    static __device__ __forceinline__ bool is_chance_move(int) { return true; }
    static __device__ __forceinline__ int round(const int8_t* b) { return at(b, 0, 2); }                   // get_round :333-334
    static __device__ __forceinline__ int score(const int8_t* b, int player) { return at(b, 0, player); } // get_score :83-84
    // A counter that grows with EVERY move (tree GC, tree.cuh): the round counter only moves once per round, but every move takes at
    // least one tile off the table (factories + centre), and a new round (<= 20 tiles dealt) outweighs a full round of takes.
    static __device__ __forceinline__ int progress(const int8_t* b) {
        int on_table = 0;
        for (int r = 3; r < 9; r++) for (int c = 0; c < 5; c++) on_table += at(b, r, c);
        return round(b) * 21 + (20 - min(on_table, 20));
    }

    // One action's legality (valid_moves :97-124).
    static __device__ bool action_valid(const int8_t* b, int a, int player) {
        const int src = a / 30, colour = (a % 30) / 6, line = a % 6;
        const bool avail = src == 0 ? at(b, 3, colour) != 0 : at(b, 3 + src, colour) > 0;
        if (!avail) return false;
        if (line == 5) return true;                                                                        // the floor always takes tiles
        const int pc = at(b, 9 + player, line);
        if (pc == -1) return at(b, 13 + 5 * player + line, (colour + line) % 5) == 0;                       // free line: the wall cell of this colour must be empty
        return pc == colour && at(b, 11 + player, line) < line + 1;                                        // started line of the same colour, not full
    }
    // WARP: legal-action bitmask into `w` (MASK_WORDS words of warp-private shared memory), visible to all lanes on return.
    static __device__ __forceinline__ void valid_mask(const int8_t* b, int player, int lane, uint32_t* w) {
#pragma unroll
        for (int k = 0; k < MASK_WORDS; k++) {
            const int a = lane + 32 * k;
            const uint32_t m = __ballot_sync(FULL, a < A && action_valid(b, a, player));
            if (lane == 0) w[k] = m;
        }
        __syncwarp();
    }

    // select_tiles_from_bag :258-269. seed != 0: the reference's deterministic draw; seed == 0 (real moves): Philox.
    static __device__ void draw(int8_t* b, int num, long long seed, Philox* rng, int8_t* result) {
        for (int t = 0; t < num; t++) {
            int total = 0;
            for (int c = 0; c < 6; c++) total += at(b, 1, c);
            if (total <= 0) return;                                                                        // the reference would divide by zero here
            long long tile;
            if (seed == 0) { tile = (long long)(rng->uniformf() * (float)total); if (tile >= total) tile = total - 1; }
            else {
                long long h = 0;
                for (int c = 0; c < 5; c++) h += (long long)at(b, 1, c) << c;
                tile = (4594591LL * (seed + h)) % total; if (tile < 0) tile += total;                      // Python modulo
            }
            int idx = 0, cum = 0;
            for (; idx < 5; idx++) { cum += at(b, 1, idx); if (cum > tile) break; }                         // searchsorted(cumsum, tile, side='right')
            if (idx > 4) idx = 4;
            result[idx]++; at(b, 1, idx)--;
        }
    }
    // setup_new_round :238-256. Returns the next player.
    static __device__ int setup_new_round(int8_t* b, long long seed, Philox* rng) {
        for (int i = 0; i < 5; i++) {
            int total = 0;
            for (int c = 0; c < 6; c++) total += at(b, 1, c);
            int8_t res[6] = {0, 0, 0, 0, 0, 0};
            if (total < 4) {
                for (int c = 0; c < 6; c++) { at(b, 4 + i, c) = (int8_t)at(b, 1, c); at(b, 1, c) = (int8_t)at(b, 2, c); at(b, 2, c) = 0; }
                draw(b, 4 - total, seed, rng, res);
                for (int c = 0; c < 6; c++) at(b, 4 + i, c) = (int8_t)(at(b, 4 + i, c) + res[c]);
            } else {
                draw(b, 4, seed, rng, res);
                for (int c = 0; c < 6; c++) at(b, 4 + i, c) = res[c];
            }
        }
        int next_player;
        if (at(b, 10, 5) == 1) { next_player = 1; at(b, 10, 5) = 0; } else { next_player = 0; at(b, 9, 5) = 0; }
        at(b, 0, 2) = (int8_t)(at(b, 0, 2) + 1);
        at(b, 3, 5) = 1;
        return next_player;
    }
    static __device__ int run_len(const int8_t* w, int r, int c, bool along_row) {                          // count_consecutive_ones :199-210
        int count = 1;
        if (along_row) { for (int k = c - 1; k >= 0 && w[6 * r + k] == 1; k--) count++; for (int k = c + 1; k < 5 && w[6 * r + k] == 1; k++) count++; }
        else { for (int k = r - 1; k >= 0 && w[6 * k + c] == 1; k--) count++; for (int k = r + 1; k < 5 && w[6 * k + c] == 1; k++) count++; }
        return count;
    }
    static __device__ int score_change(int8_t* w, int r, int c) {                                          // score_change :212-220; w = the player's five wall rows
        w[6 * r + c] = 1;
        const bool row_adj = (c > 0 && w[6 * r + c - 1] == 1) || (c < 4 && w[6 * r + c + 1] == 1);
        const bool col_adj = (r > 0 && w[6 * (r - 1) + c] == 1) || (r < 4 && w[6 * (r + 1) + c] == 1);
        if (!row_adj && !col_adj) return 1;
        return (row_adj ? run_len(w, r, c, true) : 0) + (col_adj ? run_len(w, r, c, false) : 0);
    }
    static __device__ void score_round(int8_t* b) {                                                        // score_round :161-181
        int pl[10], rw[10], col[10], n = 0;
        for (int p = 0; p < 2; p++)
            for (int r = 0; r < 5; r++)
                if (at(b, 11 + p, r) == r + 1) { pl[n] = p; rw[n] = r; col[n] = at(b, 9 + p, r); n++; }    // np.where order
        for (int i = 0; i < n; i++) {
            const int c = ((col[i] + rw[i]) % 5 + 5) % 5;
            at(b, 0, pl[i]) = (int8_t)(at(b, 0, pl[i]) + score_change(&at(b, 13 + 5 * pl[i], 0), rw[i], c));
            at(b, 13 + 5 * pl[i] + rw[i], c) = 1;
        }
        for (int i = 0; i < n; i++) { const int c = (col[i] % 6 + 6) % 6; at(b, 2, c) = (int8_t)(at(b, 2, c) + rw[i]); }
        for (int i = 0; i < n; i++) { at(b, 11 + pl[i], rw[i]) = 0; at(b, 9 + pl[i], rw[i]) = -1; }
        for (int p = 0; p < 2; p++) {
            int fl = at(b, 11 + p, 5); fl = fl > 7 ? 7 : (fl < 0 ? 0 : fl);
            const int sc = at(b, 0, p) - kAzulFloorPenalty[fl];
            at(b, 0, p) = (int8_t)(sc > 0 ? sc : 0);
            at(b, 11 + p, 5) = 0;
        }
    }
    static __device__ bool game_over(const int8_t* b) {                                                    // check_game_over :153-159
        for (int r = 0; r < 10; r++) {
            bool all = true;
            for (int c = 0; c < 5; c++) all = all && at(b, 13 + r, c) == 1;
            if (all) return true;
        }
        return false;
    }
    static __device__ void score_bonuses(int8_t* b) {                                                      // score_bonuses :183-197
        for (int p = 0; p < 2; p++) {
            const int8_t* w = b + 6 * (13 + 5 * p); int add = 0;
            for (int r = 0; r < 5; r++) { bool all = true; for (int c = 0; c < 5; c++) all = all && w[6 * r + c] == 1; if (all) add += 2; }
            for (int c = 0; c < 5; c++) { bool all = true; for (int r = 0; r < 5; r++) all = all && w[6 * r + c] == 1; if (all) add += 7; }
            for (int i = 0; i < 5; i++) { bool all = true; for (int j = 0; j < 5; j++) all = all && w[6 * j + (j + i) % 5] == 1; if (all) add += 10; }
            at(b, 0, p) = (int8_t)(at(b, 0, p) + add);
        }
    }
    // LANE: make_move :126-151. Returns the next player.
    static __device__ int make_move(int8_t* b, int move, int player, long long seed, Philox* rng) {
        int8_t* src = move < 30 ? b + 6 * 3 : b + 6 * (4 + (move - 30) / 30);
        const int colour = (move % 30) / 6, line = move % 6, num = src[colour];
        int to_floor;
        if (line == 5) to_floor = num;
        else {
            const int on_line = at(b, 11 + player, line);
            const int to_line = min(line + 1 - on_line, num);
            to_floor = num - to_line;
            at(b, 11 + player, line) = (int8_t)(on_line + to_line);
            at(b, 9 + player, line) = (int8_t)colour;
        }
        at(b, 11 + player, 5) = (int8_t)(at(b, 11 + player, 5) + to_floor);
        at(b, 2, colour) = (int8_t)(at(b, 2, colour) + to_floor);
        src[colour] = 0;
        if (move < 30) {
            if (src[5] == 1) { at(b, 11 + player, 5) = (int8_t)(at(b, 11 + player, 5) + 1); at(b, 9 + player, 5) = 1; src[5] = 0; }
        } else {
            for (int c = 0; c < 6; c++) { at(b, 3, c) = (int8_t)(at(b, 3, c) + src[c]); src[c] = 0; }
        }
        bool empty = true;
        for (int i = 6 * 4; i < 6 * 9; i++) empty = empty && b[i] == 0;
        for (int c = 0; c < 5; c++) empty = empty && at(b, 3, c) == 0;
        if (!empty) return (player + 1) % 2;
        score_round(b);
        const int next_player = setup_new_round(b, seed, rng);
        if (game_over(b)) score_bonuses(b);
        return next_player;
    }
    // check_end_game :283-302 (next_player is not used by this game). Every lane computes the same result.
    static __device__ bool ended(const int8_t* b, int next_player, float (&out)[NP], int lane) {
        out[0] = out[1] = 0.f;
        if (!game_over(b)) return false;
        int rows[2] = {0, 0};
        for (int p = 0; p < 2; p++)
            for (int r = 0; r < 5; r++) { bool all = true; for (int c = 0; c < 5; c++) all = all && at(b, 13 + 5 * p + r, c) == 1; rows[p] += all ? 1 : 0; }
        const int s0 = at(b, 0, 0), s1 = at(b, 0, 1);
        if (s0 > s1 || (s0 == s1 && rows[0] > rows[1])) { out[0] = 1.f; out[1] = -1.f; }
        else if (s1 > s0 || (s0 == s1 && rows[1] > rows[0])) { out[0] = -1.f; out[1] = 1.f; }
        else { out[0] = 0.01f; out[1] = 0.01f; }
        return true;
    }
    // WARP: swap_players :304-309 (the reference swaps whenever it is called, i.e. for next_player == 1).
    static __device__ void swap_players(int8_t* b, int nb_swaps, int lane) {
        if ((nb_swaps & 1) == 0) return;
        if (lane < 6) {
            int8_t t = b[6 * 9 + lane]; b[6 * 9 + lane] = b[6 * 10 + lane]; b[6 * 10 + lane] = t;
            t = b[6 * 11 + lane]; b[6 * 11 + lane] = b[6 * 12 + lane]; b[6 * 12 + lane] = t;
        }
        if (lane < 30) { const int8_t t = b[6 * 13 + lane]; b[6 * 13 + lane] = b[6 * 18 + lane]; b[6 * 18 + lane] = t; }
        if (lane == 31) { const int8_t t = b[0]; b[0] = b[1]; b[1] = t; }
        __syncwarp();
    }
    // LANE: init_game :86-92 (the first round's draws come from `rng`).
    static __device__ void init_game(int8_t* b, Philox* rng) {
        for (int i = 0; i < SP; i++) b[i] = 0;
        for (int c = 0; c < 5; c++) { at(b, 1, c) = 20; at(b, 9, c) = -1; at(b, 10, c) = -1; }
        (void)setup_new_round(b, 0, rng);
    }

    // get_symmetries :311-331: factory i of symmetry k is factory perm_k[i] of the board, perm_k = the k-th permutation of (0..4) in
    // lexicographic order (AzulLogic.py:4-126); policy / valids blocks of 30 move with their factory, the centre block stays.
    static __device__ int num_symmetries(const int8_t* b) { return MAX_SYM; }
    static __device__ __forceinline__ void kth_perm(int k, int (&perm)[5]) {                                // factorial number system
        int pool[5] = {0, 1, 2, 3, 4}, fact = 24;
        for (int i = 0; i < 5; i++) {
            const int q = k / fact; k -= q * fact;
            perm[i] = pool[q];
            for (int j = q; j < 4 - i; j++) pool[j] = pool[j + 1];
            if (i < 4) fact /= 4 - i;
        }
    }
    static __device__ void symmetry(const int8_t* b, const float* pi, const uint8_t* mask, int k, int lane,
                                    int8_t* ob, float* opi, uint8_t* om) {
        int perm[5]; kth_perm(k, perm);
        for (int i = lane; i < S; i += 32) {
            const int r = i / 6, c = i - 6 * r;
            ob[i] = (r >= 4 && r < 9) ? b[6 * (4 + perm[r - 4]) + c] : b[i];
        }
        for (int a = lane; a < A; a += 32) {
            const int blk = a / 30, off = a - 30 * blk;
            const int src = blk == 0 ? a : 30 * (perm[blk - 1] + 1) + off;
            opi[a] = pi[src]; om[a] = mask[src];
        }
    }
};

}  // namespace azg

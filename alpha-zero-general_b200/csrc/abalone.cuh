// abalone.cuh -- Abalone (Belgian daisy start, no dynamic komi: the shipped constants INITIAL_LAYOUT = 1,
// ENABLE_DYNAMIC_KOMI = False, abalone/AbaloneLogicNumba.py:5-6) as __device__ code: the Board jitclass of
// abalone/AbaloneLogicNumba.py:150-441 (:254-331 valid_moves, :333-374 make_move, :376-392 check_end_game,
// :394-406 swap_players, :408-441 get_symmetries with the action map of :95-148 computed on the fly).
//
// Board = the reference's int8[9][9][4] axial-hex state, byte-compatible (HWC: cell (r,q) at bytes 4(9r+q)..+3):
//   ch0 mover's marbles, ch1 opponent's marbles, ch2 board mask (4 <= r+q <= 12, 61 cells), ch3 misc:
//   misc[0,0] = byte 3 mover's score, misc[0,1] = byte 7 opponent's score, misc[0,2] = byte 11 round counter.
// Action = 378 r + 42 q + plane (anchor cell = marble with min r then min q); plane 0-5 one marble in direction d,
// 6 + 6 axis + d two marbles, 24 + 6 axis + d three marbles; axis in {E, SE, SW}.
#pragma once
#include "common.cuh"

namespace azg {

__constant__ int8_t kAbDR[6] = {0, 1, 1, 0, -1, -1};            // DIRECTIONS, AbaloneLogicNumba.py:53-60
__constant__ int8_t kAbDQ[6] = {1, 0, -1, -1, 0, 1};
__constant__ uint8_t kAbFlipDir[6] = {3, 2, 1, 0, 5, 4};         // direction under the reflection, :140

struct Abalone {
    static constexpr int GAME_ID = 3;
    static constexpr int NP = 2;
    static constexpr int D0 = 9, D1 = 9, D2 = 4;
    static constexpr int S = 324;
    static constexpr int SP = 336;                             // padded to 16 B
    static constexpr int A = 3402;                             // 9 * 9 * 42
    static constexpr int MASK_WORDS = 107;
    static constexpr int MAX_LEGAL = 160;                      // edge-space reservation per expansion (largest seen in play: 99)
    static constexpr int EDGE_FACTOR = 80;                     // mean legal moves ~61-63
    static constexpr int MAX_MOVES = 127;                      // check_end_game: round >= 127
    static constexpr int MAX_DEPTH = MAX_MOVES + 5;
    static constexpr int MAX_SYM = 12;
    typedef uint16_t act_t;

    static __device__ __forceinline__ int at(const int8_t* b, int r, int q, int ch) { return b[4 * (9 * r + q) + ch]; }
    static __device__ __forceinline__ bool on_board(const int8_t* b, int r, int q) {           // is_on_board :86-90
        return r >= 0 && r < 9 && q >= 0 && q < 9 && at(b, r, q, 2) == 1;
    }
    static __device__ __forceinline__ int round(const int8_t* b) { return b[11]; }              // misc[0,2]
    static __device__ __forceinline__ int progress(const int8_t* b) { return round(b); }           // grows with every move (tree GC, tree.cuh)
    static __device__ __forceinline__ int score(const int8_t* b, int player) { return player == 0 ? b[3] : b[7]; }
    static __device__ __forceinline__ void decode(int a, int& r, int& q, int& size, int& axis, int& d) {    // _decode_action :72-84
        const int plane = a % 42; q = (a / 42) % 9; r = a / 378; d = plane % 6;
        if (plane < 6) { size = 1; axis = 0; } else if (plane < 24) { size = 2; axis = (plane - 6) / 6; } else { size = 3; axis = (plane - 24) / 6; }
    }
    static __device__ __forceinline__ int encode(int r, int q, int size, int axis, int d) {                 // _encode_action :62-70
        const int plane = size == 1 ? d : (size == 2 ? 6 + axis * 6 + d : 24 + axis * 6 + d);
        return r * 378 + q * 42 + plane;
    }

    static __device__ __forceinline__ bool is_chance_move(int) { return false; }     // no chance in this game
    // One action's legality (valid_moves :254-331). `player` selects the marble plane (0 on canonical boards).
    static __device__ bool action_valid(const int8_t* b, int a, int player) {
        int r, q, size, axis, d; decode(a, r, q, size, axis, d);
        const int opp = 1 - player;
        if (at(b, r, q, player) == 0) return false;
        const int dr = kAbDR[d], dq = kAbDQ[d], ar = kAbDR[axis], aq = kAbDQ[axis];
        if (size == 1) { const int nr = r + dr, nq = q + dq; return on_board(b, nr, nq) && at(b, nr, nq, player) == 0 && at(b, nr, nq, opp) == 0; }
        const int r1 = r + ar, q1 = q + aq;
        if (!on_board(b, r1, q1) || at(b, r1, q1, player) == 0) return false;
        if (size == 3) { const int r2 = r1 + ar, q2 = q1 + aq; if (!(on_board(b, r2, q2) && at(b, r2, q2, player) == 1)) return false; }
        const bool inline_move = d == axis || d == (axis + 3) % 6;
        if (!inline_move) {                                    // broadside: every target cell must be free
            for (int i = 0; i < size; i++) {
                const int tr = r + i * ar + dr, tq = q + i * aq + dq;
                if (!on_board(b, tr, tq) || at(b, tr, tq, player) == 1 || at(b, tr, tq, opp) == 1) return false;
            }
            return true;
        }
        const int fr = d == axis ? r + (size - 1) * ar : r, fq = d == axis ? q + (size - 1) * aq : q;
        const int tr = fr + dr, tq = fq + dq;
        if (!on_board(b, tr, tq)) return false;
        if (at(b, tr, tq, player) == 1) return false;
        if (at(b, tr, tq, opp) == 0) return true;
        int opp_count = 0, cr = tr, cq = tq;                   // sumito: fewer opponent marbles than ours, then free cell or the edge
        for (;;) {
            if (!on_board(b, cr, cq)) return opp_count > 0;
            if (at(b, cr, cq, opp) == 1) { if (++opp_count >= size) return false; cr += dr; cq += dq; }
            else if (at(b, cr, cq, player) == 1) return false;
            else return true;
        }
    }
    // WARP: legal-action bitmask into `w` (MASK_WORDS words of warp-private SHARED memory), visible to all lanes on return.
    // Only cells holding one of the mover's marbles can start an action (valid_moves :262: the first test of every action), so the
    // lanes enumerate (own marble, plane) pairs -- at most 14 x 42 = 588 instead of 3402 actions -- and set bits with shared atomics.
    static __device__ void valid_mask(const int8_t* b, int player, int lane, uint32_t* w) {
        for (int k = lane; k < MASK_WORDS; k += 32) w[k] = 0u;
        uint32_t own[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { const int c = lane + 32 * i; own[i] = __ballot_sync(FULL, c < 81 && b[4 * c + player] != 0); }
        const int n0 = __popc(own[0]), n1 = __popc(own[1]), n = n0 + n1 + __popc(own[2]);
        __syncwarp();
        for (int idx = lane; idx < n * 42; idx += 32) {
            int mi = idx / 42; const int plane = idx - 42 * mi;
            uint32_t m; int base;
            if (mi < n0) { m = own[0]; base = 0; } else if (mi < n0 + n1) { m = own[1]; base = 32; mi -= n0; } else { m = own[2]; base = 64; mi -= n0 + n1; }
            for (int j = 0; j < mi; j++) m &= m - 1;                       // drop the mi lowest marbles of this word
            const int a = (base + __ffs(m) - 1) * 42 + plane;              // r*378 + q*42 + plane with cell = 9r + q
            if (action_valid(b, a, player)) atomicOr(&w[a >> 5], 1u << (a & 31));
        }
        __syncwarp();
    }
    // LANE: make_move :333-374. Deterministic: `seed` / `rng` are unused. Returns the next player.
    static __device__ int make_move(int8_t* b, int move, int player, long long seed, Philox* rng) {
        int r, q, size, axis, d; decode(move, r, q, size, axis, d);
        const int opp = 1 - player, dr = kAbDR[d], dq = kAbDQ[d], ar = kAbDR[axis], aq = kAbDQ[axis];
        const bool inline_move = d == axis || d == (axis + 3) % 6;
        if (size == 1 || !inline_move) {
            for (int i = 0; i < size; i++) {
                const int cr = size > 1 ? r + i * ar : r, cq = size > 1 ? q + i * aq : q;
                b[4 * (9 * cr + cq) + player] = 0;
                b[4 * (9 * (cr + dr) + cq + dq) + player] = 1;
            }
        } else {
            int fr, fq, br, bq;
            if (d == axis) { fr = r + (size - 1) * ar; fq = q + (size - 1) * aq; br = r; bq = q; }
            else { fr = r; fq = q; br = r + (size - 1) * ar; bq = q + (size - 1) * aq; }
            const int tr = fr + dr, tq = fq + dq;
            if (on_board(b, tr, tq) && at(b, tr, tq, opp) == 1) {
                int cr = tr, cq = tq;
                while (on_board(b, cr, cq) && at(b, cr, cq, opp) == 1) { cr += dr; cq += dq; }
                b[4 * (9 * tr + tq) + opp] = 0;
                if (on_board(b, cr, cq)) b[4 * (9 * cr + cq) + opp] = 1;
                else b[4 * player + 3] = (int8_t)(b[4 * player + 3] + 1);             // misc[0, player] += 1: a marble left the board
            }
            b[4 * (9 * br + bq) + player] = 0;
            b[4 * (9 * tr + tq) + player] = 1;
        }
        b[11] = (int8_t)(b[11] + 1);
        return 1 - player;
    }
    // check_end_game :376-392. Every lane computes the same result.
    static __device__ bool ended(const int8_t* b, int next_player, float (&out)[NP], int lane) {
        out[0] = out[1] = 0.f;
        if (b[3] >= 6) { out[0] = 1.f; out[1] = -1.f; return true; }
        if (b[7] >= 6) { out[0] = -1.f; out[1] = 1.f; return true; }
        if (b[11] >= 127) {
            if (b[3] > b[7]) { out[0] = 1.f; out[1] = -1.f; }
            else if (b[7] > b[3]) { out[0] = -1.f; out[1] = 1.f; }
            else { out[0] = 0.001f; out[1] = 0.001f; }
            return true;
        }
        return false;
    }
    // WARP: swap_players :394-406.
    static __device__ void swap_players(int8_t* b, int nb_swaps, int lane) {
        if ((nb_swaps & 1) == 0) return;
        for (int c = lane; c < 81; c += 32) { const int8_t t = b[4 * c]; b[4 * c] = b[4 * c + 1]; b[4 * c + 1] = t; }
        __syncwarp();
        if (lane == 0) { const int8_t t = b[3]; b[3] = b[7]; b[7] = t; }
        __syncwarp();
    }
    // LANE: init_game :167-252, Belgian daisy (deterministic).
    static __device__ void init_game(int8_t* b, Philox* rng) {
        for (int i = 0; i < SP; i++) b[i] = 0;
        for (int r = 0; r < 9; r++) for (int q = 0; q < 9; q++) if (r + q >= 4 && r + q <= 12) b[4 * (9 * r + q) + 2] = 1;
        auto fill = [&](int ch, int r, int q0, int q1) { for (int q = q0; q < q1; q++) b[4 * (9 * r + q) + ch] = 1; };
        fill(1, 0, 4, 6); fill(1, 1, 3, 6); fill(1, 2, 3, 5); fill(1, 6, 4, 6); fill(1, 7, 3, 6); fill(1, 8, 3, 5);     // opponent (white)
        fill(0, 0, 7, 9); fill(0, 1, 6, 9); fill(0, 2, 6, 8); fill(0, 6, 1, 3); fill(0, 7, 0, 3); fill(0, 8, 0, 2);     // mover (black)
    }

    // get_symmetries :408-441: 6 rotations x 2 reflections, index k = 2*rot + flip.
    static __device__ int num_symmetries(const int8_t* b) { return MAX_SYM; }
    static __device__ __forceinline__ void xform(int& r, int& q, int rot, int flip) {
        if (flip) q = 12 - r - q;                               // reflect across the vertical axis
        for (int i = 0; i < rot; i++) { const int nr = q + r - 4, nq = 8 - r; r = nr; q = nq; }   // 60 degrees clockwise around (4,4)
    }
    static __device__ int map_action(int a, int rot, int flip) {                                   // _build_action_symmetries :95-148
        int r, q, size, axis, d; decode(a, r, q, size, axis, d);
        int mr[3], mq[3];
        for (int i = 0; i < size; i++) { mr[i] = r + i * kAbDR[axis]; mq[i] = q + i * kAbDQ[axis]; xform(mr[i], mq[i], rot, flip); }
        int mi = 0;
        for (int i = 1; i < size; i++) if (mr[i] < mr[mi] || (mr[i] == mr[mi] && mq[i] < mq[mi])) mi = i;
        int new_axis = 0;
        if (size > 1) {
            const int oi = mi == 0 ? 1 : 0, ddr = mr[oi] - mr[mi], ddq = mq[oi] - mq[mi];
            if (ddr == 0 && ddq > 0) new_axis = 0; else if (ddr > 0 && ddq == 0) new_axis = 1; else if (ddr > 0 && ddq < 0) new_axis = 2;
        }
        int nd = d;
        if (flip) nd = kAbFlipDir[nd];
        nd = (nd + rot) % 6;
        return encode(mr[mi], mq[mi], size, new_axis, nd);
    }
    static __device__ void symmetry(const int8_t* b, const float* pi, const uint8_t* mask, int k, int lane,
                                    int8_t* ob, float* opi, uint8_t* om) {
        const int rot = k >> 1, flip = k & 1;
        for (int c = lane; c < 81; c += 32) { ob[4 * c] = 0; ob[4 * c + 1] = 0; ob[4 * c + 2] = 0; ob[4 * c + 3] = b[4 * c + 3]; }   // misc layer untransformed
        for (int a = lane; a < A; a += 32) { opi[a] = 0.f; om[a] = 0; }
        __threadfence_block(); __syncwarp();
        for (int c = lane; c < 81; c += 32) {
            int r = c / 9, q = c % 9;
            if (b[4 * c + 2] != 1) continue;
            int nr = r, nq = q; xform(nr, nq, rot, flip);
            const int o = 4 * (9 * nr + nq);
            ob[o] = b[4 * c]; ob[o + 1] = b[4 * c + 1]; ob[o + 2] = b[4 * c + 2];
        }
        for (int a = lane; a < A; a += 32)
            if (mask[a]) { const int m = map_action(a, rot, flip); opi[m] = pi[a]; om[m] = mask[a]; }
    }
};

}  // namespace azg

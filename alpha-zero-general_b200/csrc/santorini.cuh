// santorini.cuh -- Santorini WITHOUT god powers (the reference built with NB_GODS = 1, santorini/SantoriniConstants.py:19)
// as __device__ code: the Board jitclass of santorini/SantoriniLogicNumba.py:78-729, no-god branches only
// (:135-151 valid moves, :454-475 make_move, :552-565 check_end_game, :567-576 swap_players, :578-653 symmetries).
//
// Board = the reference's int8[5][5][3] state, byte-compatible (HWC interleaved: cell c = 5*y+x at bytes 3c..3c+2):
//   ch0 workers (+1,+2 mover / -1,-2 opponent), ch1 levels 0..4 (4 = dome), ch2 "gods_power" read through
//   .flat[i] = byte 3i+2: flat[0], flat[1] = 64 (NO_GOD owned by player / opponent), flat[2] = round counter
//   (saturates at 127), flat[8..10] = the Pan / Athena slots the no-god code still reads (always 0 here).
// Action = 81*worker + 9*move_direction + build_direction, directions 0..8 row-major, 4 = none.
// WARP functions are called by all 32 lanes, LANE functions by one lane (caller brackets with __syncwarp()).
#pragma once
#include "common.cuh"

namespace azg {

__constant__ uint8_t kSanRot[9] = {6, 3, 0, 7, 4, 1, 8, 5, 2};      // rotation_core, SantoriniConstants.py:60
__constant__ uint8_t kSanFlipLR[9] = {2, 1, 0, 5, 4, 3, 8, 7, 6};   // flipLR_core,   SantoriniConstants.py:68
__constant__ uint8_t kSanFlipUD[9] = {6, 7, 8, 3, 4, 5, 0, 1, 2};   // flipUD_core,   SantoriniConstants.py:77

struct Santorini {
    static constexpr int GAME_ID = 2;
    static constexpr int NP = 2;
    static constexpr int D0 = 5, D1 = 5, D2 = 3;
    static constexpr int S = 75;                               // observation_size() = (5,5,3)
    static constexpr int SP = 80;                              // padded to 16 B
    static constexpr int A = 162;                              // action_size() = NB_GODS*2*9*9
    static constexpr int MASK_WORDS = 6;
    static constexpr int MAX_LEGAL = 128;                      // 2 workers x 8 moves x 8 builds
    static constexpr int EDGE_FACTOR = 56;                     // edge arena = node arena x this (mean legal moves ~45-57)
    static constexpr int MAX_MOVES = 104;                      // every move builds one level: at most 100 builds fit on the board
    static constexpr int MAX_DEPTH = MAX_MOVES + 4;
    static constexpr int MAX_SYM = 8;
    static constexpr int PAN = 8, ATHENA = 9;                  // SantoriniConstants.py:16-17 (NB_GODS = 1 => index = god + player)
    typedef uint8_t act_t;

    static __device__ __forceinline__ int wk(const int8_t* b, int c) { return b[3 * c]; }
    static __device__ __forceinline__ int lv(const int8_t* b, int c) { return b[3 * c + 1]; }
    static __device__ __forceinline__ int gp(const int8_t* b, int i) { return b[3 * i + 2]; }
    static __device__ __forceinline__ void set_gp(int8_t* b, int i, int v) { b[3 * i + 2] = (int8_t)v; }

    static __device__ __forceinline__ bool is_chance_move(int) { return false; }   // no chance in this game
    static __device__ __forceinline__ int round(const int8_t* b) { return gp(b, 2); }        // get_round :655-656
    static __device__ __forceinline__ int progress(const int8_t* b) { return round(b); }           // grows with every move (tree GC, tree.cuh)
    // get_score :87-101: highest level under one of the player's workers
    static __device__ int score(const int8_t* b, int player) {
        int best = 0;
        for (int c = 0; c < 25; c++) { const int w = wk(b, c); if ((player == 0 ? w > 0 : w < 0) && lv(b, c) > best) best = lv(b, c); }
        return best;
    }
    static __device__ int find_worker(const int8_t* b, int id) {                               // _get_worker_position :667-672
        for (int c = 0; c < 25; c++) if (wk(b, c) == id) return c;
        return -1;
    }
    // One action's legality (valid_moves no-god branch :135-151 with _able_to_move_worker_to :675-701, _able_to_build :719-729).
    static __device__ bool action_valid(const int8_t* b, int a, int player) {
        const int worker = a / 81, md = (a % 81) / 9, bd = a % 9;
        if (md == 4 || bd == 4) return false;
        if (gp(b, player) <= 0) return false;                  // NO_GOD flag of the player (always 64 in this mode)
        const bool no_climb = gp(b, ATHENA + (1 - player)) > 64;
        const int wid = (worker + 1) * (player == 0 ? 1 : -1);
        const int old = find_worker(b, wid);
        if (old < 0) return false;
        const int ny = old / 5 + md / 3 - 1, nx = old % 5 + md % 3 - 1;
        if (ny < 0 || ny >= 5 || nx < 0 || nx >= 5) return false;
        const int nc = 5 * ny + nx;
        if (wk(b, nc) != 0) return false;
        const int nl = lv(b, nc);
        if (nl > 3 || nl > lv(b, old) + (no_climb ? 0 : 1)) return false;
        const int by = ny + bd / 3 - 1, bx = nx + bd % 3 - 1;
        if (by < 0 || by >= 5 || bx < 0 || bx >= 5) return false;
        const int bc = 5 * by + bx;
        const int bw = wk(b, bc);
        if (!(bw == 0 || bw == wid)) return false;             // the moving worker has left its old cell
        return lv(b, bc) < 4;
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
    // WARP: does `player` have any legal move?
    static __device__ __forceinline__ bool any_valid(const int8_t* b, int player, int lane) {
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < MASK_WORDS; k++) { const int a = lane + 32 * k; any |= __ballot_sync(FULL, a < A && action_valid(b, a, player)); }
        return any != 0;
    }
    // LANE: make_move :434-550 (power == NO_GOD). Deterministic: `seed` / `rng` are unused. Returns the next player.
    static __device__ int make_move(int8_t* b, int move, int player, long long seed, Philox* rng) {
        const int worker = move / 81, md = (move % 81) / 9, bd = move % 9;
        const int wid = (worker + 1) * (player == 0 ? 1 : -1);
        const int old = find_worker(b, wid);
        if (old >= 0) {
            const int old_level = lv(b, old);
            const int nc = 5 * (old / 5 + md / 3 - 1) + (old % 5 + md % 3 - 1);
            b[3 * old] = 0; b[3 * nc] = (int8_t)wid;
            if (bd != 4) { const int bc = 5 * (nc / 5 + bd / 3 - 1) + (nc % 5 + bd % 3 - 1); b[3 * bc + 1] = (int8_t)(b[3 * bc + 1] + 1); }
            const int new_level = lv(b, nc);
            if (gp(b, PAN + player) > 0) { if (new_level <= old_level - 2) set_gp(b, PAN + player, 65); }
            else if (gp(b, ATHENA + player) > 0) set_gp(b, ATHENA + player, 64 + (new_level > old_level ? 1 : 0));
            else { const int v = gp(b, player); set_gp(b, player, v < 64 ? v : 64); }
        }
        if (gp(b, 2) < 127) set_gp(b, 2, gp(b, 2) + 1);        // round counter :543-545
        return 1 - player;
    }
    // WARP: check_end_game :552-565. `next_player` is the player to move on this board.
    static __device__ bool ended(const int8_t* b, int next_player, float (&out)[NP], int lane) {
        out[0] = out[1] = 0.f;
        if (score(b, 0) == 3 || gp(b, PAN) > 64) { out[0] = 1.f; out[1] = -1.f; return true; }
        if (score(b, 1) == 3 || gp(b, PAN + 1) > 64) { out[0] = -1.f; out[1] = 1.f; return true; }
        if (!any_valid(b, next_player, lane)) { out[next_player] = -1.f; out[1 - next_player] = 1.f; return true; }
        return false;
    }
    // WARP: swap_players :567-576.
    static __device__ void swap_players(int8_t* b, int nb_swaps, int lane) {
        if (nb_swaps != 1) return;
        if (lane < 25) b[3 * lane] = (int8_t)(-b[3 * lane]);
        if (lane == 31) { const int8_t t = b[2]; b[2] = b[5]; b[5] = t; }      // flat[0] <-> flat[1]
        __syncwarp();
    }
    // LANE: init_game :103-120 with INIT_METHOD = 1: four distinct random cells for workers 1, -1, 2, -2.
    static __device__ void init_game(int8_t* b, Philox* rng) {
        for (int i = 0; i < SP; i++) b[i] = 0;
        const int ids[4] = {1, -1, 2, -2};
        uint32_t used = 0;
        for (int i = 0; i < 4; i++) {
            int k = (int)(rng->uniformf() * (float)(25 - i)); if (k > 24 - i) k = 24 - i;
            int c = 0;
            for (int j = 0; j < 25; j++) if (!(used >> j & 1)) { if (k == 0) { c = j; break; } k--; }
            used |= 1u << c;
            b[3 * c] = (int8_t)ids[i];
        }
        set_gp(b, 0, 64); set_gp(b, 1, 64);
    }

    // get_symmetries :578-653: identity, rot90 x1..x3, flipLR, flipUD, swap own workers, swap opponent workers.
    static __device__ int num_symmetries(const int8_t* b) { return MAX_SYM; }
    static __device__ __forceinline__ int perm_dir(int d, int k) {
        if (k >= 1 && k <= 3) { for (int i = 0; i < k; i++) d = kSanRot[d]; return d; }
        if (k == 4) return kSanFlipLR[d];
        if (k == 5) return kSanFlipUD[d];
        return d;
    }
    static __device__ void symmetry(const int8_t* b, const float* pi, const uint8_t* mask, int k, int lane,
                                    int8_t* ob, float* opi, uint8_t* om) {
        for (int c = lane; c < 25; c += 32) {
            int y = c / 5, x = c % 5, sy = y, sx = x;           // out[y][x] = in[sy][sx]
            if (k >= 1 && k <= 3) { for (int i = 0; i < k; i++) { const int ty = sx, tx = 4 - sy; sy = ty; sx = tx; } }   // np.rot90: out[i][j] = in[j][4-i]
            else if (k == 4) sx = 4 - x;
            else if (k == 5) sy = 4 - y;
            const int sc = 5 * sy + sx;
            int w = b[3 * sc];
            if (k == 6 && w > 0) w = 3 - w;                      // workers 1 <-> 2
            if (k == 7 && w < 0) w = -3 - w;                     // workers -1 <-> -2
            ob[3 * c] = (int8_t)w; ob[3 * c + 1] = b[3 * sc + 1]; ob[3 * c + 2] = b[3 * c + 2];
        }
        for (int a = lane; a < A; a += 32) {
            const int worker = a / 81, md = (a % 81) / 9, bd = a % 9;
            int dst;
            if (k == 6) dst = (1 - worker) * 81 + md * 9 + bd;  // policy halves swap with the own workers
            else dst = worker * 81 + perm_dir(md, k) * 9 + perm_dir(bd, k);
            opi[dst] = pi[a]; om[dst] = mask[a];
        }
    }
};

}  // namespace azg

// splendor.cuh -- Splendor rules as __device__ code (the Board jitclass of the reference,
// splendor/SplendorLogicNumba.py:138-479, re-designed for one-warp-per-game execution).
//
// The board is the reference's int8[ROWS][7] state, byte-compatible, held in shared
// memory by the warp that owns the game. Functions marked WARP are warp-cooperative (all 32 lanes
// call them); functions marked LANE are executed by a single lane (scalar rule code) and the caller
// brackets them with __syncwarp().
#pragma once
#include "common.cuh"

namespace azg {

// Card data (public rules of the game; values as in splendor/SplendorLogic.py:127-280), packed:
// cost of colour k in nibble k, points in bits 20..23. Index [tier][deck colour][card]; a card of
// deck colour c gives a bonus of colour kGainCol[c].
__constant__ uint32_t kCards[3][5][8] = {
    {{0x030000, 0x020001, 0x020200, 0x002201, 0x001310, 0x011101, 0x012101, 0x104000},
     {0x000003, 0x000120, 0x002002, 0x020102, 0x031001, 0x010111, 0x010112, 0x100004},
     {0x000300, 0x001200, 0x000202, 0x001022, 0x013100, 0x001111, 0x001121, 0x100040},
     {0x000030, 0x012000, 0x020020, 0x010220, 0x010013, 0x011110, 0x011210, 0x100400},
     {0x003000, 0x000012, 0x002020, 0x022010, 0x000131, 0x011011, 0x021011, 0x140000}},
    {{0x103220, 0x130320, 0x200050, 0x200035, 0x241002, 0x300060, 0, 0},
     {0x132002, 0x132030, 0x250000, 0x250003, 0x200241, 0x306000, 0, 0},
     {0x100223, 0x120303, 0x200005, 0x203500, 0x202410, 0x360000, 0, 0},
     {0x122300, 0x103032, 0x205000, 0x235000, 0x224100, 0x300006, 0, 0},
     {0x120032, 0x103203, 0x200500, 0x200350, 0x210024, 0x300600, 0, 0}},
    {{0x353303, 0x400007, 0x430036, 0x500037, 0, 0, 0, 0},
     {0x330353, 0x400700, 0x403630, 0x503700, 0, 0, 0, 0},
     {0x303533, 0x407000, 0x436300, 0x537000, 0, 0, 0, 0},
     {0x335330, 0x470000, 0x463003, 0x570003, 0, 0, 0, 0},
     {0x333035, 0x400070, 0x400363, 0x500370, 0, 0, 0, 0}}};
__constant__ uint32_t kNobles[10] = {0x304400, 0x344000, 0x300440, 0x340004, 0x300044,
                                     0x333003, 0x300333, 0x333300, 0x303330, 0x330033};
__constant__ int kGainCol[5] = {1, 3, 4, 0, 2};
// colour subsets in itertools.combinations order (sizes 1,2,3), bit k = colour k (SplendorLogic.py:76-87)
__constant__ uint8_t kGems3[25] = {1, 2, 4, 8, 16, 3, 5, 9, 17, 6, 10, 18, 12, 20, 24, 7, 11, 19, 13, 21, 25, 14, 22, 26, 28};
__constant__ uint8_t kCardPerm[3][4] = {{1, 3, 0, 2}, {2, 0, 3, 1}, {3, 2, 1, 0}};                     // SplendorLogic.py:89
__constant__ int8_t kResPerm[4][2][3] = {{{-1, -1, -1}, {-1, -1, -1}}, {{-1, -1, -1}, {-1, -1, -1}},  // SplendorLogic.py:97-102
                                         {{1, 0, 2}, {-1, -1, -1}}, {{1, 2, 0}, {2, 0, 1}}};

template <int NP_>
struct Splendor {
    static constexpr int GAME_ID = 1;
    static constexpr int NP = NP_;
    static constexpr int NN = NP + 1;                         // nobles in play
    static constexpr int COLS = 7;
    static constexpr int ROWS = 32 + 10 * NP + NP * NP;      // observation_size(), SplendorLogicNumba.py:90-92
    static constexpr int D0 = ROWS, D1 = COLS, D2 = 1;
    static constexpr int S = ROWS * COLS;                     // 392 bytes for 2 players
    static constexpr int SP = (S + 15) / 16 * 16;             // padded to 16 B for 128-bit loads (400)
    static constexpr int A = 81;
    static constexpr int MASK_WORDS = 3;
    static constexpr int MAX_LEGAL = 64;                      // upper bound used to reserve edge space (largest seen: 62)
    static constexpr int EDGE_FACTOR = 44;                    // edge arena = node arena x this (mean legal moves per expanded node ~34-42)
    static constexpr int MAX_MOVES = 62 * NP;                 // SplendorLogicNumba.py:146
    static constexpr int MAX_DEPTH = MAX_MOVES + 4;
    static constexpr int MAX_SYM = 1 + 9 + 2 * NP;
    static constexpr int GEMS = NP == 2 ? 4 : NP == 3 ? 5 : 7;
    // row map (copy_state, SplendorLogicNumba.py:207-219)
    static constexpr int R_BANK = 0, R_CARDS = 1, R_DECK = 25, R_NOBLES = 31, R_PGEMS = 32 + NP, R_PNOBLES = 32 + 2 * NP,
                         R_PCARDS = 32 + 3 * NP + NP * NP, R_PRES = 32 + 4 * NP + NP * NP;
    typedef uint8_t act_t;

    static __device__ __forceinline__ int8_t* row(int8_t* b, int r) { return b + r * COLS; }
    static __device__ __forceinline__ const int8_t* row(const int8_t* b, int r) { return b + r * COLS; }
    static __device__ __forceinline__ int sum5(const int8_t* r) { return r[0] + r[1] + r[2] + r[3] + r[4]; }
    static __device__ __forceinline__ int sum7(const int8_t* r) { return sum5(r) + r[5] + r[6]; }

    // get_round (SplendorLogicNumba.py:303-304)
    static __device__ __forceinline__ int round(const int8_t* b) { return (uint8_t)b[6]; }
    static __device__ __forceinline__ int progress(const int8_t* b) { return round(b); }           // grows with every move (tree GC, tree.cuh)
    // get_score (SplendorLogicNumba.py:151-154)
    static __device__ int score(const int8_t* b, int player) {
        int s = row(b, R_PCARDS + player)[6];
        for (int i = 0; i < NN; i++) s += row(b, R_PNOBLES + NN * player + i)[6];
        return s;
    }

    static __device__ __forceinline__ bool can_buy(const int8_t* cost, const int8_t* pg, const int8_t* pc) {
        int missing = 0, total = 0;
#pragma unroll
        for (int c = 0; c < 5; c++) { int d = (int8_t)(cost[c] - pg[c] - pc[c]); missing += d > 0 ? d : 0; total += cost[c]; }
        return missing <= pg[5] && total != 0;
    }

    // One action's legality (valid_moves and its helpers, SplendorLogicNumba.py:180-188,359-368,375-380,
    // 402-412,422-453). Pure function of the board: any lane may evaluate any action.
    static __device__ bool action_valid(const int8_t* b, int a, int player) {
        const int8_t* bank = row(b, R_BANK);
        const int8_t* pg = row(b, R_PGEMS + player);
        const int8_t* pc = row(b, R_PCARDS + player);
        const int8_t* res = row(b, R_PRES + 6 * player);
        if (a < 12) return can_buy(row(b, R_CARDS + 2 * a), pg, pc);
        if (a < 27) {
            bool empty_slot = sum5(res + 5 * COLS) == 0;          // gain row of the 3rd reserve slot
            int i = a - 12;
            int nz = i < 12 ? sum5(row(b, R_CARDS + 2 * i)) : sum5(row(b, R_DECK + 2 * (i - 12)));
            return nz != 0 && empty_slot;
        }
        if (a < 30) return can_buy(res + 2 * (a - 27) * COLS, pg, pc);
        if (a < 55) {
            int m = kGems3[a - 30], k = 0; bool ok = true;
#pragma unroll
            for (int c = 0; c < 5; c++) if (m >> c & 1) { k++; ok = ok && bank[c] >= 1; }
            return ok && sum7(pg) + k <= 10;
        }
        if (a < 60) return bank[a - 55] >= 4 && sum7(pg) + 2 <= 10;
        if (a < 75) {
            int m = kGems3[a - 60]; bool ok = true;                // first 15 subsets = sizes 1,2
#pragma unroll
            for (int c = 0; c < 5; c++) if (m >> c & 1) ok = ok && pg[c] >= 1;
            return ok;
        }
        if (a < 80) return pg[a - 75] >= 2;
        return true;                                               // 80: pass is always legal
    }

    // WARP: legal-action bitmask into `w` (MASK_WORDS words of warp-private shared memory), visible to all lanes on return.
    static __device__ __forceinline__ void valid_mask(const int8_t* b, int player, int lane, uint32_t* w) {
#pragma unroll 1
        for (int k = 0; k < MASK_WORDS; k++) {                  // not unrolled: one copy of the (large) per-action test keeps the kernels' code small
            int a = lane + 32 * k;
            bool v = a < A && action_valid(b, a, player);
            const uint32_t m = __ballot_sync(FULL, v);
            if (lane == 0) w[k] = m;
        }
        __syncwarp();
    }

    static __device__ __forceinline__ void write_card(int8_t* rows2, int tier, int colour, int idx) {
        uint32_t code = kCards[tier][colour][idx];
#pragma unroll
        for (int k = 0; k < 2 * COLS; k++) rows2[k] = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) rows2[k] = (int8_t)((code >> (4 * k)) & 15);
        rows2[COLS + kGainCol[colour]] = 1;
        rows2[COLS + 6] = (int8_t)((code >> 20) & 15);
    }

    // LANE: _get_deck_card (SplendorLogicNumba.py:306-342). seed != 0: deterministic pseudo-random index
    // (4594591*(seed + sum_c bits_c*32^c)) mod n, Python modulo. seed == 0: uniform over the remaining
    // cards (= colour proportional to count, then uniform inside the colour) from the Philox stream.
    static __device__ bool get_deck_card(int8_t* b, int tier, long long seed, Philox* rng, int8_t* out2) {
        int8_t* cnt = row(b, R_DECK + 2 * tier);
        int8_t* bits = row(b, R_DECK + 2 * tier + 1);
        int total = sum5(cnt);
        if (total == 0) return false;
        int m = 0; long long s = 0, mul = 1; uint32_t f[5];
#pragma unroll
        for (int c = 0; c < 5; c++) { f[c] = (uint8_t)bits[c]; m += __popc(f[c]); s += (long long)f[c] * mul; mul *= 32; }
        int k;
        if (seed == 0) {
            k = (int)(rng->uniformf() * (float)m); if (k >= m) k = m - 1;
        } else {
            long long x = 4594591LL * (seed + s);
            long long r = x % m; if (r < 0) r += m;
            k = (int)r;
        }
        int colour = 0, idx = 0;
#pragma unroll
        for (int c = 0; c < 5; c++) {                          // k-th set bit in (colour, MSB-first) order
            int pc = __popc(f[c]);
            if (k >= 0 && k < pc) {
                uint32_t rev = __brev(f[c]) >> 24;             // MSB-first position i  <->  bit i of rev
                uint32_t t = rev;
#pragma unroll 1
                for (int j = 0; j < k; j++) t &= t - 1;           // drop the k lowest set bits (kept rolled: this sits in every kernel that moves)
                colour = c; idx = __ffs(t) - 1; k = -1;
            } else if (k >= 0) k -= pc;
        }
        bits[colour] = (int8_t)(f[colour] & ~(128u >> idx));
        cnt[colour] -= 1;
        write_card(out2, tier, colour, idx);
        return true;
    }

    // LANE: _fill_new_card (SplendorLogicNumba.py:325-329)
    static __device__ void fill_new_card(int8_t* b, int tier, int index, long long seed, Philox* rng) {
        int8_t* slot = row(b, R_CARDS + 8 * tier + 2 * index);
        int8_t card[2 * COLS];
        bool got = get_deck_card(b, tier, seed, rng, card);
#pragma unroll
        for (int k = 0; k < 2 * COLS; k++) slot[k] = got ? card[k] : (int8_t)0;
    }

    // LANE: _give_nobles_if_earned (SplendorLogicNumba.py:465-470)
    static __device__ void give_nobles(int8_t* b, int player) {
        const int8_t* pc = row(b, R_PCARDS + player);
        for (int i = 0; i < NN; i++) {
            int8_t* noble = row(b, R_NOBLES + i);
            if (sum5(noble) <= 0) continue;
            bool ok = true;
#pragma unroll
            for (int c = 0; c < 5; c++) ok = ok && pc[c] >= noble[c];
            if (ok) {
                int8_t* dst = row(b, R_PNOBLES + NN * player + i);
#pragma unroll
                for (int c = 0; c < COLS; c++) { dst[c] = noble[c]; noble[c] = 0; }
            }
        }
    }

    // LANE: _buy_card (SplendorLogicNumba.py:331-357); card rows are passed by value
    static __device__ void buy_card(int8_t* b, const int8_t* card0, const int8_t* card1, int player) {
        int8_t* bank = row(b, R_BANK); int8_t* pg = row(b, R_PGEMS + player); int8_t* pc = row(b, R_PCARDS + player);
        int missing = 0;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            int d = (int8_t)(card0[c] - pg[c] - pc[c]); missing += d > 0 ? d : 0;
            int need = (int8_t)(card0[c] - pc[c]); need = need > 0 ? need : 0;
            int paid = need < pg[c] ? need : pg[c];
            pg[c] = (int8_t)(pg[c] - paid); bank[c] = (int8_t)(bank[c] + paid);
        }
        pg[5] = (int8_t)(pg[5] - missing); bank[5] = (int8_t)(bank[5] + missing);
#pragma unroll
        for (int c = 0; c < COLS; c++) pc[c] = (int8_t)(pc[c] + card1[c]);
        give_nobles(b, player);
    }

    // True if the move may consume the chance seed (a card is revealed from a deck): moves 0-26 buy / reserve from the table or a deck.
    // Everything else (buy a reserved card, take / give back gems, pass) gives the same child state in every universe.
    static __device__ __forceinline__ bool is_chance_move(int move) { return move < 27; }
    // LANE: make_move (SplendorLogicNumba.py:190-205) with _buy :370-373, _reserve :382-400,
    // _buy_reserve :414-420, _get_gems / _give_gems :436-463. Returns the next player.
    static __device__ int make_move(int8_t* b, int move, int player, long long seed, Philox* rng) {
        int8_t* bank = row(b, R_BANK); int8_t* pg = row(b, R_PGEMS + player);
        int8_t c0[COLS], c1[COLS];
        if (move < 12) {
            const int8_t* card = row(b, R_CARDS + 2 * move);
#pragma unroll
            for (int c = 0; c < COLS; c++) { c0[c] = card[c]; c1[c] = card[COLS + c]; }
            buy_card(b, c0, c1, player);
            fill_new_card(b, move >> 2, move & 3, seed, rng);
        } else if (move < 27) {
            int i = move - 12; int8_t* res = row(b, R_PRES + 6 * player); int slot = -1;
            for (int k = 2; k >= 0; k--) if (sum5(res + 2 * k * COLS) == 0) slot = k;       // first empty slot
            if (slot >= 0) {
                int8_t* dst = res + 2 * slot * COLS;
                if (i < 12) {
                    const int8_t* src = row(b, R_CARDS + 2 * i);
#pragma unroll
                    for (int c = 0; c < 2 * COLS; c++) dst[c] = src[c];
                    fill_new_card(b, i >> 2, i & 3, seed, rng);
                } else {
                    int8_t card[2 * COLS];
                    if (get_deck_card(b, i - 12, seed, rng, card)) {
#pragma unroll
                        for (int c = 0; c < 2 * COLS; c++) dst[c] = card[c];
                    }
                }
            }
            if (bank[5] > 0 && sum7(pg) <= 9) { pg[5] += 1; bank[5] -= 1; }
        } else if (move < 30) {
            int i = move - 27; int8_t* res = row(b, R_PRES + 6 * player); int8_t* card = res + 2 * i * COLS;
#pragma unroll
            for (int c = 0; c < COLS; c++) { c0[c] = card[c]; c1[c] = card[COLS + c]; }
            buy_card(b, c0, c1, player);
            for (int k = 2 * i * COLS; k < 4 * COLS; k++) res[k] = res[k + 2 * COLS];        // shift the reserve left
            for (int k = 4 * COLS; k < 6 * COLS; k++) res[k] = 0;
        } else if (move < 60) {
            int i = move - 30;
            if (i < 25) { int m = kGems3[i];
#pragma unroll
                for (int c = 0; c < 5; c++) if (m >> c & 1) { bank[c] -= 1; pg[c] += 1; } }
            else { bank[i - 25] -= 2; pg[i - 25] += 2; }
        } else if (move < 80) {
            int i = move - 60;
            if (i < 15) { int m = kGems3[i];
#pragma unroll
                for (int c = 0; c < 5; c++) if (m >> c & 1) { bank[c] += 1; pg[c] -= 1; } }
            else { bank[i - 15] += 2; pg[i - 15] -= 2; }
        }
        bank[6] = (int8_t)(bank[6] + 1);                       // round counter
        return (player + 1) % NP;
    }

    // check_end_game (SplendorLogicNumba.py:221-240). Every lane computes the same result; returns true if the game is over.
    // (`next_player` is part of the Game.getGameEnded signature, Game.py:64; Splendor does not use it.)
    static __device__ bool ended(const int8_t* b, int next_player, float (&out)[NP], int lane) {
#pragma unroll
        for (int p = 0; p < NP; p++) out[p] = 0.f;
        int rnd = round(b);
        if (rnd % NP != 0) return false;
        float sc[NP], mx = -1e30f;
#pragma unroll
        for (int p = 0; p < NP; p++) { sc[p] = (float)score(b, p); mx = fmaxf(mx, sc[p]); }
        if (!(mx >= 15.f || rnd >= MAX_MOVES)) return false;
        int winners = 0;
#pragma unroll
        for (int p = 0; p < NP; p++) winners += sc[p] == mx;
        bool several = winners > 1;
        if (several) {                                          // tie-break: fewer cards (penalty cards/100)
#pragma unroll
            for (int p = 0; p < NP; p++) {
                int cards = sum5(row(b, R_PCARDS + p));
                sc[p] = (float)__dsub_rn((double)sc[p], __ddiv_rn((double)cards, 100.0));
            }
            mx = -1e30f;
#pragma unroll
            for (int p = 0; p < NP; p++) mx = fmaxf(mx, sc[p]);
            winners = 0;
#pragma unroll
            for (int p = 0; p < NP; p++) winners += sc[p] == mx;
            several = winners > 1;
        }
#pragma unroll
        for (int p = 0; p < NP; p++) out[p] = sc[p] == mx ? (several ? 0.01f : 1.f) : -1.f;
        return true;
    }

    // WARP: swap_players (SplendorLogicNumba.py:244-253): new_row[i] = old_row[(i+shift) % size] on the four
    // player-owned row groups. Reads complete before any write (register staging + __syncwarp).
    static __device__ void swap_players(int8_t* b, int nb_swaps, int lane) {
        constexpr int G0 = NP * COLS, G1 = NP * NN * COLS, G2 = NP * COLS, G3 = 6 * NP * COLS;
        constexpr int TOTAL = G0 + G1 + G2 + G3;
        constexpr int PER = (TOTAL + 31) / 32;
        int8_t val[PER]; int dst[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            int i = lane + 32 * k; dst[k] = -1; val[k] = 0;
            if (i < TOTAL) {
                int base, size, unit, j;
                if (i < G0) { base = R_PGEMS; size = NP; unit = 1; j = i; }
                else if (i < G0 + G1) { base = R_PNOBLES; size = NP * NN; unit = NN; j = i - G0; }
                else if (i < G0 + G1 + G2) { base = R_PCARDS; size = NP; unit = 1; j = i - G0 - G1; }
                else { base = R_PRES; size = 6 * NP; unit = 6; j = i - G0 - G1 - G2; }
                int r = j / COLS, c = j - r * COLS;
                int src = (r + unit * nb_swaps) % size;
                val[k] = b[(base + src) * COLS + c];
                dst[k] = (base + r) * COLS + c;
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < PER; k++) if (dst[k] >= 0) b[dst[k]] = val[k];
        __syncwarp();
    }

    // LANE: init_game (SplendorLogicNumba.py:156-178): bank, decks, 12 random visible cards, NN random nobles.
    static __device__ void init_game(int8_t* b, Philox* rng) {
        for (int i = 0; i < SP; i++) b[i] = 0;
        int8_t* bank = row(b, R_BANK);
        for (int c = 0; c < 5; c++) bank[c] = GEMS;
        bank[5] = 5;
        for (int t = 0; t < 3; t++) {
            int n = t == 0 ? 8 : t == 1 ? 6 : 4;
            for (int c = 0; c < 5; c++) { row(b, R_DECK + 2 * t)[c] = (int8_t)n; row(b, R_DECK + 2 * t + 1)[c] = (int8_t)(uint8_t)(0xFF << (8 - n)); }
        }
        for (int t = 0; t < 3; t++) for (int i = 0; i < 4; i++) fill_new_card(b, t, i, 0, rng);
        uint32_t used = 0;
        for (int i = 0; i < NN; i++) {                          // np.random.choice(10, NN, replace=False)
            int k = (int)(rng->uniformf() * (float)(10 - i)); if (k > 9 - i) k = 9 - i;
            int idx = 0;
            for (int j = 0; j < 10; j++) if (!(used >> j & 1)) { if (k == 0) { idx = j; break; } k--; }
            used |= 1u << idx;
            uint32_t code = kNobles[idx]; int8_t* r = row(b, R_NOBLES + i);
            for (int c = 0; c < 5; c++) r[c] = (int8_t)((code >> (4 * c)) & 15);
            r[6] = (int8_t)((code >> 20) & 15);
        }
    }

    // number of reserved cards (SplendorLogicNumba.py:472-476)
    static __device__ int nb_reserved(const int8_t* b, int player) {
        const int8_t* res = row(b, R_PRES + 6 * player);
        for (int c = 0; c < 3; c++) if (sum5(res + 2 * c * COLS) == 0) return c;
        return 3;
    }

    // WARP: get_symmetries (SplendorLogicNumba.py:255-301). Symmetry k of (board, pi, mask) is written to
    // (ob, opi, om) by the calling warp; returns false when k is past the number of symmetries of this
    // board. Order: identity, 3 tiers x 3 slot permutations, then per player the reserve permutations.
    static __device__ int num_symmetries(const int8_t* b) {
        int k = 10;
        for (int p = 0; p < NP; p++) { int nr = nb_reserved(b, p); k += nr == 2 ? 1 : nr == 3 ? 2 : 0; }
        return k;
    }
    static __device__ void symmetry(const int8_t* b, const float* pi, const uint8_t* mask, int k, int lane,
                                    int8_t* ob, float* opi, uint8_t* om) {
        // decode k -> (kind, tier/player, permutation)
        int kind = 0, t = 0, q = 0;   // kind 0 identity, 1 card perm (tier t, perm q), 2 reserve perm (player t, perm q)
        if (k >= 1 && k < 10) { kind = 1; t = (k - 1) / 3; q = (k - 1) % 3; }
        else if (k >= 10) {
            int r = k - 10; kind = 2;
            for (int p = 0; p < NP; p++) {
                int nr = nb_reserved(b, p); int cnt = nr == 2 ? 1 : nr == 3 ? 2 : 0;
                if (r < cnt) { t = p; q = r; break; }
                r -= cnt;
            }
        }
        int nres = kind == 2 ? nb_reserved(b, t) : 0;
        for (int i = lane; i < S; i += 32) {
            int r = i / COLS, c = i - r * COLS, src = r;
            if (kind == 1 && r >= R_CARDS + 8 * t && r < R_CARDS + 8 * t + 8) {
                int rr = r - (R_CARDS + 8 * t); src = R_CARDS + 8 * t + 2 * kCardPerm[q][rr >> 1] + (rr & 1);
            } else if (kind == 2 && r >= R_PRES + 6 * t && r < R_PRES + 6 * t + 6) {
                int rr = r - (R_PRES + 6 * t); src = R_PRES + 6 * t + 2 * kResPerm[nres][q][rr >> 1] + (rr & 1);
            }
            ob[i] = b[src * COLS + c];
        }
        for (int a = lane; a < A; a += 32) {
            int src = a;
            if (kind == 1) {
                if (a >= 4 * t && a < 4 * t + 4) src = 4 * t + kCardPerm[q][a - 4 * t];
                else if (a >= 12 + 4 * t && a < 16 + 4 * t) src = 12 + 4 * t + kCardPerm[q][a - 12 - 4 * t];
            } else if (kind == 2 && t == 0 && a >= 27 && a < 30) src = 27 + kResPerm[nres][q][a - 27];
            opi[a] = pi[src]; om[a] = mask[src];
        }
    }
};

}  // namespace azg

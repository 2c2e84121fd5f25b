/*
 * azg_oracle.c -- CPU restatement of the reference's self-play hot path (TEST INFRASTRUCTURE).
 *
 * This file is the parity checker and the CPU baseline. It is NOT on the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it. The product (alpha-zero-general_b200/csrc) never links or calls it.
 *
 * Pinned against the reference itself: tests/golden/*.npz are produced by
 * oracle/gen_golden.py, which imports and runs the unmodified Python reference in the
 * build container; tests/test_oracle_*.py check every function below against them.
 *
 * Each function cites the reference code it follows (paths relative to the reference root).
 * Floating-point operation ORDER follows the machine code numba 0.65 / LLVM emits for the
 * reference's @njit(fastmath=True) helpers on the build host (x86-64 AVX2+FMA), which is
 * what produced the golden vectors:
 *   pick_highest_UCB (MCTS.py:210-230):  c1 = cpuct*sqrt(Ns) hoisted;  visited  u = Q + (c1*P)/(1+N)
 *                                        c0 = cpuct*sqrt(Ns+1e-8);     unvisited u = fma(c0, P, fpu_init)
 *   normalise (MCTS.py:250-253):         float32 sum in 4x8-lane AVX2 order, then x *= (1/sum)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int8_t i8;
typedef uint8_t u8;

#define COLS 7
#define NA 81            /* action_size(), splendor/SplendorLogicNumba.py:94-96 */
#define MAXA 3402        /* largest action space the MCTS port handles (Abalone) */
#define MAXP 4
#define MAXROWS 88       /* observation_size(4) = 32+40+16 rows */
#define MAXS (MAXROWS * COLS)

/* ---------------------------------------------------------------- game data ------------- */
/* Card/noble data of the board game, restated from splendor/SplendorLogic.py:127-280 in a
 * packed form: cost of colour k in nibble k (bits 4k..4k+3), points in bits 20..23.
 * Index [deck colour][card index]; the deck colour index c yields a bonus of colour GAIN_COL[c]. */
static const int GAIN_COL[5] = {1, 3, 4, 0, 2};
static const uint32_t CARDS_T0[5][8] = {
    {0x030000, 0x020001, 0x020200, 0x002201, 0x001310, 0x011101, 0x012101, 0x104000},
    {0x000003, 0x000120, 0x002002, 0x020102, 0x031001, 0x010111, 0x010112, 0x100004},
    {0x000300, 0x001200, 0x000202, 0x001022, 0x013100, 0x001111, 0x001121, 0x100040},
    {0x000030, 0x012000, 0x020020, 0x010220, 0x010013, 0x011110, 0x011210, 0x100400},
    {0x003000, 0x000012, 0x002020, 0x022010, 0x000131, 0x011011, 0x021011, 0x140000}};
static const uint32_t CARDS_T1[5][6] = {
    {0x103220, 0x130320, 0x200050, 0x200035, 0x241002, 0x300060},
    {0x132002, 0x132030, 0x250000, 0x250003, 0x200241, 0x306000},
    {0x100223, 0x120303, 0x200005, 0x203500, 0x202410, 0x360000},
    {0x122300, 0x103032, 0x205000, 0x235000, 0x224100, 0x300006},
    {0x120032, 0x103203, 0x200500, 0x200350, 0x210024, 0x300600}};
static const uint32_t CARDS_T2[5][4] = {
    {0x353303, 0x400007, 0x430036, 0x500037},
    {0x330353, 0x400700, 0x403630, 0x503700},
    {0x303533, 0x407000, 0x436300, 0x537000},
    {0x335330, 0x470000, 0x463003, 0x570003},
    {0x333035, 0x400070, 0x400363, 0x500370}};
static const uint32_t NOBLES[10] = {0x304400, 0x344000, 0x300440, 0x340004, 0x300044,
                                    0x333003, 0x300333, 0x333300, 0x303330, 0x330033};
static const int DECK_SIZE[3] = {8, 6, 4};
/* colour subsets in itertools.combinations order, sizes 1,2,3 (SplendorLogic.py:76-87); bit k = colour k */
static const u8 GEMS3[25] = {1, 2, 4, 8, 16, 3, 5, 9, 17, 6, 10, 18, 12, 20, 24, 7, 11, 19, 13, 21, 25, 14, 22, 26, 28};
static const u8 GEMS2[15] = {1, 2, 4, 8, 16, 3, 5, 9, 17, 6, 10, 18, 12, 20, 24};

static uint32_t card_code(int tier, int colour, int idx) {
    return tier == 0 ? CARDS_T0[colour][idx] : tier == 1 ? CARDS_T1[colour][idx] : CARDS_T2[colour][idx];
}

/* ---------------------------------------------------------------- RNG (oracle-own) ------ */
typedef struct { uint64_t s[4]; } azo_rng;
static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rng_seed(azo_rng* r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed); }
static uint64_t rng_next(azo_rng* r) {
    uint64_t* s = r->s; uint64_t res = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return res;
}
static double rng_uniform(azo_rng* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static double rng_normal(azo_rng* r) {
    double u1 = rng_uniform(r), u2 = rng_uniform(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static double rng_gamma(azo_rng* r, double a) { /* Marsaglia-Tsang */
    if (a < 1.0) { double u = rng_uniform(r); if (u < 1e-300) u = 1e-300; return rng_gamma(r, a + 1.0) * pow(u, 1.0 / a); }
    double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x = rng_normal(r), v = 1.0 + c * x;
        if (v <= 0) continue;
        v = v * v * v;
        double u = rng_uniform(r);
        if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
    }
}

/* ---------------------------------------------------------------- board layout ---------- */
/* Row map: SplendorLogicNumba.py:207-219 (copy_state). */
typedef struct { int n, nn, rows, r_nobles, r_pgems, r_pnobles, r_pcards, r_pres; } layout_t;
static layout_t layout(int n) {
    layout_t L; L.n = n; L.nn = n + 1; L.rows = 32 + 10 * n + n * n;
    L.r_nobles = 31; L.r_pgems = 32 + n; L.r_pnobles = 32 + 2 * n; L.r_pcards = 32 + 3 * n + n * n; L.r_pres = 32 + 4 * n + n * n;
    return L;
}
#define R_BANK 0
#define R_CARDS 1
#define R_DECK 25
#define ROW(b, r) ((b) + (r) * COLS)

int azo_state_rows(int n) { return 32 + 10 * n + n * n; }

static int sum5(const i8* row) { return row[0] + row[1] + row[2] + row[3] + row[4]; }
static int sum7(const i8* row) { return sum5(row) + row[5] + row[6]; }
static void write_card(i8* rows2, int tier, int colour, int idx) {
    uint32_t code = card_code(tier, colour, idx);
    memset(rows2, 0, 2 * COLS);
    for (int k = 0; k < 5; k++) rows2[k] = (i8)((code >> (4 * k)) & 15);
    rows2[COLS + GAIN_COL[colour]] = 1;
    rows2[COLS + 6] = (i8)((code >> 20) & 15);
}

/* SplendorLogicNumba.py:306-342 (_get_deck_card). Returns 0 if the deck is empty, else writes the
 * 2-row card to `out`. seed==0: true random (colour ~ remaining count, then uniform among that
 * colour's cards) drawn from `rng`; else the deterministic index (4594591*(seed+sum bits*32^c)) mod n. */
static int get_deck_card(i8* b, int tier, int64_t seed, azo_rng* rng, i8* out) {
    i8* cnt = ROW(b, R_DECK + 2 * tier);
    i8* bits = ROW(b, R_DECK + 2 * tier + 1);
    int total = sum5(cnt);
    if (total == 0) return 0;
    int colour = -1, idx = -1;
    if (seed == 0) {
        double r = rng_uniform(rng), acc = 0;
        for (int c = 0; c < 5; c++) { acc += (double)cnt[c] / (double)total; if (acc > r) { colour = c; break; } }
        if (colour < 0) for (int c = 4; c >= 0; c--) if (cnt[c] > 0) { colour = c; break; }
        u8 f = (u8)bits[colour]; int nrem = __builtin_popcount(f);
        double r2 = rng_uniform(rng); acc = 0;
        for (int i = 0; i < 8; i++) if (f & (128 >> i)) { acc += 1.0 / nrem; idx = i; if (acc > r2) break; }
    } else {
        int lc[40], li[40], m = 0; int64_t s = 0, mul = 1;
        for (int c = 0; c < 5; c++) {
            u8 f = (u8)bits[c];
            for (int i = 0; i < 8; i++) if (f & (128 >> i)) { lc[m] = c; li[m] = i; m++; }
            s += (int64_t)f * mul; mul *= 32;
        }
        int64_t x = 4594591LL * (seed + s);
        int64_t k = x % m; if (k < 0) k += m;          /* Python modulo */
        colour = lc[k]; idx = li[k];
    }
    bits[colour] = (i8)((u8)bits[colour] & ~(128 >> idx));
    cnt[colour] -= 1;
    write_card(out, tier, colour, idx);
    return 1;
}

/* SplendorLogicNumba.py:325-329 */
static void fill_new_card(i8* b, int tier, int index, int64_t seed, azo_rng* rng) {
    i8* slot = ROW(b, R_CARDS + 8 * tier + 2 * index);
    i8 card[2 * COLS];
    memset(slot, 0, 2 * COLS);
    if (get_deck_card(b, tier, seed, rng, card)) memcpy(slot, card, 2 * COLS);
}

/* SplendorLogicNumba.py:151-178 (init_game); randomness from the oracle's own RNG (numba's MT19937
 * stream is not reproducible outside numba -- parity tests use golden initial boards instead). */
void azo_init_game(i8* b, int n, uint64_t seed) {
    layout_t L = layout(n); azo_rng rng; rng_seed(&rng, seed);
    memset(b, 0, L.rows * COLS);
    int gems = n == 2 ? 4 : n == 3 ? 5 : 7;
    for (int c = 0; c < 5; c++) ROW(b, R_BANK)[c] = (i8)gems;
    ROW(b, R_BANK)[5] = 5;
    for (int t = 0; t < 3; t++)
        for (int c = 0; c < 5; c++) {
            ROW(b, R_DECK + 2 * t)[c] = (i8)DECK_SIZE[t];
            ROW(b, R_DECK + 2 * t + 1)[c] = (i8)(u8)(0xFF << (8 - DECK_SIZE[t]));
        }
    for (int t = 0; t < 3; t++) for (int i = 0; i < 4; i++) fill_new_card(b, t, i, 0, &rng);
    int perm[10]; for (int i = 0; i < 10; i++) perm[i] = i;
    for (int i = 0; i < L.nn; i++) {
        int j = i + (int)(rng_uniform(&rng) * (10 - i)); if (j > 9) j = 9;
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        uint32_t code = NOBLES[perm[i]]; i8* row = ROW(b, L.r_nobles + i);
        for (int k = 0; k < 5; k++) row[k] = (i8)((code >> (4 * k)) & 15);
        row[6] = (i8)((code >> 20) & 15);
    }
}

/* SplendorLogicNumba.py:303-304 */
int azo_get_round(const i8* b) { return (u8)ROW(b, R_BANK)[6]; }

/* SplendorLogicNumba.py:151-154 */
int azo_get_score(const i8* b, int n, int player) {
    layout_t L = layout(n); int s = ROW(b, L.r_pcards + player)[6];
    for (int i = 0; i < L.nn; i++) s += ROW(b, L.r_pnobles + L.nn * player + i)[6];
    return s;
}

static int can_buy(const i8* cost, const i8* pg, const i8* pc) {
    int missing = 0;
    for (int c = 0; c < 5; c++) { int d = (i8)(cost[c] - pg[c] - pc[c]); if (d > 0) missing += d; }
    return missing <= pg[5] && sum5(cost) != 0;
}

/* SplendorLogicNumba.py:180-188 with _valid_buy :359-368, _valid_reserve :375-380, _valid_buy_reserve
 * :402-412, _valid_get_gems(_identical) :422-434, _valid_give_gems(_identical) :443-453 */
void azo_valid_moves(const i8* b, int n, int player, u8* out) {
    layout_t L = layout(n);
    const i8* bank = ROW(b, R_BANK); const i8* pg = ROW(b, L.r_pgems + player); const i8* pc = ROW(b, L.r_pcards + player);
    const i8* res = ROW(b, L.r_pres + 6 * player);
    for (int i = 0; i < 12; i++) out[i] = (u8)can_buy(ROW(b, R_CARDS + 2 * i), pg, pc);
    int empty_slot = sum5(res + 5 * COLS) == 0;      /* gain row of the 3rd reserve slot */
    for (int i = 0; i < 12; i++) out[12 + i] = (u8)(sum5(ROW(b, R_CARDS + 2 * i)) != 0 && empty_slot);
    for (int t = 0; t < 3; t++) out[24 + t] = (u8)(sum5(ROW(b, R_DECK + 2 * t)) != 0 && empty_slot);
    for (int i = 0; i < 3; i++) out[27 + i] = (u8)can_buy(res + 2 * i * COLS, pg, pc);
    int have = sum7(pg);
    for (int i = 0; i < 25; i++) {
        int ok = 1, k = 0;
        for (int c = 0; c < 5; c++) if (GEMS3[i] >> c & 1) { k++; if (bank[c] - 1 < 0) ok = 0; }
        out[30 + i] = (u8)(ok && have + k <= 10);
    }
    for (int c = 0; c < 5; c++) out[55 + c] = (u8)(bank[c] >= 4 && have + 2 <= 10);
    for (int i = 0; i < 15; i++) {
        int ok = 1;
        for (int c = 0; c < 5; c++) if ((GEMS2[i] >> c & 1) && pg[c] - 1 < 0) ok = 0;
        out[60 + i] = (u8)ok;
    }
    for (int c = 0; c < 5; c++) out[75 + c] = (u8)(pg[c] >= 2);
    out[80] = 1;
}

/* SplendorLogicNumba.py:465-470 */
static void give_nobles(i8* b, const layout_t* L, int player) {
    i8* pc = ROW(b, L->r_pcards + player);
    for (int i = 0; i < L->nn; i++) {
        i8* noble = ROW(b, L->r_nobles + i);
        if (sum5(noble) <= 0) continue;
        int ok = 1; for (int c = 0; c < 5; c++) if (pc[c] < noble[c]) ok = 0;
        if (ok) { memcpy(ROW(b, L->r_pnobles + L->nn * player + i), noble, COLS); memset(noble, 0, COLS); }
    }
}

/* SplendorLogicNumba.py:331-357 (_buy_card) */
static void buy_card(i8* b, const layout_t* L, const i8* card0, const i8* card1, int player) {
    i8* bank = ROW(b, R_BANK); i8* pg = ROW(b, L->r_pgems + player); i8* pc = ROW(b, L->r_pcards + player);
    int missing = 0; i8 paid[5];
    for (int c = 0; c < 5; c++) {
        int d = (i8)(card0[c] - pg[c] - pc[c]); if (d > 0) missing += d;
        int need = (i8)(card0[c] - pc[c]); if (need < 0) need = 0;
        paid[c] = (i8)(need < pg[c] ? need : pg[c]);
    }
    for (int c = 0; c < 5; c++) { pg[c] -= paid[c]; bank[c] += paid[c]; }
    pg[5] = (i8)(pg[5] - missing); bank[5] = (i8)(bank[5] + missing);
    for (int c = 0; c < COLS; c++) pc[c] += card1[c];
    give_nobles(b, L, player);
}

/* SplendorLogicNumba.py:190-205 (make_move) and the helpers it dispatches to:
 * _buy :370-373, _reserve :382-400, _buy_reserve :414-420, _get_gems :436-441 area, _give_gems :455-463 */
int azo_make_move(i8* b, int n, int move, int player, int64_t seed, azo_rng* rng) {
    layout_t L = layout(n);
    i8* bank = ROW(b, R_BANK); i8* pg = ROW(b, L.r_pgems + player);
    i8 c0[COLS], c1[COLS];
    if (move < 12) {
        i8* card = ROW(b, R_CARDS + 2 * move);
        memcpy(c0, card, COLS); memcpy(c1, card + COLS, COLS);
        buy_card(b, &L, c0, c1, player);
        fill_new_card(b, move / 4, move % 4, seed, rng);
    } else if (move < 27) {
        int i = move - 12; i8* res = ROW(b, L.r_pres + 6 * player); int slot = -1;
        for (int k = 0; k < 3; k++) if (sum5(res + 2 * k * COLS) == 0) { slot = k; break; }
        if (slot >= 0) {
            if (i < 12) {
                memcpy(res + 2 * slot * COLS, ROW(b, R_CARDS + 2 * i), 2 * COLS);
                fill_new_card(b, i / 4, i % 4, seed, rng);
            } else {
                i8 card[2 * COLS];
                if (get_deck_card(b, i - 12, seed, rng, card)) memcpy(res + 2 * slot * COLS, card, 2 * COLS);
            }
        }
        if (bank[5] > 0 && sum7(pg) <= 9) { pg[5] += 1; bank[5] -= 1; }
    } else if (move < 30) {
        int i = move - 27; i8* res = ROW(b, L.r_pres + 6 * player); i8* card = res + 2 * i * COLS;
        memcpy(c0, card, COLS); memcpy(c1, card + COLS, COLS);
        buy_card(b, &L, c0, c1, player);
        if (i < 2) memmove(card, card + 2 * COLS, (size_t)(2 - i) * 2 * COLS);
        memset(res + 4 * COLS, 0, 2 * COLS);
    } else if (move < 60) {
        int i = move - 30;
        if (i < 25) { for (int c = 0; c < 5; c++) if (GEMS3[i] >> c & 1) { bank[c] -= 1; pg[c] += 1; } }
        else { bank[i - 25] -= 2; pg[i - 25] += 2; }
    } else if (move < 80) {
        int i = move - 60;
        if (i < 15) { for (int c = 0; c < 5; c++) if (GEMS2[i] >> c & 1) { bank[c] += 1; pg[c] -= 1; } }
        else { bank[i - 15] += 2; pg[i - 15] -= 2; }
    }
    bank[6] += 1;
    return (player + 1) % n;
}

/* C-callable wrapper with an explicit RNG seed for the seed==0 (true random) case. */
int azo_next_state(i8* b, int n, int move, int player, int64_t seed, uint64_t rng_seed_) {
    azo_rng r; rng_seed(&r, rng_seed_);
    return azo_make_move(b, n, move, player, seed, &r);
}

/* SplendorLogicNumba.py:221-240 (check_end_game) */
void azo_check_end_game(const i8* b, int n, float* out) {
    layout_t L = layout(n);
    for (int p = 0; p < n; p++) out[p] = 0.f;
    int round = azo_get_round(b);
    if (round % n != 0) return;
    float scores[MAXP], mx = -1e30f;
    for (int p = 0; p < n; p++) { scores[p] = (float)azo_get_score(b, n, p); if (scores[p] > mx) mx = scores[p]; }
    if (!(mx >= 15.f || round >= 62 * n)) return;
    int winners = 0; for (int p = 0; p < n; p++) winners += scores[p] == mx;
    int several = winners > 1;
    if (several) {
        for (int p = 0; p < n; p++) {
            int cards = sum5(ROW(b, L.r_pcards + p));
            scores[p] = (float)((double)scores[p] - (double)cards / 100.);
        }
        mx = -1e30f; for (int p = 0; p < n; p++) if (scores[p] > mx) mx = scores[p];
        winners = 0; for (int p = 0; p < n; p++) winners += scores[p] == mx;
        several = winners > 1;
    }
    for (int p = 0; p < n; p++) out[p] = scores[p] == mx ? (several ? 0.01f : 1.f) : -1.f;
}

/* SplendorLogicNumba.py:244-253 (swap_players): new[i] = old[(i+shift) % size] on 4 row groups */
static void roll_rows(i8* base, int size, int shift) {
    i8 tmp[MAXROWS * COLS];
    memcpy(tmp, base, (size_t)size * COLS);
    for (int i = 0; i < size; i++) memcpy(base + i * COLS, tmp + ((i + shift) % size) * COLS, COLS);
}
void azo_swap_players(i8* b, int n, int nb_swaps) {
    layout_t L = layout(n);
    roll_rows(ROW(b, L.r_pgems), n, nb_swaps);
    roll_rows(ROW(b, L.r_pnobles), n * L.nn, L.nn * nb_swaps);
    roll_rows(ROW(b, L.r_pcards), n, nb_swaps);
    roll_rows(ROW(b, L.r_pres), 6 * n, 6 * nb_swaps);
}

/* SplendorLogicNumba.py:255-301 (get_symmetries). Writes up to 1+9+2n triples, returns the count. */
static const int CARD_PERM[3][4] = {{1, 3, 0, 2}, {2, 0, 3, 1}, {3, 2, 1, 0}};
static const int RES_PERM[4][2][3] = {{{-1, -1, -1}, {-1, -1, -1}}, {{-1, -1, -1}, {-1, -1, -1}},
                                      {{1, 0, 2}, {-1, -1, -1}}, {{1, 2, 0}, {2, 0, 1}}};
int azo_symmetries(const i8* b, int n, const float* pi, const u8* valids, i8* ob, float* opi, u8* ov) {
    layout_t L = layout(n); int S = L.rows * COLS, k = 0;
    memcpy(ob, b, S); memcpy(opi, pi, NA * sizeof(float)); memcpy(ov, valids, NA); k++;
    for (int t = 0; t < 3; t++)
        for (int q = 0; q < 3; q++) {
            i8* o = ob + (size_t)k * S; float* p = opi + (size_t)k * NA; u8* v = ov + (size_t)k * NA;
            memcpy(o, b, S); memcpy(p, pi, NA * sizeof(float)); memcpy(v, valids, NA);
            for (int i = 0; i < 4; i++) {
                int src = CARD_PERM[q][i];
                memcpy(ROW(o, R_CARDS + 8 * t + 2 * i), ROW(b, R_CARDS + 8 * t + 2 * src), 2 * COLS);
                p[4 * t + i] = pi[4 * t + src]; p[12 + 4 * t + i] = pi[12 + 4 * t + src];
                v[4 * t + i] = valids[4 * t + src]; v[12 + 4 * t + i] = valids[12 + 4 * t + src];
            }
            k++;
        }
    for (int pl = 0; pl < n; pl++) {
        const i8* res = ROW(b, L.r_pres + 6 * pl);
        int nres = 3; for (int c = 0; c < 3; c++) if (sum5(res + 2 * c * COLS) == 0) { nres = c; break; }
        for (int q = 0; q < 2; q++) {
            if (RES_PERM[nres][q][0] < 0) continue;
            i8* o = ob + (size_t)k * S; float* p = opi + (size_t)k * NA; u8* v = ov + (size_t)k * NA;
            memcpy(o, b, S); memcpy(p, pi, NA * sizeof(float)); memcpy(v, valids, NA);
            for (int i = 0; i < 3; i++) {
                int src = RES_PERM[nres][q][i];
                memcpy(ROW(o, L.r_pres + 6 * pl + 2 * i), res + 2 * src * COLS, 2 * COLS);
                if (pl == 0) { p[27 + i] = pi[27 + src]; v[27 + i] = valids[27 + src]; }
            }
            k++;
        }
    }
    return k;
}

/* ================================================================ Abalone (Belgian daisy) ===============
 * abalone/AbaloneLogicNumba.py with the shipped constants INITIAL_LAYOUT = 1, ENABLE_DYNAMIC_KOMI = False (:5-6).
 * State int8[9][9][4] (cell (r,q) at bytes 4(9r+q)..+3: mover's marble, opponent's marble, board mask, misc);
 * misc[0,0] = byte 3 mover's score, misc[0,1] = byte 7 opponent's score, misc[0,2] = byte 11 round.
 * Action = 378 r + 42 q + plane (:62-84). */
#define ABA_S 324
#define ABA_A 3402
static const int ABA_DR[6] = {0, 1, 1, 0, -1, -1}, ABA_DQ[6] = {1, 0, -1, -1, 0, 1};          /* DIRECTIONS :53-60 */
static int aba_at(const i8* b, int r, int q, int ch) { return b[4 * (9 * r + q) + ch]; }
static int aba_on(const i8* b, int r, int q) { return r >= 0 && r < 9 && q >= 0 && q < 9 && aba_at(b, r, q, 2) == 1; }   /* :86-90 */
static int aba_enc(int r, int q, int size, int axis, int d) { int plane = size == 1 ? d : (size == 2 ? 6 + axis * 6 + d : 24 + axis * 6 + d); return r * 378 + q * 42 + plane; }
static void aba_dec(int a, int* r, int* q, int* size, int* axis, int* d) {
    int plane = a % 42; *q = (a / 42) % 9; *r = a / 378; *d = plane % 6;
    if (plane < 6) { *size = 1; *axis = 0; } else if (plane < 24) { *size = 2; *axis = (plane - 6) / 6; } else { *size = 3; *axis = (plane - 24) / 6; }
}
int azo_aba_get_round(const i8* b) { return b[11]; }
int azo_aba_get_score(const i8* b, int player) { return player == 0 ? b[3] : b[7]; }
/* valid_moves :254-331 */
void azo_aba_valid_moves(const i8* b, int player, u8* out) {
    memset(out, 0, ABA_A);
    int opp = 1 - player;
    for (int r = 0; r < 9; r++) for (int q = 0; q < 9; q++) {
        if (aba_at(b, r, q, player) == 0) continue;
        for (int d = 0; d < 6; d++) {
            int nr = r + ABA_DR[d], nq = q + ABA_DQ[d];
            if (aba_on(b, nr, nq) && aba_at(b, nr, nq, player) == 0 && aba_at(b, nr, nq, opp) == 0) out[aba_enc(r, q, 1, 0, d)] = 1;
        }
        for (int axis = 0; axis < 3; axis++) {
            int r1 = r + ABA_DR[axis], q1 = q + ABA_DQ[axis];
            if (!aba_on(b, r1, q1) || aba_at(b, r1, q1, player) == 0) continue;
            int r2 = r1 + ABA_DR[axis], q2 = q1 + ABA_DQ[axis];
            int max_size = (aba_on(b, r2, q2) && aba_at(b, r2, q2, player) == 1) ? 3 : 2;
            for (int size = 2; size <= max_size; size++)
                for (int d = 0; d < 6; d++) {
                    int inl = d == axis || d == (axis + 3) % 6;
                    if (!inl) {
                        int ok = 1;
                        for (int i = 0; i < size; i++) {
                            int tr = r + i * ABA_DR[axis] + ABA_DR[d], tq = q + i * ABA_DQ[axis] + ABA_DQ[d];
                            if (!aba_on(b, tr, tq) || aba_at(b, tr, tq, player) == 1 || aba_at(b, tr, tq, opp) == 1) { ok = 0; break; }
                        }
                        if (ok) out[aba_enc(r, q, size, axis, d)] = 1;
                    } else {
                        int fr = d == axis ? r + (size - 1) * ABA_DR[axis] : r, fq = d == axis ? q + (size - 1) * ABA_DQ[axis] : q;
                        int tr = fr + ABA_DR[d], tq = fq + ABA_DQ[d];
                        if (!aba_on(b, tr, tq)) continue;
                        if (aba_at(b, tr, tq, player) == 1) continue;
                        if (aba_at(b, tr, tq, opp) == 0) { out[aba_enc(r, q, size, axis, d)] = 1; continue; }
                        int cnt = 0, cr = tr, cq = tq, ok = 0;
                        for (;;) {
                            if (!aba_on(b, cr, cq)) { if (cnt > 0) ok = 1; break; }
                            if (aba_at(b, cr, cq, opp) == 1) { cnt++; if (cnt >= size) break; cr += ABA_DR[d]; cq += ABA_DQ[d]; }
                            else if (aba_at(b, cr, cq, player) == 1) break;
                            else { ok = 1; break; }
                        }
                        if (ok) out[aba_enc(r, q, size, axis, d)] = 1;
                    }
                }
        }
    }
}
/* make_move :333-374 */
int azo_aba_make_move(i8* b, int move, int player) {
    int r, q, size, axis, d; aba_dec(move, &r, &q, &size, &axis, &d);
    int inl = d == axis || d == (axis + 3) % 6, opp = 1 - player;
    if (size == 1 || !inl) {
        for (int i = 0; i < size; i++) {
            int cr = size > 1 ? r + i * ABA_DR[axis] : r, cq = size > 1 ? q + i * ABA_DQ[axis] : q;
            b[4 * (9 * cr + cq) + player] = 0; b[4 * (9 * (cr + ABA_DR[d]) + cq + ABA_DQ[d]) + player] = 1;
        }
    } else {
        int fr, fq, br, bq;
        if (d == axis) { fr = r + (size - 1) * ABA_DR[axis]; fq = q + (size - 1) * ABA_DQ[axis]; br = r; bq = q; }
        else { fr = r; fq = q; br = r + (size - 1) * ABA_DR[axis]; bq = q + (size - 1) * ABA_DQ[axis]; }
        int tr = fr + ABA_DR[d], tq = fq + ABA_DQ[d];
        if (aba_on(b, tr, tq) && aba_at(b, tr, tq, opp) == 1) {
            int cr = tr, cq = tq;
            while (aba_on(b, cr, cq) && aba_at(b, cr, cq, opp) == 1) { cr += ABA_DR[d]; cq += ABA_DQ[d]; }
            b[4 * (9 * tr + tq) + opp] = 0;
            if (aba_on(b, cr, cq)) b[4 * (9 * cr + cq) + opp] = 1; else b[4 * player + 3] = (i8)(b[4 * player + 3] + 1);
        }
        b[4 * (9 * br + bq) + player] = 0; b[4 * (9 * tr + tq) + player] = 1;
    }
    b[11] = (i8)(b[11] + 1);
    return 1 - player;
}
/* check_end_game :376-392 */
void azo_aba_check_end_game(const i8* b, float* out) {
    out[0] = out[1] = 0.f;
    if (b[3] >= 6) { out[0] = 1.f; out[1] = -1.f; return; }
    if (b[7] >= 6) { out[0] = -1.f; out[1] = 1.f; return; }
    if (b[11] >= 127) {
        if (b[3] > b[7]) { out[0] = 1.f; out[1] = -1.f; } else if (b[7] > b[3]) { out[0] = -1.f; out[1] = 1.f; } else { out[0] = out[1] = 0.001f; }
    }
}
/* swap_players :394-406 */
void azo_aba_swap_players(i8* b, int nb_swaps) {
    if (nb_swaps % 2 != 1) return;
    for (int c = 0; c < 81; c++) { i8 t = b[4 * c]; b[4 * c] = b[4 * c + 1]; b[4 * c + 1] = t; }
    i8 t = b[3]; b[3] = b[7]; b[7] = t;
}
/* init_game :167-252, Belgian daisy */
void azo_aba_init_game(i8* b) {
    memset(b, 0, ABA_S);
    for (int r = 0; r < 9; r++) for (int q = 0; q < 9; q++) if (r + q >= 4 && r + q <= 12) b[4 * (9 * r + q) + 2] = 1;
    static const int OPP[6][3] = {{0, 4, 6}, {1, 3, 6}, {2, 3, 5}, {6, 4, 6}, {7, 3, 6}, {8, 3, 5}}, ME[6][3] = {{0, 7, 9}, {1, 6, 9}, {2, 6, 8}, {6, 1, 3}, {7, 0, 3}, {8, 0, 2}};
    for (int i = 0; i < 6; i++) { for (int q = OPP[i][1]; q < OPP[i][2]; q++) b[4 * (9 * OPP[i][0] + q) + 1] = 1; for (int q = ME[i][1]; q < ME[i][2]; q++) b[4 * (9 * ME[i][0] + q)] = 1; }
}
static void aba_xform(int* r, int* q, int rot, int flip) {
    if (flip) *q = 12 - *r - *q;
    for (int i = 0; i < rot; i++) { int nr = *q + *r - 4, nq = 8 - *r; *r = nr; *q = nq; }
}
/* _build_action_symmetries :95-148, one entry */
static int aba_map_action(int a, int rot, int flip) {
    static const int FLIPD[6] = {3, 2, 1, 0, 5, 4};
    int r, q, size, axis, d; aba_dec(a, &r, &q, &size, &axis, &d);
    int mr[3], mq[3];
    for (int i = 0; i < size; i++) { mr[i] = r + i * ABA_DR[axis]; mq[i] = q + i * ABA_DQ[axis]; aba_xform(&mr[i], &mq[i], rot, flip); }
    int mi = 0; for (int i = 1; i < size; i++) if (mr[i] < mr[mi] || (mr[i] == mr[mi] && mq[i] < mq[mi])) mi = i;
    int na = 0;
    if (size > 1) { int oi = mi == 0 ? 1 : 0, dr = mr[oi] - mr[mi], dq = mq[oi] - mq[mi]; if (dr == 0 && dq > 0) na = 0; else if (dr > 0 && dq == 0) na = 1; else if (dr > 0 && dq < 0) na = 2; }
    int nd = d; if (flip) nd = FLIPD[nd]; nd = (nd + rot) % 6;
    return aba_enc(mr[mi], mq[mi], size, na, nd);
}
/* get_symmetries :408-441: 12 = 6 rotations x 2 reflections, k = 2*rot + flip */
int azo_aba_symmetries(const i8* b, const float* pi, const u8* valids, i8* ob, float* opi, u8* ov) {
    for (int k = 0; k < 12; k++) {
        int rot = k / 2, flip = k % 2;
        i8* o = ob + k * ABA_S; float* op = opi + (size_t)k * ABA_A; u8* om = ov + (size_t)k * ABA_A;
        memset(o, 0, ABA_S); memset(op, 0, sizeof(float) * ABA_A); memset(om, 0, ABA_A);
        for (int r = 0; r < 9; r++) for (int q = 0; q < 9; q++) if (aba_at(b, r, q, 2) == 1) {
            int nr = r, nq = q; aba_xform(&nr, &nq, rot, flip);
            for (int ch = 0; ch < 3; ch++) o[4 * (9 * nr + nq) + ch] = b[4 * (9 * r + q) + ch];
        }
        for (int c = 0; c < 81; c++) o[4 * c + 3] = b[4 * c + 3];
        for (int a = 0; a < ABA_A; a++) if (valids[a]) { int m = aba_map_action(a, rot, flip); op[m] = pi[a]; om[m] = valids[a]; }
    }
    return 12;
}

/* ================================================================ Santorini, no gods ====================
 * santorini/SantoriniLogicNumba.py built with NB_GODS = 1 (santorini/SantoriniConstants.py:19): state int8[5][5][3]
 * (cell c = 5y+x at bytes 3c..3c+2: worker, level, gods_power), gods_power.flat[i] = byte 3i+2; flat[0], flat[1] = 64
 * (NO_GOD owned), flat[2] = round counter; flat[8..10] are the Pan / Athena slots the no-god code still reads.
 * Action = 81*worker + 9*move_direction + build_direction (SantoriniConstants.py:23-34). */
#define SAN_S 75
#define SAN_A 162
#define SAN_PAN 8
#define SAN_ATHENA 9
static int san_wk(const i8* b, int c) { return b[3 * c]; }
static int san_lv(const i8* b, int c) { return b[3 * c + 1]; }
static int san_gp(const i8* b, int i) { return b[3 * i + 2]; }
int azo_sant_get_round(const i8* b) { return san_gp(b, 2); }                                  /* :655-656 */
int azo_sant_get_score(const i8* b, int player) {                                             /* :87-101 */
    int best = 0;
    for (int c = 0; c < 25; c++) { int w = san_wk(b, c); if ((player == 0 ? w > 0 : w < 0) && san_lv(b, c) > best) best = san_lv(b, c); }
    return best;
}
static int san_find(const i8* b, int id) { for (int c = 0; c < 25; c++) if (san_wk(b, c) == id) return c; return -1; }   /* :667-672 */
/* _able_to_move_worker_to :675-701 (no swap / push without gods) */
static int san_can_move(const i8* b, int old, int ny, int nx, int no_climb) {
    if (ny < 0 || ny >= 5 || nx < 0 || nx >= 5) return 0;
    int nc = 5 * ny + nx;
    if (san_wk(b, nc) != 0) return 0;
    if (san_lv(b, nc) > 3) return 0;
    if (san_lv(b, nc) > san_lv(b, old) + (no_climb ? 0 : 1)) return 0;
    return 1;
}
/* _able_to_build :719-729 */
static int san_can_build(const i8* b, int y, int x, int ignore) {
    if (y < 0 || y >= 5 || x < 0 || x >= 5) return 0;
    int w = san_wk(b, 5 * y + x);
    if (!(w == 0 || w == ignore)) return 0;
    return san_lv(b, 5 * y + x) < 4;
}
/* valid_moves, NO_GOD branch :135-151 */
void azo_sant_valid_moves(const i8* b, int player, u8* out) {
    memset(out, 0, SAN_A);
    if (san_gp(b, player) <= 0) return;
    int no_climb = san_gp(b, SAN_ATHENA + (1 - player)) > 64;
    for (int worker = 0; worker < 2; worker++) {
        int wid = (worker + 1) * (player == 0 ? 1 : -1), old = san_find(b, wid);
        if (old < 0) continue;
        for (int md = 0; md < 9; md++) {
            if (md == 4) continue;
            int ny = old / 5 + md / 3 - 1, nx = old % 5 + md % 3 - 1;
            if (!san_can_move(b, old, ny, nx, no_climb)) continue;
            for (int bd = 0; bd < 9; bd++) {
                if (bd == 4) continue;
                if (!san_can_build(b, ny + bd / 3 - 1, nx + bd % 3 - 1, wid)) continue;
                out[81 * worker + 9 * md + bd] = 1;
            }
        }
    }
}
/* make_move :434-550, power == NO_GOD; returns the next player */
int azo_sant_make_move(i8* b, int move, int player) {
    int worker = move / 81, md = (move % 81) / 9, bd = move % 9;
    int wid = (worker + 1) * (player == 0 ? 1 : -1), old = san_find(b, wid);
    if (old >= 0) {
        int old_level = san_lv(b, old);
        int nc = 5 * (old / 5 + md / 3 - 1) + (old % 5 + md % 3 - 1);
        b[3 * old] = 0; b[3 * nc] = (i8)wid;
        if (bd != 4) { int bc = 5 * (nc / 5 + bd / 3 - 1) + (nc % 5 + bd % 3 - 1); b[3 * bc + 1] = (i8)(b[3 * bc + 1] + 1); }
        int new_level = san_lv(b, nc);
        if (san_gp(b, SAN_PAN + player) > 0) { if (new_level <= old_level - 2) b[3 * (SAN_PAN + player) + 2] = 65; }
        else if (san_gp(b, SAN_ATHENA + player) > 0) b[3 * (SAN_ATHENA + player) + 2] = (i8)(64 + (new_level > old_level ? 1 : 0));
        else { int v = san_gp(b, player); b[3 * player + 2] = (i8)(v < 64 ? v : 64); }
    }
    if (san_gp(b, 2) < 127) b[3 * 2 + 2] = (i8)(san_gp(b, 2) + 1);
    return 1 - player;
}
/* check_end_game :552-565 */
void azo_sant_check_end_game(const i8* b, int next_player, float* out) {
    out[0] = out[1] = 0.f;
    if (azo_sant_get_score(b, 0) == 3 || san_gp(b, SAN_PAN) > 64) { out[0] = 1.f; out[1] = -1.f; return; }
    if (azo_sant_get_score(b, 1) == 3 || san_gp(b, SAN_PAN + 1) > 64) { out[0] = -1.f; out[1] = 1.f; return; }
    u8 v[SAN_A]; azo_sant_valid_moves(b, next_player, v);
    int any = 0; for (int a = 0; a < SAN_A; a++) any |= v[a];
    if (!any) { out[next_player] = -1.f; out[1 - next_player] = 1.f; }
}
/* swap_players :567-576 */
void azo_sant_swap_players(i8* b, int nb_swaps) {
    if (nb_swaps != 1) return;
    for (int c = 0; c < 25; c++) b[3 * c] = (i8)(-b[3 * c]);
    i8 t = b[2]; b[2] = b[5]; b[5] = t;
}
/* init_game :103-120, INIT_METHOD = 1 (random worker squares; oracle-own RNG, the reference uses numba's MT19937) */
void azo_sant_init_game(i8* b, uint64_t seed) {
    azo_rng r; rng_seed(&r, seed);
    memset(b, 0, SAN_S);
    const int ids[4] = {1, -1, 2, -2}; uint32_t used = 0;
    for (int i = 0; i < 4; i++) {
        int k = (int)(rng_uniform(&r) * (25 - i)), c = 0;
        for (int j = 0; j < 25; j++) if (!(used >> j & 1)) { if (k == 0) { c = j; break; } k--; }
        used |= 1u << c; b[3 * c] = (i8)ids[i];
    }
    b[2] = 64; b[5] = 64;
}
/* get_symmetries :578-653: identity, rot90 x1..3, flipLR, flipUD, swap own workers, swap opponent workers */
static const int SAN_ROT[9] = {6, 3, 0, 7, 4, 1, 8, 5, 2}, SAN_FLR[9] = {2, 1, 0, 5, 4, 3, 8, 7, 6}, SAN_FUD[9] = {6, 7, 8, 3, 4, 5, 0, 1, 2};
int azo_sant_symmetries(const i8* b, const float* pi, const u8* valids, i8* ob, float* opi, u8* ov) {
    for (int k = 0; k < 8; k++) {
        i8* o = ob + k * SAN_S; float* op = opi + k * SAN_A; u8* om = ov + k * SAN_A;
        for (int c = 0; c < 25; c++) {
            int y = c / 5, x = c % 5, sy = y, sx = x;
            if (k >= 1 && k <= 3) for (int i = 0; i < k; i++) { int ty = sx, tx = 4 - sy; sy = ty; sx = tx; }      /* np.rot90 */
            else if (k == 4) sx = 4 - x;
            else if (k == 5) sy = 4 - y;
            int sc = 5 * sy + sx, w = b[3 * sc];
            if (k == 6 && w > 0) w = 3 - w;
            if (k == 7 && w < 0) w = -3 - w;
            o[3 * c] = (i8)w; o[3 * c + 1] = b[3 * sc + 1]; o[3 * c + 2] = b[3 * c + 2];
        }
        for (int a = 0; a < SAN_A; a++) {
            int worker = a / 81, md = (a % 81) / 9, bd = a % 9, dst;
            if (k == 6) dst = (1 - worker) * 81 + md * 9 + bd;
            else {
                int m2 = md, b2 = bd;
                if (k >= 1 && k <= 3) for (int i = 0; i < k; i++) { m2 = SAN_ROT[m2]; b2 = SAN_ROT[b2]; }
                else if (k == 4) { m2 = SAN_FLR[md]; b2 = SAN_FLR[bd]; }
                else if (k == 5) { m2 = SAN_FUD[md]; b2 = SAN_FUD[bd]; }
                dst = worker * 81 + m2 * 9 + b2;
            }
            op[dst] = pi[a]; om[dst] = valids[a];
        }
    }
    return 8;
}

/* ---------------------------------------------------------------- Azul (2 players) ------- */
/* azul/AzulLogicNumba.py (rules) and azul/AzulLogic.py:4-126 (the 120 factory permutations). ROUND-2 GROUNDWORK: rules + the game-generic MCTS (game id 3), pinned by
 * tests/golden/azul_kat.npz and azul_mcts.npz; no CUDA plugin or net for this game yet.
 * State int8[23][6] (:6-24): row 0 scores (P0, P1, round), 1 bag, 2 discards, 3 centre (+ first-player token in column 5), 4-8 the five
 * factories, 9-10 pattern-line colours of P0/P1 (-1 = empty; column 5 = holds the token), 11-12 tiles per pattern line (column 5 = floor),
 * 13-17 / 18-22 the walls. Action = 30 source + 6 colour + line, source 0 = centre, line 5 = floor (:27-48). */
#define AZU_S 138
#define AZU_A 180
#define AZ(b, r, c) (b)[6 * (r) + (c)]
int azo_azul_get_round(const i8* b) { return AZ(b, 0, 2); }                                      /* get_round :333-334 */
int azo_azul_get_score(const i8* b, int player) { return AZ(b, 0, player); }                     /* get_score :83-84 */

void azo_azul_valid_moves(const i8* b, int player, u8* out) {                                    /* valid_moves :97-124 */
    const i8* pc = &AZ(b, 9 + player, 0); const i8* pn = &AZ(b, 11 + player, 0);
    for (int src = 0; src < 6; src++)
        for (int colour = 0; colour < 5; colour++) {
            const int avail = src == 0 ? AZ(b, 3, colour) != 0 : AZ(b, 3 + src, colour) > 0;     /* centre: astype(bool); factory: > 0 */
            for (int line = 0; line < 6; line++) {
                const int line_free = line == 5 ? 1 : pc[line] == -1;
                const int wall_free = line == 5 ? 1 : AZ(b, 13 + 5 * player + line, (colour + line) % 5) == 0;
                const int partial = pc[line] == colour && pn[line] < line + 1;
                out[src * 30 + colour * 6 + line] = (u8)(avail && ((line_free && wall_free) || partial));
            }
        }
}

/* select_tiles_from_bag :258-269. seed != 0: the reference's deterministic draw; seed == 0 (true random there): the oracle's own RNG. */
static void azul_draw(i8* b, int num, int64_t seed, azo_rng* rng, i8* result) {
    for (int t = 0; t < num; t++) {
        int total = 0; for (int c = 0; c < 6; c++) total += AZ(b, 1, c);
        if (total <= 0) return;                                                                   /* the reference would divide by zero here */
        int64_t tile;
        if (seed == 0) tile = (int64_t)(rng_uniform(rng) * total);
        else {
            int64_t h = 0; for (int c = 0; c < 5; c++) h += (int64_t)AZ(b, 1, c) << c;
            tile = (4594591 * (seed + h)) % total; if (tile < 0) tile += total;                   /* Python modulo */
        }
        int idx = 0, cum = 0;
        for (; idx < 5; idx++) { cum += AZ(b, 1, idx); if (cum > tile) break; }                   /* searchsorted(cumsum, tile, side='right') */
        if (idx > 4) idx = 4;
        result[idx]++; AZ(b, 1, idx)--;
    }
}
static int azul_setup_new_round(i8* b, int64_t seed, azo_rng* rng) {                             /* setup_new_round :238-256 */
    for (int i = 0; i < 5; i++) {
        int total = 0; for (int c = 0; c < 6; c++) total += AZ(b, 1, c);
        i8 res[6] = {0, 0, 0, 0, 0, 0};
        if (total < 4) {
            for (int c = 0; c < 6; c++) { AZ(b, 4 + i, c) = AZ(b, 1, c); AZ(b, 1, c) = AZ(b, 2, c); AZ(b, 2, c) = 0; }
            azul_draw(b, 4 - total, seed, rng, res);
            for (int c = 0; c < 6; c++) AZ(b, 4 + i, c) = (i8)(AZ(b, 4 + i, c) + res[c]);
        } else {
            azul_draw(b, 4, seed, rng, res);
            for (int c = 0; c < 6; c++) AZ(b, 4 + i, c) = res[c];
        }
    }
    int next_player;
    if (AZ(b, 10, 5) == 1) { next_player = 1; AZ(b, 10, 5) = 0; } else { next_player = 0; AZ(b, 9, 5) = 0; }
    AZ(b, 0, 2) = (i8)(AZ(b, 0, 2) + 1);
    AZ(b, 3, 5) = 1;
    return next_player;
}
static int azul_run(const i8* w, int r, int c, int along_row) {                                   /* count_consecutive_ones :199-210 */
    int count = 1;
    if (along_row) { for (int k = c - 1; k >= 0 && w[6 * r + k] == 1; k--) count++; for (int k = c + 1; k < 5 && w[6 * r + k] == 1; k++) count++; }
    else { for (int k = r - 1; k >= 0 && w[6 * k + c] == 1; k--) count++; for (int k = r + 1; k < 5 && w[6 * k + c] == 1; k++) count++; }
    return count;
}
static int azul_score_change(i8* w, int r, int c) {                                              /* score_change :212-220; w = the player's 5 wall rows */
    w[6 * r + c] = 1;
    const int row_adj = (c > 0 && w[6 * r + c - 1] == 1) || (c < 4 && w[6 * r + c + 1] == 1);
    const int col_adj = (r > 0 && w[6 * (r - 1) + c] == 1) || (r < 4 && w[6 * (r + 1) + c] == 1);
    if (!row_adj && !col_adj) return 1;
    return (row_adj ? azul_run(w, r, c, 1) : 0) + (col_adj ? azul_run(w, r, c, 0) : 0);
}
static void azul_score_round(i8* b) {                                                            /* score_round :161-181 */
    static const int FLOOR_PENALTY[8] = {0, 1, 2, 4, 6, 8, 11, 14};
    int pl[10], rw[10], col[10], n = 0;
    for (int p = 0; p < 2; p++) for (int r = 0; r < 5; r++) if (AZ(b, 11 + p, r) == r + 1) { pl[n] = p; rw[n] = r; col[n] = AZ(b, 9 + p, r); n++; }   /* np.where order */
    for (int i = 0; i < n; i++) {
        const int c = ((col[i] + rw[i]) % 5 + 5) % 5;
        AZ(b, 0, pl[i]) = (i8)(AZ(b, 0, pl[i]) + azul_score_change(&AZ(b, 13 + 5 * pl[i], 0), rw[i], c));
        AZ(b, 13 + 5 * pl[i] + rw[i], c) = 1;
    }
    for (int i = 0; i < n; i++) AZ(b, 2, (col[i] % 6 + 6) % 6) = (i8)(AZ(b, 2, (col[i] % 6 + 6) % 6) + rw[i]);
    for (int i = 0; i < n; i++) { AZ(b, 11 + pl[i], rw[i]) = 0; AZ(b, 9 + pl[i], rw[i]) = -1; }
    for (int p = 0; p < 2; p++) {
        int fl = AZ(b, 11 + p, 5); if (fl > 7) fl = 7; if (fl < 0) fl = 0;
        const int sc = AZ(b, 0, p) - FLOOR_PENALTY[fl];
        AZ(b, 0, p) = (i8)(sc > 0 ? sc : 0);
        AZ(b, 11 + p, 5) = 0;
    }
}
static int azul_game_over(const i8* b) {                                                         /* check_game_over :153-159 */
    for (int r = 0; r < 10; r++) { int all = 1; for (int c = 0; c < 5; c++) all &= AZ(b, 13 + r, c) == 1; if (all) return 1; }
    return 0;
}
static void azul_score_bonuses(i8* b) {                                                          /* score_bonuses :183-197 */
    for (int p = 0; p < 2; p++) {
        const i8* w = &AZ(b, 13 + 5 * p, 0); int add = 0;
        for (int r = 0; r < 5; r++) { int all = 1; for (int c = 0; c < 5; c++) all &= w[6 * r + c] == 1; if (all) add += 2; }
        for (int c = 0; c < 5; c++) { int all = 1; for (int r = 0; r < 5; r++) all &= w[6 * r + c] == 1; if (all) add += 7; }
        for (int i = 0; i < 5; i++) { int all = 1; for (int j = 0; j < 5; j++) all &= w[6 * j + (j + i) % 5] == 1; if (all) add += 10; }
        AZ(b, 0, p) = (i8)(AZ(b, 0, p) + add);
    }
}
int azo_azul_make_move(i8* b, int move, int player, int64_t seed, uint64_t rng_seed_) {         /* make_move :126-151 */
    azo_rng rng; rng_seed(&rng, rng_seed_);
    i8* src = move < 30 ? &AZ(b, 3, 0) : &AZ(b, 4 + (move - 30) / 30, 0);
    const int colour = (move % 30) / 6, line = move % 6, num = src[colour];
    int to_floor;
    if (line == 5) to_floor = num;
    else {
        const int on_line = AZ(b, 11 + player, line);
        const int to_line = line + 1 - on_line < num ? line + 1 - on_line : num;
        to_floor = num - to_line;
        AZ(b, 11 + player, line) = (i8)(on_line + to_line);
        AZ(b, 9 + player, line) = (i8)colour;
    }
    AZ(b, 11 + player, 5) = (i8)(AZ(b, 11 + player, 5) + to_floor);
    AZ(b, 2, colour) = (i8)(AZ(b, 2, colour) + to_floor);
    src[colour] = 0;
    if (move < 30) {
        if (src[5] == 1) { AZ(b, 11 + player, 5) = (i8)(AZ(b, 11 + player, 5) + 1); AZ(b, 9 + player, 5) = 1; src[5] = 0; }
    } else {
        for (int c = 0; c < 6; c++) { AZ(b, 3, c) = (i8)(AZ(b, 3, c) + src[c]); src[c] = 0; }
    }
    int empty = 1;
    for (int f = 0; f < 5; f++) for (int c = 0; c < 6; c++) empty &= AZ(b, 4 + f, c) == 0;
    for (int c = 0; c < 5; c++) empty &= AZ(b, 3, c) == 0;
    if (!empty) return (player + 1) % 2;
    azul_score_round(b);
    const int next_player = azul_setup_new_round(b, seed, &rng);
    if (azul_game_over(b)) azul_score_bonuses(b);
    return next_player;
}
void azo_azul_check_end_game(const i8* b, float* out) {                                          /* check_end_game :283-302 */
    out[0] = out[1] = 0.f;
    if (!azul_game_over(b)) return;
    int rows[2] = {0, 0};
    for (int p = 0; p < 2; p++) for (int r = 0; r < 5; r++) { int all = 1; for (int c = 0; c < 5; c++) all &= AZ(b, 13 + 5 * p + r, c) == 1; rows[p] += all; }
    const int s0 = AZ(b, 0, 0), s1 = AZ(b, 0, 1);
    if (s0 > s1 || (s0 == s1 && rows[0] > rows[1])) { out[0] = 1.f; out[1] = -1.f; }
    else if (s1 > s0 || (s0 == s1 && rows[1] > rows[0])) { out[0] = -1.f; out[1] = 1.f; }
    else { out[0] = 0.01f; out[1] = 0.01f; }
}
void azo_azul_swap_players(i8* b) {                                                              /* swap_players :304-309 (unconditional) */
    i8 t = AZ(b, 0, 0); AZ(b, 0, 0) = AZ(b, 0, 1); AZ(b, 0, 1) = t;
    for (int c = 0; c < 6; c++) {
        t = AZ(b, 9, c); AZ(b, 9, c) = AZ(b, 10, c); AZ(b, 10, c) = t;
        t = AZ(b, 11, c); AZ(b, 11, c) = AZ(b, 12, c); AZ(b, 12, c) = t;
        for (int r = 0; r < 5; r++) { t = AZ(b, 13 + r, c); AZ(b, 13 + r, c) = AZ(b, 18 + r, c); AZ(b, 18 + r, c) = t; }
    }
}
void azo_azul_init_game(i8* b, uint64_t seed) {                                                  /* init_game :86-92 (draws from the oracle's RNG) */
    azo_rng rng; rng_seed(&rng, seed);
    memset(b, 0, AZU_S);
    for (int c = 0; c < 5; c++) { AZ(b, 1, c) = 20; AZ(b, 9, c) = -1; AZ(b, 10, c) = -1; }
    (void)azul_setup_new_round(b, 0, &rng);
}
/* get_symmetries :311-331: all 120 orders of the five factories (AzulLogic.py:4-126 lists them in lexicographic order). */
int azo_azul_symmetries(const i8* b, const float* pi, const u8* valids, i8* ob, float* opi, u8* ov) {
    int perm[5] = {0, 1, 2, 3, 4}, k = 0;
    for (;;) {
        i8* o = ob + (size_t)k * AZU_S; float* op = opi + (size_t)k * AZU_A; u8* om = ov + (size_t)k * AZU_A;
        memcpy(o, b, AZU_S); memcpy(op, pi, sizeof(float) * AZU_A); memcpy(om, valids, AZU_A);
        for (int i = 0; i < 5; i++) {
            memcpy(&AZ(o, 4 + i, 0), &AZ(b, 4 + perm[i], 0), 6);
            memcpy(op + 30 * (i + 1), pi + 30 * (perm[i] + 1), sizeof(float) * 30);
            memcpy(om + 30 * (i + 1), valids + 30 * (perm[i] + 1), 30);
        }
        k++;
        int i = 3; while (i >= 0 && perm[i] > perm[i + 1]) i--;                                   /* next lexicographic permutation */
        if (i < 0) break;
        int j = 4; while (perm[j] < perm[i]) j--;
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        for (int a = i + 1, z = 4; a < z; a++, z--) { t = perm[a]; perm[a] = perm[z]; perm[z] = t; }
    }
    return k;
}

/* ---------------------------------------------------------------- nets ------------------ */
/* (1) hash-net: see oracle/hashnet.py (test-only deterministic prior/value). */
static uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
void azo_hashnet_a(const i8* b, int S, const u8* valids, int A, int n, float* pi, float* v) {
    uint32_t h = 0x811C9DC5u;
    for (int i = 0; i < S; i++) h = (h ^ (u8)b[i]) * 16777619u;
    int64_t w[MAXA], W = 0, ksum = 0, k[MAXA]; int best = -1; int64_t bw = -1;
    for (int a = 0; a < A; a++) {
        w[a] = valids[a] ? 256 + (fmix32(h + (uint32_t)a * 0x9E3779B1u) & 1023) : 0;
        W += w[a]; if (w[a] > bw) { bw = w[a]; best = a; }
    }
    for (int a = 0; a < A; a++) { k[a] = (w[a] * 4096) / W; ksum += k[a]; }
    k[best] += 4096 - ksum;
    for (int a = 0; a < A; a++) pi[a] = (float)k[a] / 4096.0f;
    int j = (int)(fmix32(h ^ 0xABCDEF01u) % 129u) - 64;
    float v0 = (float)j / 64.0f;
    v[0] = v0; for (int p = 1; p < n; p++) v[p] = -v0 / (float)(n - 1);
}
void azo_hashnet(const i8* b, int S, const u8* valids, int n, float* pi, float* v) { azo_hashnet_a(b, S, valids, NA, n, pi, v); }

/* (2) SplendorNNet version 80, eval mode (splendor/SplendorNNet.py:149-204,259-280,397-404,440).
 * Weights arrive as one flat float32 blob in the tensor order listed in oracle/oracle.py:V80_ORDER. */
typedef struct { const float *w, *g, *b, *m, *v; } lin_bn;
typedef struct { lin_bn expand, dw, project; const float *fc1w, *fc1b, *fc2w, *fc2b; } irblock;
typedef struct {
    int nv;                 /* number of board rows (56 for 2 players) */
    lin_bn first; irblock blk[3];
    const float *pi2w, *pi2b, *pi4w, *pi4b, *v2w, *v2b, *v4w, *v4b;
    int np;
} v80_net;

static const float* take(const float** p, size_t n) { const float* r = *p; *p += n; return r; }
static lin_bn take_lin_bn(const float** p, int out, int in, int ch) {
    lin_bn l; l.w = take(p, (size_t)out * in); l.g = take(p, ch); l.b = take(p, ch); l.m = take(p, ch); l.v = take(p, ch); return l;
}
static int make_divisible8(int v) { int n = (v + 4) / 8 * 8; if (n < 8) n = 8; if (n < 0.9 * v) n += 8; return n; }   /* torchvision _make_divisible(v, 8) */
static void v80_bind(v80_net* N, const float* blob, int nv, int np) {
    const float* p = blob; int E = 3 * nv, Q = make_divisible8(E / 4); /* _make_divisible(E // 4, 8): 40 / 56 / 64 for nv = 56 / 71 / 88 */
    N->nv = nv; N->np = np;
    N->first = take_lin_bn(&p, nv, nv, nv);
    for (int k = 0; k < 3; k++) {
        irblock* B = &N->blk[k];
        B->expand = take_lin_bn(&p, E, nv, E);
        B->dw = take_lin_bn(&p, 7, 7, E);
        B->fc1w = take(&p, (size_t)Q * E); B->fc1b = take(&p, Q); B->fc2w = take(&p, (size_t)E * Q); B->fc2b = take(&p, E);
        B->project = take_lin_bn(&p, nv, E, nv);
    }
    N->pi2w = take(&p, (size_t)NA * nv * 7); N->pi2b = take(&p, NA); N->pi4w = take(&p, NA * NA); N->pi4b = take(&p, NA);
    N->v2w = take(&p, (size_t)np * nv * 7); N->v2b = take(&p, np); N->v4w = take(&p, (size_t)np * np); N->v4b = take(&p, np);
}
static float bn_apply(const lin_bn* l, int ch, float x) { return (x - l->m[ch]) / sqrtf(l->v[ch] + 1e-5f) * l->g[ch] + l->b[ch]; }
static float relu6f(float x) { return x < 0 ? 0 : x > 6 ? 6 : x; }
static float act(float x, int hs) { return hs ? x * relu6f(x + 3.f) / 6.f : (x > 0 ? x : 0); }

/* token-axis linear + BN (+act): y[o][f] = act(bn_o(sum_i W[o][i] x[i][f]))  (LinearNormActivation, non-depthwise) */
static void token_linear(const lin_bn* l, int out, int in, const float* x, float* y, int activation /*0 none,1 relu,2 hs*/) {
    for (int o = 0; o < out; o++)
        for (int f = 0; f < 7; f++) {
            float s = 0; for (int i = 0; i < in; i++) s += l->w[o * in + i] * x[i * 7 + f];
            s = bn_apply(l, o, s);
            y[o * 7 + f] = activation == 0 ? s : act(s, activation == 2);
        }
}
static void ir_block(const irblock* B, int nv, const float* x, float* y, int hs, int se_max) {
    int E = 3 * nv, Q = make_divisible8(E / 4);
    float e[3 * MAXROWS * 7], d[3 * MAXROWS * 7], sq[3 * MAXROWS], hid[64], sc[3 * MAXROWS], pr[MAXROWS * 7];
    token_linear(&B->expand, E, nv, x, e, hs ? 2 : 1);
    for (int c = 0; c < E; c++)                                  /* "depthwise": shared Linear(7->7) on the feature axis, BN per channel */
        for (int g = 0; g < 7; g++) {
            float s = 0; for (int f = 0; f < 7; f++) s += B->dw.w[g * 7 + f] * e[c * 7 + f];
            d[c * 7 + g] = act(bn_apply(&B->dw, c, s), hs);
        }
    for (int c = 0; c < E; c++) {                                /* squeeze: avg (trunk) or max (heads) over the 7 features */
        float s = se_max ? -1e30f : 0.f;
        for (int f = 0; f < 7; f++) s = se_max ? (d[c * 7 + f] > s ? d[c * 7 + f] : s) : s + d[c * 7 + f];
        sq[c] = se_max ? s : s / 7.f;
    }
    for (int q = 0; q < Q; q++) { float s = B->fc1b[q]; for (int c = 0; c < E; c++) s += B->fc1w[q * E + c] * sq[c]; hid[q] = s > 0 ? s : 0; }
    for (int c = 0; c < E; c++) { float s = B->fc2b[c]; for (int q = 0; q < Q; q++) s += B->fc2w[c * Q + q] * hid[q]; sc[c] = relu6f(s + 3.f) / 6.f; }
    for (int c = 0; c < E; c++) for (int f = 0; f < 7; f++) d[c * 7 + f] *= sc[c];
    token_linear(&B->project, nv, E, d, pr, 0);
    for (int i = 0; i < nv * 7; i++) y[i] = pr[i] + x[i];
}
static void v80_forward(const v80_net* N, const i8* board, const u8* valids, float* pi, float* v) {
    int nv = N->nv, F = nv * 7;
    float x[MAXROWS * 7], x0[MAXROWS * 7], t[MAXROWS * 7], hp[MAXROWS * 7], hv[MAXROWS * 7], h1[NA], logit[NA], hv1[MAXP];
    for (int i = 0; i < F; i++) x[i] = (float)board[i];
    token_linear(&N->first, nv, nv, x, x0, 0);
    ir_block(&N->blk[0], nv, x0, t, 0, 0);
    ir_block(&N->blk[1], nv, t, hp, 1, 1);
    ir_block(&N->blk[2], nv, t, hv, 1, 1);
    for (int o = 0; o < NA; o++) { float s = N->pi2b[o]; for (int i = 0; i < F; i++) s += N->pi2w[o * F + i] * hp[i]; h1[o] = s > 0 ? s : 0; }
    float mx = -INFINITY;
    for (int o = 0; o < NA; o++) {
        float s = N->pi4b[o]; for (int i = 0; i < NA; i++) s += N->pi4w[o * NA + i] * h1[i];
        logit[o] = valids[o] ? s : -1e8f; if (logit[o] > mx) mx = logit[o];
    }
    float se = 0; for (int o = 0; o < NA; o++) se += expf(logit[o] - mx);
    float lse = logf(se);
    for (int o = 0; o < NA; o++) pi[o] = expf(logit[o] - mx - lse);   /* exp(log_softmax), GenericNNetWrapper.py:119 */
    for (int p = 0; p < N->np; p++) { float s = N->v2b[p]; for (int i = 0; i < F; i++) s += N->v2w[p * F + i] * hv[i]; hv1[p] = s > 0 ? s : 0; }
    for (int p = 0; p < N->np; p++) { float s = N->v4b[p]; for (int q = 0; q < N->np; q++) s += N->v4w[p * N->np + q] * hv1[q]; v[p] = tanhf(s); }
}
void azo_v80_forward(const float* blob, int n_players, int batch, const i8* boards, const u8* valids, float* pi, float* v) {
    v80_net N; int nv = azo_state_rows(n_players); v80_bind(&N, blob, nv, n_players);
    for (int b = 0; b < batch; b++) v80_forward(&N, boards + (size_t)b * nv * 7, valids + (size_t)b * NA, pi + (size_t)b * NA, v + (size_t)b * n_players);
}

/* ---------------------------------------------------------------- AzulNNet V84 ------------ */
/* azul/AzulNNet.py:84-111 (layers), :127-137 (forward), eval mode; same building blocks as SplendorNNet V80 on 23 tokens x 6 features:
 * first_layer Linear(23->23)+BN, trunk InvertedResidual1d(23->115->23, ReLU, SE avg), policy head InvertedResidual1d(23->115->46,
 * Hardswish, SE avg, no residual) -> Linear(276->180)+ReLU -> Linear(180->180), value head InvertedResidual1d(23->46->23, Hardswish,
 * SE avg) -> Linear(138->2)+ReLU -> Linear(2->2). blob = state_dict tensors in the order of oracle.py:v80_order() (same module names).
 * ROUND-2 GROUNDWORK (no CUDA kernel yet); pinned by tests/golden/azul_v84_*.npz at 1e-5. */
#define V84_NV 23
#define V84_F 6
typedef struct { lin_bn expand, dw, project; const float *fc1w, *fc1b, *fc2w, *fc2b; int in, E, out, Q, hs; } ir84;
typedef struct { lin_bn first; ir84 blk[3]; const float *pi2w, *pi2b, *pi4w, *pi4b, *v2w, *v2b, *v4w, *v4b; } v84_net;
static void v84_bind(v84_net* N, const float* blob) {
    static const int E_[3] = {115, 115, 46}, OUT_[3] = {23, 46, 23}, Q_[3] = {32, 32, 16}, HS_[3] = {0, 1, 1};   /* Q = _make_divisible(E // 4, 8) */
    const float* p = blob;
    N->first = take_lin_bn(&p, V84_NV, V84_NV, V84_NV);
    for (int k = 0; k < 3; k++) {
        ir84* B = &N->blk[k]; B->in = V84_NV; B->E = E_[k]; B->out = OUT_[k]; B->Q = Q_[k]; B->hs = HS_[k];
        B->expand = take_lin_bn(&p, B->E, B->in, B->E);
        B->dw = take_lin_bn(&p, V84_F, V84_F, B->E);
        B->fc1w = take(&p, (size_t)B->Q * B->E); B->fc1b = take(&p, B->Q); B->fc2w = take(&p, (size_t)B->E * B->Q); B->fc2b = take(&p, B->E);
        B->project = take_lin_bn(&p, B->out, B->E, B->out);
    }
    N->pi2w = take(&p, (size_t)AZU_A * 46 * V84_F); N->pi2b = take(&p, AZU_A); N->pi4w = take(&p, (size_t)AZU_A * AZU_A); N->pi4b = take(&p, AZU_A);
    N->v2w = take(&p, 2 * V84_NV * V84_F); N->v2b = take(&p, 2); N->v4w = take(&p, 4); N->v4b = take(&p, 2);
}
static void token_linear6(const lin_bn* l, int out, int in, const float* x, float* y, int activation /*0 none,1 relu,2 hs*/) {
    for (int o = 0; o < out; o++)
        for (int f = 0; f < V84_F; f++) {
            float s = 0; for (int i = 0; i < in; i++) s += l->w[o * in + i] * x[i * V84_F + f];
            s = bn_apply(l, o, s);
            y[o * V84_F + f] = activation == 0 ? s : act(s, activation == 2);
        }
}
static void ir84_block(const ir84* B, const float* x, float* y) {
    const int E = B->E, Q = B->Q;
    float e[115 * V84_F], d[115 * V84_F], sq[115], hid[32], sc[115];
    token_linear6(&B->expand, E, B->in, x, e, B->hs ? 2 : 1);
    for (int c = 0; c < E; c++)                                  /* "depthwise": shared Linear(6->6) on the feature axis, BN per channel */
        for (int g = 0; g < V84_F; g++) {
            float s = 0; for (int f = 0; f < V84_F; f++) s += B->dw.w[g * V84_F + f] * e[c * V84_F + f];
            d[c * V84_F + g] = act(bn_apply(&B->dw, c, s), B->hs);
        }
    for (int c = 0; c < E; c++) { float s = 0.f; for (int f = 0; f < V84_F; f++) s += d[c * V84_F + f]; sq[c] = s / (float)V84_F; }   /* AdaptiveAvgPool1d(1) */
    for (int q = 0; q < Q; q++) { float s = B->fc1b[q]; for (int c = 0; c < E; c++) s += B->fc1w[q * E + c] * sq[c]; hid[q] = s > 0 ? s : 0; }
    for (int c = 0; c < E; c++) { float s = B->fc2b[c]; for (int q = 0; q < Q; q++) s += B->fc2w[c * Q + q] * hid[q]; sc[c] = relu6f(s + 3.f) / 6.f; }
    for (int c = 0; c < E; c++) for (int f = 0; f < V84_F; f++) d[c * V84_F + f] *= sc[c];
    token_linear6(&B->project, B->out, E, d, y, 0);
    if (B->in == B->out) for (int i = 0; i < B->out * V84_F; i++) y[i] += x[i];           /* use_res_connect */
}
static void v84_forward(const v84_net* N, const i8* board, const u8* valids, float* pi, float* v) {
    float x[V84_NV * V84_F], x0[V84_NV * V84_F], t[V84_NV * V84_F], hp[46 * V84_F], hv[V84_NV * V84_F], h1[AZU_A], logit[AZU_A], hv1[2];
    for (int i = 0; i < V84_NV * V84_F; i++) x[i] = (float)board[i];
    token_linear6(&N->first, V84_NV, V84_NV, x, x0, 0);
    ir84_block(&N->blk[0], x0, t);
    ir84_block(&N->blk[1], t, hp);
    ir84_block(&N->blk[2], t, hv);
    const int FP = 46 * V84_F, FV = V84_NV * V84_F;
    for (int o = 0; o < AZU_A; o++) { float s = N->pi2b[o]; for (int i = 0; i < FP; i++) s += N->pi2w[o * FP + i] * hp[i]; h1[o] = s > 0 ? s : 0; }
    float mx = -INFINITY;
    for (int o = 0; o < AZU_A; o++) {
        float s = N->pi4b[o]; for (int i = 0; i < AZU_A; i++) s += N->pi4w[o * AZU_A + i] * h1[i];
        logit[o] = valids[o] ? s : -1e8f; if (logit[o] > mx) mx = logit[o];
    }
    float se = 0; for (int o = 0; o < AZU_A; o++) se += expf(logit[o] - mx);
    const float lse = logf(se);
    for (int o = 0; o < AZU_A; o++) pi[o] = expf(logit[o] - mx - lse);   /* exp(log_softmax), GenericNNetWrapper.py:119 */
    for (int p = 0; p < 2; p++) { float s = N->v2b[p]; for (int i = 0; i < FV; i++) s += N->v2w[p * FV + i] * hv[i]; hv1[p] = s > 0 ? s : 0; }
    for (int p = 0; p < 2; p++) { float s = N->v4b[p]; for (int q = 0; q < 2; q++) s += N->v4w[p * 2 + q] * hv1[q]; v[p] = tanhf(s); }
}
void azo_v84_forward(const float* blob, int batch, const i8* boards, const u8* valids, float* pi, float* v) {
    v84_net N; v84_bind(&N, blob);
    for (int b = 0; b < batch; b++) v84_forward(&N, boards + (size_t)b * AZU_S, valids + (size_t)b * AZU_A, pi + (size_t)b * AZU_A, v + (size_t)b * 2);
}

/* ---------------------------------------------------------------- SantoriniNNet V89 ------ */
/* santorini/SantoriniNNet.py:70-84 (SimpleResBlock), :16-40 (SimpleHead), :194-217 (layers), :273-279 (forward), eval mode.
 * blob = state_dict tensors in the order of oracle.py:v89_order() (conv weight, then BN weight/bias/mean/var; heads). */
typedef struct { const float *w, *g, *b, *m, *v; } conv_bn;
typedef struct { conv_bn first, c1[5], c2[5], pi, vv; const float *pifc, *pifcb, *vfc1, *vfc1b, *vfc2, *vfc2b; } v89_net;
static conv_bn take_conv_bn(const float** p, int out, int cin, int k) {
    conv_bn c; c.w = take(p, (size_t)out * cin * k * k); c.g = take(p, out); c.b = take(p, out); c.m = take(p, out); c.v = take(p, out); return c;
}
static void v89_bind(v89_net* N, const float* blob) {
    const float* p = blob;
    N->first = take_conv_bn(&p, 64, 2, 3);
    for (int i = 0; i < 5; i++) { N->c1[i] = take_conv_bn(&p, 64, 64, 3); N->c2[i] = take_conv_bn(&p, 64, 64, 3); }
    N->pi = take_conv_bn(&p, 2, 64, 1); N->pifc = take(&p, 162 * 50); N->pifcb = take(&p, 162);
    N->vv = take_conv_bn(&p, 1, 64, 1); N->vfc1 = take(&p, 64 * 25); N->vfc1b = take(&p, 64); N->vfc2 = take(&p, 2 * 64); N->vfc2b = take(&p, 2);
}
/* Conv2d(cin->cout, k x k, padding k/2, no bias) + BatchNorm2d (eval) on a 5x5 plane stack x[cin][25] -> y[cout][25] */
static void conv_bn_apply(const conv_bn* c, int cout, int cin, int k, const float* x, float* y) {
    int r = k / 2;
    for (int o = 0; o < cout; o++)
        for (int py = 0; py < 5; py++)
            for (int px = 0; px < 5; px++) {
                float s = 0;
                for (int ci = 0; ci < cin; ci++)
                    for (int ky = 0; ky < k; ky++)
                        for (int kx = 0; kx < k; kx++) {
                            int yy = py + ky - r, xx = px + kx - r;
                            if (yy < 0 || yy >= 5 || xx < 0 || xx >= 5) continue;
                            s += c->w[((o * cin + ci) * k + ky) * k + kx] * x[ci * 25 + yy * 5 + xx];
                        }
                y[o * 25 + py * 5 + px] = (s - c->m[o]) / sqrtf(c->v[o] + 1e-5f) * c->g[o] + c->b[o];
            }
}
static void v89_forward(const v89_net* N, const i8* board, const u8* valids, float* pi, float* v) {
    float x[2 * 25], a[64 * 25], h[64 * 25], t[64 * 25];
    for (int c = 0; c < 2; c++) for (int pos = 0; pos < 25; pos++) x[c * 25 + pos] = (float)board[pos * 3 + c];   /* permute(0,3,1,2), channels 0-1 */
    conv_bn_apply(&N->first, 64, 2, 3, x, a);
    for (int i = 0; i < 64 * 25; i++) a[i] = a[i] > 0 ? a[i] : 0;
    for (int blk = 0; blk < 5; blk++) {
        conv_bn_apply(&N->c1[blk], 64, 64, 3, a, h);
        for (int i = 0; i < 64 * 25; i++) h[i] = h[i] > 0 ? h[i] : 0;
        conv_bn_apply(&N->c2[blk], 64, 64, 3, h, t);
        for (int i = 0; i < 64 * 25; i++) { float s = t[i] + a[i]; a[i] = s > 0 ? s : 0; }
    }
    float pf[50], vf[25], logit[SAN_A], hv[64];
    conv_bn_apply(&N->pi, 2, 64, 1, a, pf);
    for (int i = 0; i < 50; i++) pf[i] = pf[i] > 0 ? pf[i] : 0;
    float mx = -INFINITY;
    for (int o = 0; o < SAN_A; o++) {
        float s = N->pifcb[o]; for (int k = 0; k < 50; k++) s += N->pifc[o * 50 + k] * pf[k];
        logit[o] = valids[o] ? s : -1e8f; if (logit[o] > mx) mx = logit[o];
    }
    float se = 0; for (int o = 0; o < SAN_A; o++) se += expf(logit[o] - mx);
    float lse = logf(se);
    for (int o = 0; o < SAN_A; o++) pi[o] = expf(logit[o] - mx - lse);
    conv_bn_apply(&N->vv, 1, 64, 1, a, vf);
    for (int i = 0; i < 25; i++) vf[i] = vf[i] > 0 ? vf[i] : 0;
    for (int j = 0; j < 64; j++) { float s = N->vfc1b[j]; for (int k = 0; k < 25; k++) s += N->vfc1[j * 25 + k] * vf[k]; hv[j] = s > 0 ? s : 0; }
    for (int o = 0; o < 2; o++) { float s = N->vfc2b[o]; for (int j = 0; j < 64; j++) s += N->vfc2[o * 64 + j] * hv[j]; v[o] = tanhf(s); }
}
void azo_v89_forward(const float* blob, int batch, const i8* boards, const u8* valids, float* pi, float* v) {
    v89_net N; v89_bind(&N, blob);
    for (int b = 0; b < batch; b++) v89_forward(&N, boards + (size_t)b * SAN_S, valids + (size_t)b * SAN_A, pi + (size_t)b * SAN_A, v + (size_t)b * 2);
}

/* ---------------------------------------------------------------- AbaloneNNet V21 -------- */
/* abalone/AbaloneNNet.py:117-156 (layers; torchvision InvertedResidual 24->48->24, kernel 3, no SE, ReLU, BatchNorm2d eps 1e-5),
 * :173-202 (forward), eval mode. blob = state_dict tensors in the order of oracle.py:v21_order(). Planes are [c][9*r+q]. */
typedef struct { conv_bn first, ex[4], dw[4], pr[4], pi, vc; const float *mw, *mb, *f1, *f1b, *f2, *f2b; } v21_net;
static void v21_bind(v21_net* N, const float* blob) {
    const float* p = blob;
    N->first = take_conv_bn(&p, 24, 3, 3);
    for (int i = 0; i < 4; i++) { N->ex[i] = take_conv_bn(&p, 48, 24, 1); N->dw[i] = take_conv_bn(&p, 48, 1, 3); N->pr[i] = take_conv_bn(&p, 24, 48, 1); }
    N->mw = take(&p, 16 * 6); N->mb = take(&p, 16);
    N->pi = take_conv_bn(&p, 42, 24, 1); N->vc = take_conv_bn(&p, 4, 24, 1);
    N->f1 = take(&p, 64 * 340); N->f1b = take(&p, 64); N->f2 = take(&p, 2 * 64); N->f2b = take(&p, 2);
}
/* Conv2d(k x k, padding k/2, no bias; depthwise when dwise) + BatchNorm2d (eval) on 9x9 planes */
static void conv9_bn(const conv_bn* c, int cout, int cin, int k, int dwise, const float* x, float* y) {
    int r = k / 2;
    for (int o = 0; o < cout; o++)
        for (int py = 0; py < 9; py++)
            for (int px = 0; px < 9; px++) {
                float s = 0;
                for (int ci = 0; ci < (dwise ? 1 : cin); ci++)
                    for (int ky = 0; ky < k; ky++)
                        for (int kx = 0; kx < k; kx++) {
                            int yy = py + ky - r, xx = px + kx - r;
                            if (yy < 0 || yy >= 9 || xx < 0 || xx >= 9) continue;
                            s += c->w[((o * (dwise ? 1 : cin) + ci) * k + ky) * k + kx] * x[(dwise ? o : ci) * 81 + yy * 9 + xx];
                        }
                y[o * 81 + py * 9 + px] = (s - c->m[o]) / sqrtf(c->v[o] + 1e-5f) * c->g[o] + c->b[o];
            }
}
static void v21_forward(const v21_net* N, const i8* board, const u8* valids, float* pi, float* v) {
    static __thread float x[3 * 81], a[24 * 81], e[48 * 81], d[48 * 81], t[24 * 81], lg[42 * 81], vf[4 * 81], logit[ABA_A];
    for (int c = 0; c < 3; c++) for (int pos = 0; pos < 81; pos++) x[c * 81 + pos] = (float)board[pos * 4 + c];
    conv9_bn(&N->first, 24, 3, 3, 0, x, a);
    for (int i = 0; i < 24 * 81; i++) a[i] = a[i] > 0 ? a[i] : 0;
    for (int blk = 0; blk < 4; blk++) {
        conv9_bn(&N->ex[blk], 48, 24, 1, 0, a, e); for (int i = 0; i < 48 * 81; i++) e[i] = e[i] > 0 ? e[i] : 0;
        conv9_bn(&N->dw[blk], 48, 48, 3, 1, e, d); for (int i = 0; i < 48 * 81; i++) d[i] = d[i] > 0 ? d[i] : 0;
        conv9_bn(&N->pr[blk], 24, 48, 1, 0, d, t); for (int i = 0; i < 24 * 81; i++) a[i] = a[i] + t[i];
    }
    conv9_bn(&N->pi, 42, 24, 1, 0, a, lg);
    float mx = -INFINITY;
    for (int pos = 0; pos < 81; pos++) for (int pl = 0; pl < 42; pl++) {         /* permute(0,2,3,1): action = 42*pos + plane */
        int o = pos * 42 + pl; logit[o] = valids[o] ? lg[pl * 81 + pos] : -1e8f; if (logit[o] > mx) mx = logit[o];
    }
    float se = 0; for (int o = 0; o < ABA_A; o++) se += expf(logit[o] - mx);
    float lse = logf(se);
    for (int o = 0; o < ABA_A; o++) pi[o] = expf(logit[o] - mx - lse);
    float meta[6], me[16], vc[340], h[64];
    for (int j = 0; j < 6; j++) meta[j] = (float)board[j * 4 + 3];                /* input[:, 0, 0:6, 3] */
    for (int o = 0; o < 16; o++) { float sacc = N->mb[o]; for (int j = 0; j < 6; j++) sacc += N->mw[o * 6 + j] * meta[j]; me[o] = sacc > 0 ? sacc : 0; }
    conv9_bn(&N->vc, 4, 24, 1, 0, a, vf);
    for (int i = 0; i < 324; i++) vc[i] = vf[i] > 0 ? vf[i] : 0;
    for (int i = 0; i < 16; i++) vc[324 + i] = me[i];
    for (int o = 0; o < 64; o++) { float sacc = N->f1b[o]; for (int k = 0; k < 340; k++) sacc += N->f1[o * 340 + k] * vc[k]; h[o] = sacc > 0 ? sacc : 0; }
    for (int o = 0; o < 2; o++) { float sacc = N->f2b[o]; for (int k = 0; k < 64; k++) sacc += N->f2[o * 64 + k] * h[k]; v[o] = tanhf(sacc); }
}
void azo_v21_forward(const float* blob, int batch, const i8* boards, const u8* valids, float* pi, float* v) {
    v21_net N; v21_bind(&N, blob);
    for (int b = 0; b < batch; b++) v21_forward(&N, boards + (size_t)b * ABA_S, valids + (size_t)b * ABA_A, pi + (size_t)b * ABA_A, v + (size_t)b * 2);
}

/* ---------------------------------------------------------------- MCTS ------------------ */
#define NAN_Q (-42.0)
static const int64_t MAGIC_SEEDS[8] = {31416, 1, 14142, 42, 27183, 2, 16180, 7};   /* MCTS.py:14 */

typedef struct {
    int num_players, numMCTSSims, ratio_fullMCTS, universes, forced_playouts, no_mem_optim, net_kind /*0 hash,1 v80,2 v89,3 v21,4 v84*/;
    double cpuct, fpu, dirichletAlpha, prob_fullMCTS, temperature2;
    int game /*0 splendor, 1 santorini without gods, 2 abalone*/;
} azo_cfg;

typedef struct node {
    i8 key[MAXS];
    int has_es, expanded, r; float Es[MAXP];
    int64_t Ns; float Qs;            /* per-action arrays Vs/Ps/Qsa/Nsa live in the tree's slabs (row = node index, A entries) */
    int used;
} node_t;

typedef struct {
    azo_cfg cfg; int S, A; v80_net net; v89_net net89; v21_net net21; v84_net net84; const float* blob;
    node_t* nodes; int* table; int cap, tcap, count;
    u8* sVs; float* sPs; double* sQsa; int64_t* sNsa;        /* [cap][A] slabs */
    int dirichlet_noise, step, last_cleaning; int64_t random_seed;
    azo_rng rng;
    double inj_u_full;               /* >= 0: injected playout-cap coin for the next getActionProb (replay of a recorded episode) */
    /* counters */
    int64_t n_sims, n_expansions, n_node_visits, n_nn_evals;
} azo_mcts;

static uint64_t key_hash(const i8* k, int S) { uint64_t h = 1469598103934665603ULL; for (int i = 0; i < S; i++) h = (h ^ (u8)k[i]) * 1099511628211ULL; return h; }
static void table_rebuild(azo_mcts* m) {
    for (int i = 0; i < m->tcap; i++) m->table[i] = -1;
    for (int i = 0; i < m->count; i++) { uint64_t h = key_hash(m->nodes[i].key, m->S) & (uint64_t)(m->tcap - 1); while (m->table[h] >= 0) h = (h + 1) & (uint64_t)(m->tcap - 1); m->table[h] = i; }
}
#define N_VS(m, nd) ((m)->sVs + (size_t)((nd) - (m)->nodes) * (size_t)(m)->A)
#define N_PS(m, nd) ((m)->sPs + (size_t)((nd) - (m)->nodes) * (size_t)(m)->A)
#define N_QSA(m, nd) ((m)->sQsa + (size_t)((nd) - (m)->nodes) * (size_t)(m)->A)
#define N_NSA(m, nd) ((m)->sNsa + (size_t)((nd) - (m)->nodes) * (size_t)(m)->A)
static void slabs_alloc(azo_mcts* m) {
    size_t n = (size_t)m->cap * (size_t)m->A;
    m->sVs = (u8*)realloc(m->sVs, n); m->sPs = (float*)realloc(m->sPs, n * sizeof(float));
    m->sQsa = (double*)realloc(m->sQsa, n * sizeof(double)); m->sNsa = (int64_t*)realloc(m->sNsa, n * sizeof(int64_t));
}
static void grow(azo_mcts* m) {
    m->cap *= 2; m->tcap *= 2;
    m->nodes = (node_t*)realloc(m->nodes, sizeof(node_t) * (size_t)m->cap);
    m->table = (int*)realloc(m->table, sizeof(int) * (size_t)m->tcap);
    slabs_alloc(m);
    table_rebuild(m);
}
static node_t* lookup(azo_mcts* m, const i8* key) {
    uint64_t h = key_hash(key, m->S) & (uint64_t)(m->tcap - 1);
    while (m->table[h] >= 0) { node_t* nd = &m->nodes[m->table[h]]; if (memcmp(nd->key, key, (size_t)m->S) == 0) return nd; h = (h + 1) & (uint64_t)(m->tcap - 1); }
    return NULL;
}
static node_t* insert(azo_mcts* m, const i8* key) {
    if (m->count + 1 > m->cap) grow(m);
    node_t* nd = &m->nodes[m->count]; memset(nd, 0, sizeof(*nd)); memcpy(nd->key, key, (size_t)m->S);
    uint64_t h = key_hash(key, m->S) & (uint64_t)(m->tcap - 1); while (m->table[h] >= 0) h = (h + 1) & (uint64_t)(m->tcap - 1);
    m->table[h] = m->count++; return nd;
}

azo_mcts* azo_mcts_new(const azo_cfg* cfg, const float* blob, int dirichlet_noise, uint64_t seed) {
    azo_mcts* m = (azo_mcts*)calloc(1, sizeof(azo_mcts));
    m->cfg = *cfg; m->blob = blob;
    if (cfg->game == 1) { m->S = SAN_S; m->A = SAN_A; } else if (cfg->game == 2) { m->S = ABA_S; m->A = ABA_A; } else if (cfg->game == 3) { m->S = AZU_S; m->A = AZU_A; } else { m->S = azo_state_rows(cfg->num_players) * COLS; m->A = NA; }
    if (cfg->net_kind == 1) v80_bind(&m->net, blob, azo_state_rows(cfg->num_players), cfg->num_players);
    if (cfg->net_kind == 2) v89_bind(&m->net89, blob);
    if (cfg->net_kind == 3) v21_bind(&m->net21, blob);
    if (cfg->net_kind == 4) v84_bind(&m->net84, blob);
    m->cap = m->A > 1000 ? 512 : 4096; m->tcap = 4 * m->cap; m->nodes = (node_t*)malloc(sizeof(node_t) * (size_t)m->cap); m->table = (int*)malloc(sizeof(int) * (size_t)m->tcap);
    slabs_alloc(m);
    m->count = 0; table_rebuild(m); m->dirichlet_noise = dirichlet_noise; m->random_seed = -1; rng_seed(&m->rng, seed); m->inj_u_full = -1.0;
    return m;
}
void azo_mcts_free(azo_mcts* m) { if (m) { free(m->nodes); free(m->table); free(m->sVs); free(m->sPs); free(m->sQsa); free(m->sNsa); free(m); } }
void azo_mcts_reset(azo_mcts* m) { m->count = 0; m->last_cleaning = 0; table_rebuild(m); }
void azo_mcts_stats(const azo_mcts* m, int64_t* out) {
    int64_t nt = 0, sns = 0;
    for (int i = 0; i < m->count; i++) { if (!m->nodes[i].expanded) nt++; else sns += m->nodes[i].Ns; }
    out[0] = m->count; out[1] = nt; out[2] = sns; out[3] = m->n_sims; out[4] = m->n_expansions; out[5] = m->n_node_visits; out[6] = m->n_nn_evals;
}

/* float32 sum in the order numba's vectorised np.sum uses on the build host (see file header) */
static float sum_f32_avx2(const float* x, int n) {
    float s = 0.f; int i = 0;
    if (n >= 32) {
        float acc[32]; for (int j = 0; j < 32; j++) acc[j] = 0.f;
        int nb = n / 32;
        for (int b = 0; b < nb; b++) for (int j = 0; j < 32; j++) acc[j] = acc[j] + x[32 * b + j];
        float t[8], u[4];
        for (int j = 0; j < 8; j++) t[j] = (acc[8 + j] + acc[j]) + (acc[16 + j] + acc[24 + j]);
        for (int j = 0; j < 4; j++) u[j] = t[j + 4] + t[j];
        s = (u[0] + u[2]) + (u[1] + u[3]);
        i = 32 * nb;
    }
    if ((n & 28) != 0 && (n & ~3) > i) {
        float q0 = s, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        for (; i < (n & ~3); i += 4) { q0 += x[i]; q1 += x[i + 1]; q2 += x[i + 2]; q3 += x[i + 3]; }
        s = (q0 + q2) + (q1 + q3);
    }
    for (; i < n; i++) s += x[i];
    return s;
}
/* MCTS.py:250-253 */
static void normalise_f32(float* x, int n) { float inv = 1.0f / sum_f32_avx2(x, n); for (int i = 0; i < n; i++) x[i] = x[i] * inv; }
/* MCTS.py:255-261 */
static void softmax_temp(float* P, int n, double T) {
    if (T == 1.0) return;
    double r[MAXA], s = 0; for (int i = 0; i < n; i++) { r[i] = pow((double)P[i], 1.0 / T); s += r[i]; }
    double inv = 1.0 / s; for (int i = 0; i < n; i++) P[i] = (float)(r[i] * inv);
}
/* ---- game dispatch of the Game.py methods the search calls (MCTS.py:125-173) ---- */
static void g_ended(const azo_mcts* m, const i8* b, int next_player, float* out) {
    if (m->cfg.game == 1) azo_sant_check_end_game(b, next_player, out); else if (m->cfg.game == 2) azo_aba_check_end_game(b, out); else if (m->cfg.game == 3) azo_azul_check_end_game(b, out); else azo_check_end_game(b, m->cfg.num_players, out);
}
static void g_valid(const azo_mcts* m, const i8* b, u8* out) {
    if (m->cfg.game == 1) azo_sant_valid_moves(b, 0, out); else if (m->cfg.game == 2) azo_aba_valid_moves(b, 0, out); else if (m->cfg.game == 3) azo_azul_valid_moves(b, 0, out); else azo_valid_moves(b, m->cfg.num_players, 0, out);
}
static int g_move(const azo_mcts* m, i8* b, int a, int player, int64_t seed, azo_rng* rng) {
    return m->cfg.game == 1 ? azo_sant_make_move(b, a, player) : m->cfg.game == 2 ? azo_aba_make_move(b, a, player) : m->cfg.game == 3 ? azo_azul_make_move(b, a, player, seed, 1) : azo_make_move(b, m->cfg.num_players, a, player, seed, rng);
}
static void g_swap(const azo_mcts* m, i8* b, int nb) {
    if (m->cfg.game == 1) azo_sant_swap_players(b, nb); else if (m->cfg.game == 2) azo_aba_swap_players(b, nb); else if (m->cfg.game == 3) { if (nb & 1) azo_azul_swap_players(b); } else azo_swap_players(b, m->cfg.num_players, nb);
}
static int g_round(const azo_mcts* m, const i8* b) { return m->cfg.game == 1 ? azo_sant_get_round(b) : m->cfg.game == 2 ? azo_aba_get_round(b) : m->cfg.game == 3 ? azo_azul_get_round(b) : azo_get_round(b); }

/* MCTS.py:187-197; `noise` (length = number of legal actions) is either injected by the caller
 * (parity tests replay the reference's draws) or sampled here. */
static void apply_dir_noise(azo_mcts* m, float* P, const u8* Vs, const double* noise) {
    const int A = m->A;
    int L = 0; for (int a = 0; a < A; a++) L += Vs[a] != 0;
    double tmp[MAXA];
    if (!noise) {
        double alpha = m->cfg.dirichletAlpha > 0 ? m->cfg.dirichletAlpha : 10.0 / L, s = 0;
        for (int i = 0; i < L; i++) { tmp[i] = rng_gamma(&m->rng, alpha); s += tmp[i]; }
        for (int i = 0; i < L; i++) tmp[i] /= s;
        noise = tmp;
    }
    int k = 0;
    for (int a = 0; a < A; a++) if (Vs[a]) { float t1 = 0.75f * P[a]; P[a] = (float)((double)t1 + 0.25 * noise[k]); k++; }
}

/* MCTS.py:210-230 */
static int pick_highest_ucb(const node_t* nd, const u8* Vs, const float* Ps, const double* Qsa, const int64_t* Nsa, int A, double cpuct, int forced, int64_t n_iter, double fpu) {
    double best = -INFINITY; int best_a = -1;
    double fpu_init = fpu > 0 ? (double)nd->Qs - fpu : fpu;
    double c0 = cpuct * sqrt((double)nd->Ns + 1e-8), c1 = cpuct * sqrt((double)nd->Ns), kn = (double)n_iter * 0.5;
    for (int a = 0; a < A; a++) {
        if (!Vs[a]) continue;
        if (forced && Nsa[a] < (int64_t)sqrt(kn * (double)Ps[a])) return a;
        double u;
        if (Qsa[a] != NAN_Q) u = Qsa[a] + (c1 * (double)Ps[a]) / (double)(Nsa[a] + 1);
        else u = fma(c0, (double)Ps[a], fpu_init);
        if (u > best) { best = u; best_a = a; }
    }
    return best_a;
}

/* MCTS.py:105-184 (search), unrolled from recursion into select / leaf / backup. */
static void search(azo_mcts* m, const i8* root, int dir_noise, int forced, const double* noise) {
    int n = m->cfg.num_players, S = m->S, depth = 0; const int A = m->A;
    static __thread int path_node[256], path_a[256], path_np[256];   /* node INDICES: insert() may realloc m->nodes */
    i8 cur[MAXS]; memcpy(cur, root, (size_t)S);
    float v[MAXP];
    azo_rng dummy; rng_seed(&dummy, 1);
    m->n_sims++;
    for (;;) {
        node_t* nd = lookup(m, cur);
        if (!nd || !nd->has_es) {
            float Es[MAXP]; g_ended(m, cur, 0, Es);                       /* MCTS.py:131 getGameEnded(canonicalBoard, 0) */
            int any = 0; for (int p = 0; p < n; p++) any |= Es[p] != 0.f;
            if (!nd) { nd = insert(m, cur); nd->r = g_round(m, cur); }
            nd->has_es = 1; memcpy(nd->Es, Es, sizeof(float) * (size_t)n);
            if (any) { memcpy(v, Es, sizeof(float) * (size_t)n); break; }
        } else {
            int any = 0; for (int p = 0; p < n; p++) any |= nd->Es[p] != 0.f;
            if (any) { memcpy(v, nd->Es, sizeof(float) * (size_t)n); break; }
        }
        if (!nd->expanded) {
            u8* Vs = N_VS(m, nd); float* Ps = N_PS(m, nd); double* Qsa = N_QSA(m, nd); int64_t* Nsa = N_NSA(m, nd);
            g_valid(m, cur, Vs);
            if (m->cfg.net_kind == 0) azo_hashnet_a(cur, S, Vs, A, n, Ps, v);
            else if (m->cfg.net_kind == 2) v89_forward(&m->net89, cur, Vs, Ps, v);
            else if (m->cfg.net_kind == 3) v21_forward(&m->net21, cur, Vs, Ps, v);
            else if (m->cfg.net_kind == 4) v84_forward(&m->net84, cur, Vs, Ps, v);
            else v80_forward(&m->net, cur, Vs, Ps, v);
            m->n_nn_evals++; m->n_expansions++;
            if (depth == 0 && dir_noise) { softmax_temp(Ps, A, m->cfg.temperature2); apply_dir_noise(m, Ps, Vs, noise); }
            normalise_f32(Ps, A);
            nd->Ns = 0; for (int a = 0; a < A; a++) { Qsa[a] = NAN_Q; Nsa[a] = 0; }
            nd->Qs = v[0]; nd->expanded = 1;
            break;
        }
        if (depth == 0 && dir_noise) { softmax_temp(N_PS(m, nd), A, m->cfg.temperature2); apply_dir_noise(m, N_PS(m, nd), N_VS(m, nd), noise); normalise_f32(N_PS(m, nd), A); }
        int a = pick_highest_ucb(nd, N_VS(m, nd), N_PS(m, nd), N_QSA(m, nd), N_NSA(m, nd), m->A, m->cfg.cpuct, depth == 0 && forced, m->step, m->cfg.fpu);
        m->n_node_visits++;
        int np_ = g_move(m, cur, a, 0, m->random_seed, &dummy);           /* MCTS.py:233-248 */
        if (np_ != 0) g_swap(m, cur, np_);
        path_node[depth] = (int)(nd - m->nodes); path_a[depth] = a; path_np[depth] = np_; depth++;
    }
    for (int d = depth - 1; d >= 0; d--) {
        float t[MAXP]; for (int p = 0; p < n; p++) t[(p + path_np[d]) % n] = v[p];       /* np.roll(v, next_player) */
        memcpy(v, t, sizeof(float) * (size_t)n);
        node_t* nd = &m->nodes[path_node[d]]; int a = path_a[d];
        double* Qsa = N_QSA(m, nd); int64_t* Nsa = N_NSA(m, nd);
        Qsa[a] = ((double)Nsa[a] * Qsa[a] + (double)v[0]) / (double)(Nsa[a] + 1);
        nd->Qs = ((float)(nd->Ns + 1) * nd->Qs + v[0]) / (float)(nd->Ns + 2);
        Nsa[a] += 1; nd->Ns += 1;
    }
}

/* MCTS.py:49-103 (getActionProb). noise: injected Dirichlet vector or NULL. Outputs: probs f64[A], q f32[np],
 * raw_counts i64[A] (root Nsa before policy-target pruning). Returns is_full_search. */
int azo_mcts_get_action_prob(azo_mcts* m, const i8* cb, double temp, int force_full, const double* noise,
                             double* probs, float* q, int64_t* raw_counts) {
    const azo_cfg* c = &m->cfg; int n = c->num_players; const int A = m->A;
    int full = force_full || ((m->inj_u_full >= 0 ? m->inj_u_full : rng_uniform(&m->rng)) < c->prob_fullMCTS);
    int nsims = full ? c->numMCTSSims : c->numMCTSSims / c->ratio_fullMCTS;
    int forced = full && c->forced_playouts;
    for (m->step = 0; m->step < nsims; m->step++) {
        m->random_seed = c->universes > 0 ? MAGIC_SEEDS[m->step % c->universes] : -1;
        search(m, cb, m->step == 0 && full && m->dirichlet_noise, forced, noise);
    }
    if (nsims > 0) m->step = nsims - 1;
    node_t* root = lookup(m, cb);
    double counts[MAXA];
    const int64_t* rNsa = N_NSA(m, root); const float* rPs = N_PS(m, root);
    for (int a = 0; a < A; a++) { counts[a] = (double)rNsa[a]; if (raw_counts) raw_counts[a] = rNsa[a]; }
    q[0] = root->Qs; for (int p = 1; p < n; p++) q[p] = -root->Qs / (float)(n - 1);
    if (forced) {
        double best = 0; for (int a = 0; a < A; a++) if (counts[a] > best) best = counts[a];
        for (int a = 0; a < A; a++) {
            double cnt = counts[a];
            if (cnt != best) { float t = 0.5f * rPs[a]; t = t * (float)nsims; cnt = cnt - (double)(int64_t)sqrt((double)t); }
            counts[a] = cnt > 1 ? cnt : 0;
        }
    }
    if (!c->no_mem_optim) {                                                /* MCTS.py:86-91 */
        int r = g_round(m, cb);
        if (r > m->last_cleaning + 20) {
            int w = 0;
            for (int i = 0; i < m->count; i++) if (!(m->nodes[i].r < r - 5)) {
                if (w != i) {
                    size_t Az = (size_t)m->A;
                    m->nodes[w] = m->nodes[i];
                    memcpy(m->sVs + w * Az, m->sVs + i * Az, Az); memcpy(m->sPs + w * Az, m->sPs + i * Az, Az * sizeof(float));
                    memcpy(m->sQsa + w * Az, m->sQsa + i * Az, Az * sizeof(double)); memcpy(m->sNsa + w * Az, m->sNsa + i * Az, Az * sizeof(int64_t));
                }
                w++;
            }
            m->count = w; table_rebuild(m); m->last_cleaning = r;
        }
    }
    if (temp <= 0.02) {
        double best = -1; int nb = 0; for (int a = 0; a < A; a++) { if (counts[a] > best) { best = counts[a]; nb = 1; } else if (counts[a] == best) nb++; }
        int pick = (int)(rng_uniform(&m->rng) * nb), k = 0; if (pick >= nb) pick = nb - 1;
        for (int a = 0; a < A; a++) { probs[a] = 0; if (counts[a] == best) { if (k == pick) probs[a] = 1; k++; } }
        return full;
    }
    double s = 0; for (int a = 0; a < A; a++) { counts[a] = pow(counts[a], 1.0 / temp); s += counts[a]; }
    for (int a = 0; a < A; a++) probs[a] = counts[a] / s;
    return full;
}

/* ---------------------------------------------------------------- episode (Coach.py:37-84) */
typedef struct { int64_t sims, expansions, node_visits, nn_evals, plies, examples, games; double seconds; } azo_run_stats;

static double temp_for_selfplay(double t_begin, double t_end, double half_life, int n) {   /* Coach.py:266-271 */
    if (half_life < 0) return n > -half_life ? t_end : t_begin;
    return t_end + (t_begin - t_end) * pow(0.5, n / half_life);
}

/* One self-play game. Example boards/policies are only counted (the CPU baseline measures search
 * throughput); `max_plies` > 0 truncates the game (bounded sample for bench.py). */
static void execute_episode(azo_mcts* m, uint64_t seed, double t0, double t1, double half, int max_plies, azo_run_stats* st) {
    int n = m->cfg.num_players, S = m->S; const int A = m->A; azo_rng rng; rng_seed(&rng, seed ^ 0xA5A5A5A5ULL);
    i8 board[MAXS], cb[MAXS];
    if (m->cfg.game == 1) azo_sant_init_game(board, seed); else if (m->cfg.game == 2) azo_aba_init_game(board); else if (m->cfg.game == 3) azo_azul_init_game(board, seed); else azo_init_game(board, n, seed);
    int player = 0, step = 0; azo_mcts_reset(m);
    double probs[MAXA]; float q[MAXP], r[MAXP];
    for (;;) {
        step++;
        memcpy(cb, board, (size_t)S); if (player) g_swap(m, cb, player);
        int64_t before = m->n_sims;
        int full = azo_mcts_get_action_prob(m, cb, 1.0, 0, NULL, probs, q, NULL);
        (void)before;
        double T = temp_for_selfplay(t0, t1, half, step), w[MAXA], s = 0;
        int action = -1;
        if (T == 0) { double b = -1; for (int a = 0; a < A; a++) if (probs[a] > b) { b = probs[a]; action = a; } }
        else {
            for (int a = 0; a < A; a++) { w[a] = pow(probs[a], 1.0 / T); s += w[a]; }
            double u = rng_uniform(&rng) * s, acc = 0;
            for (int a = 0; a < A; a++) { if (w[a] > 0) { action = a; acc += w[a]; if (acc > u) break; } }
        }
        if (full) { u8 V[MAXA]; g_valid(m, cb, V); st->examples += 1; }
        player = g_move(m, board, action, player, 0, &rng);
        st->plies++;
        g_ended(m, board, player, r);                             /* Coach.py:73 getGameEnded(board, curPlayer) */
        int any = 0; for (int p = 0; p < n; p++) any |= r[p] != 0.f;
        if (any || (max_plies > 0 && step >= max_plies)) break;
    }
    st->games++;
}

/* Coach.executeEpisode (Coach.py:37-84) replayed with INJECTED randomness (tests): every random input the reference consumes is
 * supplied by the caller, per ply p = episodeStep - 1 < P:
 *   u_full[p]      the playout-cap coin (MCTS.py:58)                 noise[p*noise_stride ..]  the root Dirichlet draw (MCTS.py:187-197), or NULL
 *   u_move[p]      the uniform of random_pick's np.random.choice (Coach.py:289-292: cdf = cumsum(p) / cdf[-1], searchsorted right)
 *   chance_seed[p] random_seed of the real move (non-zero => the deterministic draw of make_move; games without chance ignore it)
 * Records the UN-augmented examples (canonical board, pi = probs of getActionProb(temp=1), z = roll(r, -player), valids, q) of the
 * full-search plies (Coach.py:65-69,76-82). Returns the number of examples, or -1 if P plies did not finish the game. */
int azo_execute_episode_inj(azo_mcts* m, const i8* init_board, int P, const double* u_full, const double* u_move, const int64_t* chance_seed,
                            const double* noise, int noise_stride, double t0, double t1, double half, int max_examples,
                            i8* ex_board, float* ex_pi, float* ex_z, u8* ex_valid, float* ex_q, int* out_plies, int* out_actions, u8* out_full) {
    const int n = m->cfg.num_players, S = m->S, A = m->A;
    i8 board[MAXS], cb[MAXS]; memcpy(board, init_board, (size_t)S);
    int player = 0, step = 0, n_ex = 0; azo_mcts_reset(m);
    static __thread double probs[MAXA], w[MAXA]; float q[MAXP], r[MAXP];
    int* ex_player = (int*)malloc(sizeof(int) * (size_t)(max_examples > 0 ? max_examples : 1));
    azo_rng dummy; rng_seed(&dummy, 1);
    for (;;) {
        if (step >= P) { free(ex_player); *out_plies = step; return -1; }
        step++;
        memcpy(cb, board, (size_t)S); if (player) g_swap(m, cb, player);
        m->inj_u_full = u_full[step - 1];
        const int full = azo_mcts_get_action_prob(m, cb, 1.0, 0, noise ? noise + (size_t)(step - 1) * (size_t)noise_stride : NULL, probs, q, NULL);
        m->inj_u_full = -1.0;
        const double T = temp_for_selfplay(t0, t1, half, step);
        int action = -1;
        if (T == 0) { double b = -1; for (int a = 0; a < A; a++) if (probs[a] > b) { b = probs[a]; action = a; } }
        else {                                                    /* applyTemperatureAndNormalize + np.random.choice, Coach.py:278-292 */
            double s = 0; for (int a = 0; a < A; a++) { w[a] = pow(probs[a], 1.0 / T); s += w[a]; }
            double acc = 0; for (int a = 0; a < A; a++) { w[a] = w[a] / s; acc += w[a]; w[a] = acc; }
            const double last = w[A - 1], u = u_move[step - 1];
            action = 0; for (int a = 0; a < A; a++) if (w[a] / last <= u) action = a + 1;
            if (action >= A) action = A - 1;
        }
        if (full && n_ex < max_examples) {
            u8* V = ex_valid + (size_t)n_ex * (size_t)A; g_valid(m, cb, V);
            memcpy(ex_board + (size_t)n_ex * (size_t)S, cb, (size_t)S);
            for (int a = 0; a < A; a++) ex_pi[(size_t)n_ex * (size_t)A + a] = (float)probs[a];
            for (int p = 0; p < n; p++) ex_q[(size_t)n_ex * n + p] = q[p];
            ex_player[n_ex] = player; n_ex++;
        }
        if (out_actions) out_actions[step - 1] = action;
        if (out_full) out_full[step - 1] = (u8)full;
        player = g_move(m, board, action, player, chance_seed ? chance_seed[step - 1] : 1, &dummy);
        g_ended(m, board, player, r);
        int any = 0; for (int p = 0; p < n; p++) any |= r[p] != 0.f;
        if (any) break;
    }
    for (int e = 0; e < n_ex; e++) for (int p = 0; p < n; p++) ex_z[(size_t)e * n + p] = r[(p + ex_player[e]) % n];   /* np.roll(r, -player) */
    free(ex_player); *out_plies = step;
    return n_ex;
}

typedef struct { azo_cfg cfg; const float* blob; int games, max_plies; uint64_t seed; double t0, t1, half; azo_run_stats st; } worker_t;
static void* worker(void* arg) {
    worker_t* w = (worker_t*)arg;
    azo_mcts* m = azo_mcts_new(&w->cfg, w->blob, w->cfg.dirichletAlpha != 0, w->seed);
    for (int g = 0; g < w->games; g++) execute_episode(m, w->seed * 1000003ULL + (uint64_t)g, w->t0, w->t1, w->half, w->max_plies, &w->st);
    w->st.sims = m->n_sims; w->st.expansions = m->n_expansions; w->st.node_visits = m->n_node_visits; w->st.nn_evals = m->n_nn_evals;
    azo_mcts_free(m); return NULL;
}
/* CPU baseline: `threads` independent workers (the author's own scaling method is independent
 * processes, README.md:175-176), each playing `games_per_thread` self-play games. out[8]. */
void azo_selfplay_bench(const azo_cfg* cfg, const float* blob, int threads, int games_per_thread, int max_plies,
                        double t_begin, double t_end, double half_life, uint64_t seed, double* out) {
    worker_t* ws = (worker_t*)calloc((size_t)threads, sizeof(worker_t)); pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    struct timespec a, b; clock_gettime(CLOCK_MONOTONIC, &a);
    for (int i = 0; i < threads; i++) { ws[i].cfg = *cfg; ws[i].blob = blob; ws[i].games = games_per_thread; ws[i].max_plies = max_plies; ws[i].seed = seed + (uint64_t)i; ws[i].t0 = t_begin; ws[i].t1 = t_end; ws[i].half = half_life; pthread_create(&th[i], NULL, worker, &ws[i]); }
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    double secs = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
    for (int k = 0; k < 8; k++) out[k] = 0;
    for (int i = 0; i < threads; i++) { out[0] += (double)ws[i].st.sims; out[1] += (double)ws[i].st.expansions; out[2] += (double)ws[i].st.node_visits; out[3] += (double)ws[i].st.nn_evals; out[4] += (double)ws[i].st.plies; out[5] += (double)ws[i].st.examples; out[6] += (double)ws[i].st.games; }
    out[7] = secs; free(ws); free(th);
}

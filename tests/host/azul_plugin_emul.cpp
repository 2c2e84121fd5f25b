// Host emulation of the Azul plugin csrc/azul.cuh: the CUDA qualifiers are defined away and the 32 lanes of a warp
// function are run one after the other (none of the emulated functions exchanges data between lanes), so the device rules can be
// checked against the reference goldens without a GPU. Test infrastructure; built by tests/test_oracle_azul.py with g++.
#include <algorithm>
#include <cstdint>
#include <cstring>
#define AZG_HOST_EMUL 1
#define __device__
#define __forceinline__ inline
#define __constant__ static const
namespace azg {
constexpr unsigned FULL = 0xFFFFFFFFu;
using std::min;
struct Philox { uint64_t s; float uniformf() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (float)((s >> 40) & 0xFFFFFF) / 16777216.0f; } };
static inline unsigned __ballot_sync(unsigned, bool) { return 0u; }     // valid_mask is not emulated (action_valid is called per action instead)
static inline void __syncwarp() {}
}  // namespace azg
#include "../../alpha-zero-general_b200/csrc/azul.cuh"
using azg::Azul;

extern "C" {
int emul_sizes(int* out) { out[0] = Azul::S; out[1] = Azul::SP; out[2] = Azul::A; out[3] = Azul::MAX_SYM; out[4] = Azul::NP; return 0; }
void emul_valid(const int8_t* board, int player, uint8_t* out) {
    int8_t b[Azul::SP] = {0}; memcpy(b, board, Azul::S);
    for (int a = 0; a < Azul::A; a++) out[a] = Azul::action_valid(b, a, player) ? 1 : 0;
}
int emul_make_move(int8_t* board, int move, int player, long long seed) {
    int8_t b[Azul::SP] = {0}; memcpy(b, board, Azul::S);
    azg::Philox rng{12345};
    const int np = Azul::make_move(b, move, player, seed, &rng);
    memcpy(board, b, Azul::S);
    return np;
}
int emul_ended(const int8_t* board, float* out) {
    int8_t b[Azul::SP] = {0}; memcpy(b, board, Azul::S);
    float es[Azul::NP]; const bool over = Azul::ended(b, 0, es, 0);
    out[0] = es[0]; out[1] = es[1];
    return over ? 1 : 0;
}
void emul_swap(int8_t* board, int nb_swaps) {
    int8_t b[Azul::SP] = {0}; memcpy(b, board, Azul::S);
    for (int lane = 0; lane < 32; lane++) Azul::swap_players(b, nb_swaps, lane);
    memcpy(board, b, Azul::S);
}
int emul_round(const int8_t* board) { return Azul::round(board); }
int emul_score(const int8_t* board, int player) { return Azul::score(board, player); }
int emul_symmetries(const int8_t* board, const float* pi, const uint8_t* mask, int8_t* ob, float* opi, uint8_t* om) {
    int8_t b[Azul::SP] = {0}; memcpy(b, board, Azul::S);
    const int n = Azul::num_symmetries(b);
    for (int k = 0; k < n; k++)
        for (int lane = 0; lane < 32; lane++)
            Azul::symmetry(b, pi, mask, k, lane, ob + (size_t)k * Azul::S, opi + (size_t)k * Azul::A, om + (size_t)k * Azul::A);
    return n;
}
void emul_init(int8_t* board, uint64_t seed) {
    int8_t b[Azul::SP]; azg::Philox rng{seed};
    Azul::init_game(b, &rng);
    memcpy(board, b, Azul::S);
}
}

// Compile-only check of the round-2 draft plugin csrc/next/azul.cuh (it is not part of the library yet): instantiates every member in a
// kernel so that nvcc type-checks and generates sm_100a code for it. Nothing is launched; the host part prints the plugin's sizes.
#include <cstdio>
#include "../../alpha-zero-general_b200/csrc/next/azul.cuh"
using namespace azg;
__global__ void k_touch(int8_t* boards, float* pi, uint8_t* mask, int8_t* ob, float* opi, uint8_t* om, int* out) {
    __shared__ int8_t b[Azul::SP];
    __shared__ uint32_t w[Azul::MASK_WORDS];
    const int lane = threadIdx.x & 31;
    Philox rng(1, 2, 3);
    if (lane == 0) Azul::init_game(b, &rng);
    __syncwarp();
    Azul::valid_mask(b, 0, lane, w);
    int np = 0;
    if (lane == 0) np = Azul::make_move(b, __ffs(w[0]) - 1, 0, 31416, &rng);
    np = __shfl_sync(FULL, np, 0);
    __syncwarp();
    Azul::swap_players(b, np, lane);
    float es[Azul::NP];
    const bool over = Azul::ended(b, 0, es, lane);
    for (int k = 0; k < Azul::num_symmetries(b); k++) Azul::symmetry(b, pi, mask, k, lane, ob + k * Azul::S, opi + k * Azul::A, om + k * Azul::A);
    if (lane == 0) { out[0] = Azul::round(b) + Azul::score(b, 0) + (over ? 1 : 0) + (Azul::is_chance_move(0) ? 1 : 0); for (int i = 0; i < Azul::S; i++) boards[i] = b[i]; }
}
int main() {
    printf("azul plugin: S=%d SP=%d A=%d MASK_WORDS=%d MAX_SYM=%d\n", Azul::S, Azul::SP, Azul::A, Azul::MASK_WORDS, Azul::MAX_SYM);
    return (void*)k_touch == nullptr;
}

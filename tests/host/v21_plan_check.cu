// Host-only check of the tile plan of the AbaloneNNet V21 kernel (net_v21.cuh: v21_plan). No kernel is launched.
// Prints "ok" when, for every batch size and SM count tried, the plan covers all leaves, never over-provisions by more than one
// tile, and uses 2-leaf tiles only when they fit in a single round of CTAs.
#include <cstdio>
#include "../../alpha-zero-general_b200/csrc/net_v21.cuh"
int main() {
    using namespace azg;
    const int sms[] = {1, 7, 108, 132, 148, 160};
    for (int n_sm : sms)
        for (int n = 1; n <= 20000; n++) {
            const V21Plan p = v21_plan(n, n_sm);
            const long cover = 4L * p.n_big + 2L * p.n_small;
            if (p.n_big < 0 || p.n_small < 0 || cover < n) { printf("FAIL cover n=%d sm=%d big=%d small=%d\n", n, n_sm, p.n_big, p.n_small); return 1; }
            if (cover - n >= 4 || (p.n_small > 0 && cover - n >= 2)) { printf("FAIL waste n=%d sm=%d big=%d small=%d\n", n, n_sm, p.n_big, p.n_small); return 1; }
            if (p.n_small > n_sm) { printf("FAIL small tiles exceed one round n=%d sm=%d small=%d\n", n, n_sm, p.n_small); return 1; }
            if (p.n_small > 0 && p.n_big % n_sm != 0) { printf("FAIL small tiles after a partial round n=%d sm=%d big=%d\n", n, n_sm, p.n_big); return 1; }
        }
    const V21Plan b = v21_plan(2048, 148);              // the bench shard: 3 full rounds of 4-leaf tiles + 136 2-leaf tiles
    if (b.n_big != 444 || b.n_small != 136) { printf("FAIL bench plan %d %d\n", b.n_big, b.n_small); return 1; }
    if (v21_smem_bytes() > 227 * 1024) { printf("FAIL smem %zu\n", v21_smem_bytes()); return 1; }
    printf("ok\n");
    return 0;
}

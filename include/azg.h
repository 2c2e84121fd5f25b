/*
 * azg.h -- C ABI of the B200-native AlphaZero self-play engine (libazg_b200.so).
 *
 * The reference (cestpasphoto/alpha-zero-general) is pure Python: its "FFI" for the hot path is
 * the duck-typed plugin surface Game.py / NeuralNet.py / MCTS.py / Coach.py. Each entry point below
 * names the reference interface it replaces (file:line relative to the reference root). A ctypes
 * binding of exactly these symbols is what alpha-zero-general_b200/lib.py loads, and what a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; azg_last_error() gives the message
 *    (thread-local).
 *  - all buffers are caller-owned. A pointer may be HOST memory (numpy) or DEVICE memory
 *    (torch.Tensor.data_ptr()): the library detects which (cudaPointerGetAttributes) and stages host
 *    buffers through its own device scratch, synchronising before it returns. With device pointers the
 *    work is only enqueued on `stream` (a cudaStream_t passed as void*, NULL = default stream).
 *  - boards are int8, row-major, `state_bytes` bytes per board, tightly packed [n][state_bytes].
 *  - masks are uint8 0/1 [n][action_size]; policies float32 [n][action_size]; values float32 [n][num_players].
 *  - handles are not thread-safe; one engine per GPU / process.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef AZG_H
#define AZG_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AZG_ABI_VERSION 4
#define AZG_N_STATS 20                                  /* entries azg_engine_stats writes */

enum { AZG_GAME_SPLENDOR = 1,                            /* GameSwitcher.py:3-13 ('splendor'), 2 players            */
       AZG_GAME_SANTORINI = 2,                           /* 'santorini' built with NB_GODS = 1 (SantoriniConstants.py:19) */
       AZG_GAME_ABALONE = 3,                             /* 'abalone', Belgian daisy start (AbaloneLogicNumba.py:5-6)      */
       AZG_GAME_AZUL = 4 };                              /* 'azul', 2 players (azul/AzulLogicNumba.py)                     */
enum { AZG_NET_HASH = 0,                                 /* deterministic test net (tests only)                       */
       AZG_NET_SPLENDOR_V80 = 80,                        /* splendor/SplendorNNet.py version 80 (shipped 2-player net) */
       AZG_NET_SANTORINI_V89 = 89,                       /* santorini/SantoriniNNet.py version 89 (shipped no-god net) */
       AZG_NET_ABALONE_V21 = 21,                         /* abalone/AbaloneNNet.py version 21 (shipped Belgian-daisy net) */
       AZG_NET_AZUL_V84 = 84 };                          /* azul/AzulNNet.py version 84 (shipped 2-player net)            */

typedef struct {
    int32_t game_id, num_players;
    int32_t state_rows, state_cols, state_depth;         /* Game.getBoardSize   Game.py:22 (depth 1 = 2-D board) */
    int32_t state_bytes;
    int32_t action_size;                                 /* Game.getActionSize  Game.py:29  */
    int32_t max_symmetries;                              /* upper bound of len(getSymmetries()) */
    int32_t max_game_len;                                /* upper bound on plies per game */
} azg_game_info_t;

int azg_abi_version(void);
const char* azg_last_error(void);
int azg_device_count(void);                               /* 0 when no CUDA device is usable */
int azg_set_device(int device);                           /* device used by handles created afterwards on this thread */
int azg_game_info(int game_id, int num_players, azg_game_info_t* out);

/* ---- batched game step: the *LogicNumba.Board methods behind the Game facade ------------------- */
/* Game.getInitBoard (Game.py:14; splendor/SplendorLogicNumba.py:151-178). One board per seed; the draws
 * come from a counter-based device RNG keyed by the seed (the reference uses numba's MT19937 stream). */
int azg_game_init(int game_id, int num_players, int n, const uint64_t* seeds, int8_t* boards, void* stream);
/* Game.getValidMoves (Game.py:51; SplendorLogicNumba.py:180-188). players may be NULL (= all 0). */
int azg_game_valid(int game_id, int num_players, int n, const int8_t* boards, const int32_t* players,
                   uint8_t* mask, void* stream);
/* Game.getNextState (Game.py:36-49; SplendorLogicNumba.py:190-205,306-357). seeds[i]==0 means a true
 * random chance draw taken from the device RNG keyed by rng_keys[i] (may be NULL => key i). */
int azg_game_next(int game_id, int num_players, int n, const int8_t* boards, const int32_t* players,
                  const int32_t* actions, const int64_t* seeds, const uint64_t* rng_keys,
                  int8_t* out_boards, int32_t* out_next_player, void* stream);
/* Game.getGameEnded(board, next_player) (Game.py:64; SplendorLogicNumba.py:221-240; SantoriniLogicNumba.py:552-565):
 * float32[n][num_players]. next_players = the player to move on each board (NULL = all 0; Splendor ignores it). */
int azg_game_ended(int game_id, int num_players, int n, const int8_t* boards, const int32_t* next_players, float* out, void* stream);
/* Game.getCanonicalForm (Game.py:97; SplendorLogicNumba.py:244-253). */
int azg_game_canonical(int game_id, int num_players, int n, const int8_t* boards, const int32_t* players,
                       int8_t* out_boards, void* stream);
/* Game.getRound / Game.getScore (Game.py:76,87; SplendorLogicNumba.py:151-154,303-304).
 * rounds int32[n]; scores int32[n][num_players]; either output may be NULL. */
int azg_game_round_score(int game_id, int num_players, int n, const int8_t* boards, int32_t* rounds,
                         int32_t* scores, void* stream);
/* Game.getSymmetries (Game.py:113; SplendorLogicNumba.py:255-301). Outputs are [n][max_symmetries][..];
 * out_k[i] = number of valid entries for board i. */
int azg_game_symmetries(int game_id, int num_players, int n, const int8_t* boards, const float* pi,
                        const uint8_t* mask, int8_t* out_boards, float* out_pi, uint8_t* out_mask,
                        int32_t* out_k, void* stream);

/* ---- policy/value net: GenericNNetWrapper.predict / predict_server (GenericNNetWrapper.py:94-157) ---- */
typedef struct azg_net azg_net;
/* weights: float32 blob = the net's state_dict tensors concatenated in the order documented in
 * alpha-zero-general_b200/nnet.py (V80_TENSOR_ORDER / V89_TENSOR_ORDER / V21_TENSOR_ORDER / V84_TENSOR_ORDER); host or device pointer. NULL for AZG_NET_HASH. */
int azg_net_create(int net_kind, int game_id, int num_players, const float* weights, size_t n_weights, azg_net** out);
int azg_net_load(azg_net* net, const float* weights, size_t n_weights);      /* new weights, same architecture */
/* pi = softmax over legal actions (what predict returns after np.exp), v = tanh value vector. */
int azg_net_forward(azg_net* net, int n, const int8_t* boards, const uint8_t* mask, float* pi, float* v, void* stream);
int azg_net_destroy(azg_net* net);

/* ---- search engine: MCTS.py:19-261 for n_games concurrent, independent trees ---------------------- */
typedef struct {
    int32_t game_id, num_players;
    int32_t n_games;                 /* concurrent trees (= threads of Coach.executeEpisodes_batch, Coach.py:86) */
    int32_t numMCTSSims;             /* main.py:125 */
    int32_t ratio_fullMCTS;          /* main.py:131 */
    int32_t universes;               /* main.py:133; 0 => seed -1 */
    int32_t forced_playouts;         /* main.py:132 */
    int32_t no_mem_optim;            /* main.py:153 (tree GC is a semantic no-op; kept for flag parity) */
    int32_t dirichlet_noise;         /* MCTS(..., dirichlet_noise=) MCTS.py:24; Coach passes dirichletAlpha!=0 */
    int32_t node_cap, edge_cap;      /* per-game arena sizes; 0 => defaults derived from numMCTSSims */
    double cpuct, fpu, dirichletAlpha, prob_fullMCTS;      /* main.py:126-130 */
    double temperature[3];           /* main.py:135 ; [2] is the root-prior softmax temperature */
    double tempThreshold;            /* main.py:136 */
    uint64_t seed;                   /* device RNG seed (Dirichlet, PCR coin flips, move sampling, chance) */
    uint64_t first_game;             /* global id of slot 0. Every RNG stream is keyed by (seed, first_game + slot, games started in the
                                        slot, ply), so what a slot plays does not depend on how slots are sharded over ranks: rank r of a
                                        world that shards n_total slots passes first_game = dist.shard_games(n_total, r, world)[0]. */
} azg_engine_cfg;

typedef struct azg_engine azg_engine;
int azg_engine_create(const azg_engine_cfg* cfg, azg_net* net, azg_engine** out);
int azg_engine_destroy(azg_engine* e);
/* MCTS.reset_all_search_trees (MCTS.py:199-203) for all games, or one game slot if game >= 0. */
int azg_engine_reset(azg_engine* e, int game);

/* MCTS.getActionProb (MCTS.py:49-103) for games [0,n): n_sims simulations each from roots[i] (canonical
 * boards). Trees persist across calls (tree reuse) until azg_engine_reset.
 *   full_search  uint8[n] or NULL (=all full): 1 selects numMCTSSims, 0 numMCTSSims/ratio; Dirichlet at sim 0,
 *                forced playouts and policy-target pruning exactly as MCTS.py:58-65,75-80. 2 = no search for that slot this call
 *                (its tree is kept, its outputs are whatever the tree already holds for the root, normally all zero): lets a
 *                caller that owns several engines search each position with the engine whose turn it is (Arena.py:75-76).
 *   noise        float64[n][action_size] or NULL: injected Dirichlet draws (first L entries used, L = number
 *                of legal actions); NULL => drawn on device.
 *   out_counts   int32[n][action_size]  root Nsa after policy-target pruning (MCTS.py:75-80)
 *   out_raw      int32[n][action_size]  root Nsa before pruning (may be NULL)
 *   out_q        float32[n][num_players] (MCTS.py:71-72) (may be NULL)
 * Policies are counts/sum(counts) (temp=1); other temperatures are a host-side power of the counts. */
int azg_engine_search(azg_engine* e, int n, const int8_t* roots, const uint8_t* full_search, const double* noise,
                      int32_t* out_counts, int32_t* out_raw, float* out_q, void* stream);

/* Coach.executeEpisodes (Coach.py:86-148): keeps all n_games slots playing (refilling finished games)
 * until at least `min_episodes` games have finished or `max_moves` lock-step plies were played, or (when min_episodes > 0) the
 * example ring is more than half full -- drain it with azg_engine_examples and call again (stats[10] counts finished episodes).
 * Training examples (un-augmented: one per full-search ply) are appended to the engine's example ring
 * and fetched with azg_engine_examples. Returns counters in out_stats (see azg_engine_stats). */
int azg_engine_selfplay(azg_engine* e, int min_episodes, int max_moves, void* stream);
/* Injected randomness for azg_engine_selfplay (parity tests: replay of an episode recorded from the reference's Coach.executeEpisode,
 * oracle/gen_golden_selfplay.py). Per slot g and ply p = episodeStep - 1 < n_plies the engine takes
 *   u_full[g][p]      instead of MCTS.rng.random()  -- the playout-cap coin, MCTS.py:58
 *   noise[g][p][0..L) instead of MCTS.rng.dirichlet -- the root Dirichlet draw of a full search, MCTS.py:187-197 (NULL: drawn on device)
 *   u_move[g][p]      instead of the uniform np.random.choice consumes in random_pick, Coach.py:289-292
 *   chance_seed[g][p] as random_seed of the real move (Coach.py:71 passes 0 = true random; non-zero = make_move's deterministic draw)
 * and starts slot g from init_boards[g] instead of Game.getInitBoard. Every slot then plays exactly ONE game and idles (no refill).
 * Buffers (host or device) are copied; inj == NULL returns the engine to its own device RNG. Resets all self-play slots. */
typedef struct {
    int32_t n_plies;
    const int8_t* init_boards;       /* [n_games][state_bytes] */
    const double* u_full;            /* [n_games][n_plies] */
    const double* u_move;            /* [n_games][n_plies] */
    const int64_t* chance_seed;      /* [n_games][n_plies] */
    const double* noise;             /* [n_games][n_plies][action_size] or NULL */
} azg_selfplay_inject;
int azg_engine_selfplay_inject(azg_engine* e, const azg_selfplay_inject* inj);
/* Drains up to `cap` finished-game examples: boards int8[cap][S], pi f32[cap][A], z f32[cap][np],
 * valids u8[cap][A], q f32[cap][np]; *out_n = number written. (tuple layout of Coach.py:76-82). The buffers may be DEVICE memory
 * (torch tensors): the examples then never leave HBM (device-to-device copies), which is what the multi-GPU gather and a training
 * step on the same GPU use. */
int azg_engine_examples_pending(azg_engine* e, int32_t* out_n);     /* examples waiting in the ring (size the buffers of azg_engine_examples) */
int azg_engine_examples(azg_engine* e, int cap, int8_t* boards, float* pi, float* z, uint8_t* valids, float* q, int32_t* out_n);

/* MCTS.nodes_data[stringRepresentation(board)] (MCTS.py:37-39: the tuple (Es, Vs, Ps, Ns, Qsa, Nsa, r, Qs)) for n boards, query i looked
 * up in the tree of slot slots[i] (NULL = slot i). found int32[n]: 0 = not in the tree, 1 = expanded node, 2 = terminal node (es / round
 * only). Dense A-wide rows like the reference's arrays: vs u8[n][A], ps f32[n][A], qsa f64[n][A] (-42 = never visited), nsa i32[n][A];
 * es f32[n][np], ns i32[n], round i32[n], qs f32[n]. Any output but `found` may be NULL. */
int azg_engine_node(azg_engine* e, int n, const int32_t* slots, const int8_t* boards, int32_t* found, float* es, uint8_t* vs, float* ps,
                    int32_t* ns, double* qsa, int32_t* nsa, int32_t* round, float* qs, void* stream);
/* The self-play slots as Coach.executeEpisode's locals (Coach.py:55-60): absolute-frame boards int8[n_games][S], curPlayer, episodeStep,
 * and whether the slot holds a running game. Any output may be NULL. */
int azg_engine_selfplay_state(azg_engine* e, int8_t* boards, int32_t* players, int32_t* plies, int32_t* active);

/* Counters since creation: [0] sims [1] node_visits (select steps) [2] expansions (nodes with priors)
 * [3] nn_evals [4] terminal_hits [5] arena_overflows [6] gc_runs [7] max_nodes_in_a_tree
 * [8] sum_legal (over expansions) [9] moves_played [10] episodes_finished [11] examples_recorded
 * [12] kernels_launched [13] gc_sweeps (tier-2 reachability GCs, see tree.cuh) [14] node_cap [15] sum_legal_visited (sum of n_legal over select steps)
 * [16] sum_legal_root_scans (edges scanned by k_select at the roots) [17] sum_legal_refreshed (edges scanned by k_backup when it
 * refreshes the cached PUCT choice of the nodes on the path) [18] examples_dropped (example ring full: must stay 0)
 * [19] gc_trims (tier-3 GCs: the reused tree itself was too large for the arena and lost its deepest nodes; raise node_cap to avoid).  out must hold AZG_N_STATS int64. */
int azg_engine_stats(azg_engine* e, int64_t* out_stats);

/* Per-kernel device timing (CUDA events on the launching stream around every launch of the search loop).
 * enable=1 starts collecting for every launch (6 event records per simulation: ~4 % slower loop), enable=N>1 times only every N-th
 * lock-step simulation (sampled: negligible overhead; the per-move kernels are always timed), enable=0 stops.
 * azg_engine_kernel_times drains the events: out8 = [0] select ms [1] leaf-eval (net) ms [2] expand+backup ms
 * [3] other (gc, move begin/end, finish) ms [4] profiled lock-step simulations [5] select launches
 * [6] net launches [7] backup launches. */
int azg_engine_profile(azg_engine* e, int enable);
int azg_engine_kernel_times(azg_engine* e, double* out8);

/* ---- debug / profiling hooks (no reference counterpart; used by scripts/dbg_*.py) ----
 * azg_net_prof: SM-clock timestamps of the phases of the first tiles CTA 0 of the V80 tensor-core kernel processed (nets created with
 *   AZG_V80_PROF=1 in the environment); out64 receives 64 int64. azg_net_prof_ctas: the same nets, last launch: per CTA (160 x 4 int64)
 *   {kernel entry, prologue done, exit} in globaltimer ns and the SM id.
 * azg_debug_selprof: per-phase cycle sums of k_select / k_backup, only filled by a library built with -DAZG_SEL_PROF=1|2; out16
 *   receives 16 uint64.
 * Environment switches read at handle creation: AZG_V80_KERNEL=fp32 (CUDA-core V80 forward instead of the tcgen05 kernel, for A/B
 * runs), AZG_TREE_REPLAY=0 (k_select without path replay; results are identical). */
int azg_net_prof(azg_net* net, long long* out64);
int azg_net_prof_ctas(azg_net* net, long long* out640);
int azg_debug_selprof(unsigned long long* out16);

#ifdef __cplusplus
}
#endif
#endif

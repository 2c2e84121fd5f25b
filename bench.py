#!/usr/bin/env python
"""bench.py -- self-play MCTS sims/sec of the B200 engine (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's engine
  python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU arm: the oracle port of the reference path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)

Workload (config.workload): BASELINE.json configs[2] -- Splendor 2-player with chance-node card draws
(universes=3), 16384 concurrent self-play games per GPU, numMCTSSims=800, every move a full search, SplendorNNet
V80 random-init (seed 0), root Dirichlet noise on.  One STEP = one self-play ply of every game = 800 lock-step
simulations x n_games trees (select -> batched V80 forward -> expand+backup), plus the move itself.

  value  : device-resident arm. azg_engine_selfplay plays W warm-up plies then K timed plies entirely on the GPU
           (Coach.executeEpisodes equivalent); sims counted by the engine's own counters; CUDA events on the
           launching stream; max over ranks.
  e2e    : the same K plies driven through the reference-facing plugin calls with HOST buffers
           (MCTS.getActionProb-> azg_engine_search, Game.getNextState/getGameEnded/getCanonicalForm ->
           azg_game_*), pinned host memory, every host<->device copy inside the timed region.
  roofline : dominant kernel by device time (CUDA events around every launch inside the timed region).
  cpu_baseline : oracle port (oracle/azg_oracle.c) of the same path on all host threads, bounded sample.

The working set (trees: tens of GB) is far larger than L2 (126 MB), so no explicit L2 flush is needed (config.l2).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'selfplay_mcts_sims_per_sec'
UNIT = 'sims/s'
N_PL = 2
# per-game workload constants: board bytes, actions, multiply-adds of one leaf evaluation (SURVEY.md section 8a row a17:
# V80 1.04 MFLOP, V89 18.52 MFLOP), default concurrent games (BASELINE.json configs[2] / configs[1]), universes
GAMES = {
    'splendor': dict(S=392, A=81, flops=2 * 521205, games=16384, universes=3, net='SplendorNNet V80 (142406 params, random init seed 0)',
                     tag='splendor2p_chance_universes3'),
    'santorini': dict(S=75, A=162, flops=2 * 9259428, games=4096, universes=1, node_cap_per_sim=16, net='SantoriniNNet V89 (381454 params, random init seed 0)',
                      tag='santorini_nogods'),
    'abalone': dict(S=324, A=3402, flops=2 * 1050360, games=2048, universes=1, net='AbaloneNNet V21 (35862 params, random init seed 0)',
                    tag='abalone_belgian_daisy', sims=1600),
    'azul': dict(S=138, A=180, flops=2 * 203708, games=8192, universes=2, node_cap_per_sim=12, net='AzulNNet V84 (118 k params, random init seed 0)', tag='azul2p_universes2'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--game', default='splendor', choices=sorted(GAMES), help='splendor = BASELINE.json configs[2] (headline metric), santorini = configs[1], abalone = configs[4] (2048 games per GPU), azul = the first SURVEY 8f game')
    ap.add_argument('--games', type=int, default=0, help='concurrent games per GPU (0 = the config default: 16384 splendor / 4096 santorini / 2048 abalone / 8192 azul)')
    ap.add_argument('--sims', type=int, default=0, help='numMCTSSims (0 = the config default: 800, abalone 1600)')
    ap.add_argument('--node-cap', type=int, default=0, help='nodes per tree arena (0 = 6 x sims + 320; santorini 16 x, azul 12 x)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-pcr', action='store_true', help='skip the secondary line with the reference-default playout-cap randomisation (prob_fullMCTS 0.25, ratio 5)')
    ap.add_argument('--no-iteration', action='store_true', help='skip the secondary whole-iteration leg (complete games -> example drain -> NCCL gather -> symmetries)')
    ap.add_argument('--iter-sims', type=int, default=0, help='numMCTSSims of the whole-iteration leg (0 = 40: complete games within seconds)')
    ap.add_argument('--iter-games', type=int, default=0, help='concurrent games per GPU of the whole-iteration leg (0 = --games)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-plies', type=int, default=0, help='plies per thread of the cpu_baseline sample (0 = 12 splendor / 1 santorini: ~10-30 s)')
    a = ap.parse_args()
    a.games = a.games or GAMES[a.game]['games']
    a.sims = a.sims or GAMES[a.game].get('sims', 800)
    a.cpu_plies = a.cpu_plies or {'splendor': 12, 'santorini': 1, 'abalone': 2, 'azul': 12}[a.game]
    return a


def mcts_args(sims, game='splendor'):
    # main.py defaults; config C3 (splendor): universes=3; every move a full search
    return dict(numMCTSSims=sims, cpuct=1.25, fpu=0.0, universes=GAMES[game]['universes'], dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                tempThreshold=10, prob_fullMCTS=1.0, ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d['hbm_gbs']), bf16=float(d['bf16_tflops']), bf16_sustained=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])),
                    src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src='fallback (B200_PROFILING.md)')


# --------------------------------------------------------------------------- clocks --------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ''
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm), reasons=sorted(reasons))


# --------------------------------------------------------------------------- reference (CPU) arm -------
def cpu_sample(sims, plies, threads, seed=1, game='splendor'):
    """Oracle port of Coach.executeEpisode / MCTS.search / the game's Board / the net forward on `threads` host threads,
    each playing one self-play game truncated after `plies` plies. Returns the oracle's counters."""
    from oracle import oracle as O
    from azg_b200.nnet import random_v80_state_dict, random_v89_state_dict, random_v21_state_dict, random_v84_state_dict
    kind = {'splendor': 1, 'santorini': 2, 'abalone': 3, 'azul': 4}[game]; gid = {'splendor': O.GAME_SPLENDOR, 'santorini': O.GAME_SANTORINI, 'abalone': O.GAME_ABALONE, 'azul': O.GAME_AZUL}[game]
    a = mcts_args(sims, game)
    cfg = O.make_cfg(numMCTSSims=sims, net_kind=kind, universes=a['universes'], prob_fullMCTS=1.0, cpuct=a['cpuct'], fpu=a['fpu'],
                     dirichletAlpha=a['dirichletAlpha'], temperature2=a['temperature'][2], game=gid)
    blob = {'splendor': lambda: O.v80_blob(random_v80_state_dict(0)), 'santorini': lambda: O.v89_blob(random_v89_state_dict(0)),
            'abalone': lambda: O.v21_blob(random_v21_state_dict(0)), 'azul': lambda: O.v84_blob(random_v84_state_dict(0))}[game]()
    return O.selfplay_bench(cfg, blob, threads, 1, max_plies=plies, temperature=a['temperature'][:2], tempThreshold=a['tempThreshold'], seed=seed)


def port_vs_reference(port_value):
    """The C port is faster than the reference's own Numba + torch-CPU path. profiles/r02_reference_cpu.json holds both timed on the
    same cores of the build container (scripts/time_reference_cpu.py runs the UNMODIFIED Coach.executeEpisode); the measured
    port/reference factor converts a port figure into what the reference itself would do on these cores."""
    p = os.path.join(ROOT, 'profiles', 'r02_reference_cpu.json')
    if not os.path.exists(p):
        return {}
    d = json.load(open(p)); f = float(d['port_over_reference']['all_cores'])
    return {'port_over_reference': f, 'reference_equivalent_value': port_value / f,
            'reference_measured': {'where': d['where'], 'cpu_model': d['cpu_model'], 'one_process_sims_per_s': d['ref_1proc']['sims_per_s_sum'],
                                   'P_processes': d['ref_Pproc']['procs'], 'P_processes_sims_per_s': d['ref_Pproc']['sims_per_s_sum'],
                                   'stub_net_one_process_sims_per_s': d['ref_stub_1proc']['sims_per_s_sum'], 'source': 'profiles/r02_reference_cpu.json'}}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    plies = 2                                               # one step = every host thread plays 2 plies (2 x sims sims)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(args.sims, 1, threads, game=args.game)
    t0 = time.perf_counter(); sims = 0
    for k in range(args.steps):
        r = cpu_sample(args.sims, plies, threads, seed=100 + k, game=args.game); sims += r['sims']
    dt = time.perf_counter() - t0
    val = sims / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32/f64',
            'data': 'synthetic', 'config': workload_cfg(args, threads),
            'cpu_baseline': dict({'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': f'{threads} host threads x 1 {args.game} self-play game x {plies} plies x {args.sims} sims per step, {args.steps} steps; '
                                       'oracle/azg_oracle.c (C port of the reference path; the Python/numba reference cannot travel to the GPU box)'},
                                 **port_vs_reference(val)),
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def workload_cfg(args, cpu_threads=None):
    gm = GAMES[args.game]
    c = {'workload': f'{gm["tag"]}_{args.games}games_per_gpu_{args.sims}sims_randinit', 'game': args.game, 'num_players': 2,
         'games_per_gpu': args.games, 'numMCTSSims': args.sims, 'net': gm['net'],
         'universes': gm['universes'], 'prob_fullMCTS': 1.0, 'dirichlet_noise': True, 'step': 'one self-play ply of every game (numMCTSSims lock-step simulations per tree)',
         'parallelism': f'games sharded over {args.gpus} GPU(s) by global slot id; no collective inside the search loop; the iteration leg gathers the examples over NCCL', 'l2': 'working set >> L2 (tree arenas of tens of GB); no flush needed'}
    if cpu_threads:
        c['cpu_threads'] = cpu_threads
    return c


# --------------------------------------------------------------------------- e2e (host-buffer) arm -----
class HostLoop:
    """Coach.executeEpisode written against the plugin calls, batched over all games, with pinned HOST buffers."""

    def __init__(self, torch, game, eng, n, seed):
        import azg_b200.lib as lib
        self.lib = lib; self.L = lib.load(); self.game = game; self.eng = eng; self.n = n
        S_BYTES, N_ACT = game.info.state_bytes, game.info.action_size
        self.N_ACT = N_ACT
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        self.board = pin((n, S_BYTES), torch.int8); self.board2 = pin((n, S_BYTES), torch.int8); self.roots = pin((n, S_BYTES), torch.int8)
        self.player = pin((n,), torch.int32); self.player2 = pin((n,), torch.int32); self.action = pin((n,), torch.int32)
        self.seeds = pin((n,), torch.int64); self.keys = pin((n,), torch.int64).view(np.uint64); self.ended = pin((n, N_PL), torch.float32)
        self.counts = pin((n, N_ACT), torch.int32); self.q = pin((n, N_PL), torch.float32)
        self.rng = np.random.default_rng(seed); self.key_ctr = 1 << 20
        self.board[:] = game.init_batch(np.arange(1, n + 1, dtype=np.uint64) + (seed << 32)).reshape(n, S_BYTES)
        self.player[:] = 0; self.seeds[:] = 0
        self.roots[:] = self.board
        self.h2d = 0; self.d2h = 0

    def step(self):
        lib, L, g, n = self.lib, self.L, self.game, self.n
        p = lib.ptr; N_ACT = self.N_ACT
        # MCTS.getActionProb for every game (host roots in, host counts out)
        lib.check(L.azg_engine_search(self.eng.h, n, p(self.roots), None, None, p(self.counts), None, p(self.q), None))   # getActionProb returns probs and q; the un-pruned counts are not asked for
        self.h2d += self.roots.nbytes; self.d2h += self.counts.nbytes + self.q.nbytes
        # Coach.py:63 random_pick with temp_for_selfplay ~ 1 (early plies): sample from the visit counts
        if N_ACT <= 512:
            c = self.counts.astype(np.float64); cs = np.cumsum(c, axis=1); u = self.rng.random(n) * cs[:, -1]
            self.action[:] = np.minimum((cs <= u[:, None]).sum(axis=1), N_ACT - 1)
        else:   # large action spaces: over the NON-ZERO counts only (a dense cumsum over n x 3402 Abalone actions cost 9 % of that game's e2e step)
            r, a = np.nonzero(self.counts)                          # row-major: the visited actions of game 0, then game 1, ...
            cs = np.cumsum(self.counts[r, a], dtype=np.int64)
            last = np.searchsorted(r, np.arange(n), side='right') - 1   # index of every game's last visited action
            end = cs[last]; start = np.concatenate(([0], end[:-1]))
            u = start + np.floor(self.rng.random(n) * (end - start)).astype(np.int64)
            self.action[:] = a[np.minimum(np.searchsorted(cs, u, side='right'), last)]
        # Game.getNextState with a true random chance draw (random_seed=0), Coach.py:71
        self.keys[:] = np.arange(self.key_ctr, self.key_ctr + n, dtype=np.uint64); self.key_ctr += n
        lib.check(L.azg_game_next(g.game_id, N_PL, n, p(self.board), p(self.player), p(self.action), p(self.seeds), p(self.keys),
                                  p(self.board2), p(self.player2), None))
        self.h2d += self.board.nbytes + self.player.nbytes + self.action.nbytes + self.seeds.nbytes + self.keys.nbytes
        self.d2h += self.board2.nbytes + self.player2.nbytes
        self.board, self.board2 = self.board2, self.board; self.player, self.player2 = self.player2, self.player
        # Game.getGameEnded, Coach.py:73
        lib.check(L.azg_game_ended(g.game_id, N_PL, n, p(self.board), p(self.player), p(self.ended), None))
        self.h2d += self.board.nbytes + self.player.nbytes; self.d2h += self.ended.nbytes
        done = np.flatnonzero(self.ended.any(axis=1))
        for i in done:                                              # finished game: new game + fresh tree in that slot (Coach.py:93-98)
            self.board[i] = g.init_batch(np.array([self.key_ctr + int(i)], dtype=np.uint64)).reshape(-1); self.player[i] = 0
            self.eng.reset(int(i))
        # Game.getCanonicalForm, Coach.py:61
        lib.check(L.azg_game_canonical(g.game_id, N_PL, n, p(self.board), p(self.player), p(self.roots), None))
        self.h2d += self.board.nbytes + self.player.nbytes; self.d2h += self.roots.nbytes


# --------------------------------------------------------------------------- whole-iteration leg -------
def run_iteration(torch, dist, args, game, net, rank, world, dev, barrier, max_over_ranks, sum_over_ranks, gather_examples, shard_games):
    """One self-play ITERATION as the reference runs it (Coach.py:105-148 executeEpisodes, then learn() consumes the examples in one
    process): every slot plays complete games until n_games episodes have finished on this rank (refill of finished slots, terminal
    nodes, example hand-over to the ring, ring drained device-to-device whenever it is half full), then the un-augmented examples
    of all ranks are gathered to rank 0 over NCCL (point-to-point, exact byte ranges) and getSymmetries runs on rank 0's GPU.
    Smaller numMCTSSims than the headline so that complete games fit in seconds; everything is inside the timed region."""
    from azg_b200 import lib
    from azg_b200.mcts import Engine
    sims = args.iter_sims or 40
    n_games = args.iter_games or args.games
    a = mcts_args(sims, args.game)
    first_game, _ = shard_games(n_games * world, rank, world)
    eng = Engine(game, net, a, n_games=n_games, dirichlet_noise=True, seed=2000, node_cap=8 * sims + 256, first_game=first_game)
    eng.selfplay(max_moves=2)                                                     # warm-up plies
    warm = tuple(torch.zeros((4, 8), dtype=torch.uint8, device=dev) for _ in range(5))
    for _ in range(2):
        gather_examples(warm, dst=0)                                              # warm-up of the collective (NCCL opens its p2p channels lazily)
    s0 = eng.stats()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    stream = torch.cuda.current_stream()
    barrier()
    ev[0].record(stream)
    parts = []
    while True:
        left = n_games - (eng.stats()['episodes_finished'] - s0['episodes_finished'])
        if left <= 0:
            break
        eng.selfplay(min_episodes=left)
        parts.append(eng.examples_device(dev))                                    # ring -> torch CUDA tensors (device-to-device)
    local = tuple(torch.cat([p[i] for p in parts]) for i in range(5))
    ev[1].record(stream)
    (gb, gpi, gz, gva, gq), gst = gather_examples(local, dst=0, return_stats=True)
    ev[2].record(stream)
    n_aug = 0
    if rank == 0 and len(gb):                                                     # getSymmetries after the gather, on device pointers, in chunks
        L = lib.load(); K = game.info.max_symmetries; S = game.info.state_bytes; A = game.info.action_size; CH = 32768
        ob = torch.empty((CH, K, S), dtype=torch.int8, device=dev); opi = torch.empty((CH, K, A), dtype=torch.float32, device=dev)
        om = torch.empty((CH, K, A), dtype=torch.uint8, device=dev); ok = torch.empty((CH,), dtype=torch.int32, device=dev)
        tot = torch.zeros((), dtype=torch.int64, device=dev)
        gbf = gb.reshape(len(gb), S)
        for i in range(0, len(gb), CH):
            m = min(CH, len(gb) - i)
            lib.check(L.azg_game_symmetries(game.game_id, N_PL, m, lib.ptr(gbf[i:i + m]), lib.ptr(gpi[i:i + m]), lib.ptr(gva[i:i + m]), lib.ptr(ob), lib.ptr(opi),
                                            lib.ptr(om), lib.ptr(ok), None))
            tot += ok[:m].sum()
        n_aug = int(tot.item())
    ev[3].record(stream)
    barrier()
    s1 = eng.stats()
    ms_play = max_over_ranks(ev[0].elapsed_time(ev[1])); ms_gather = max_over_ranks(ev[1].elapsed_time(ev[2])); ms_sym = max_over_ranks(ev[2].elapsed_time(ev[3]))
    ms_total = max_over_ranks(ev[0].elapsed_time(ev[3]))
    d = {k: s1[k] - s0[k] for k in ('sims', 'moves_played', 'episodes_finished', 'examples_recorded', 'terminal_hits', 'gc_runs', 'gc_sweeps', 'arena_overflows', 'node_visits')}
    sims_total = sum_over_ranks(d['sims']); ex_total = sum_over_ranks(int(len(local[0]))); ep_total = sum_over_ranks(d['episodes_finished'])
    rec = game.info.state_bytes + 4 * game.info.action_size + game.info.action_size + 8 * N_PL
    out = {'numMCTSSims': sims, 'games_per_gpu': n_games, 'episodes_finished': ep_total, 'examples_unaugmented': ex_total, 'examples_after_symmetries_rank0': n_aug,
           'bytes_per_example': rec, 'sims': sims_total, 'sims_per_sec_incl_gather': sims_total / (ms_total * 1e-3), 'sims_per_sec_selfplay_only': sims_total / (ms_play * 1e-3),
           'selfplay_ms': ms_play, 'gather_ms': ms_gather, 'symmetries_ms': ms_sym, 'total_ms': ms_total, 'gather_share': ms_gather / ms_total,
           'collective': ('none (1 rank)' if world == 1 else f'NCCL point-to-point gather to rank 0 (batch_isend_irecv = one ncclGroup; {world - 1} senders, exact byte ranges, '
                          'no padding, device memory on both sides), after one all_gather of the counts'),
           'gather_bytes_received_rank0': gst['bytes_received'] if rank == 0 else None,
           'gather_GBps_into_rank0': (gst['bytes_received'] / (ms_gather * 1e-3) / 1e9) if (rank == 0 and world > 1 and ms_gather > 0) else None,
           'rank0_counters': dict(d, examples_dropped=s1['examples_dropped'], mean_depth=d['node_visits'] / max(d['sims'], 1))}
    eng.close()
    return out


# --------------------------------------------------------------------------- main (B200 arm) -----------
def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return 0
    import torch
    import torch.distributed as dist
    import azg_b200
    from azg_b200 import lib
    from azg_b200.mcts import Engine
    if not torch.cuda.is_available() or lib.device_count() <= 0:
        raise SystemExit('bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    lib.check(lib.load().azg_set_device(local))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    gm = GAMES[args.game]
    if args.game == 'splendor':
        game = azg_b200.SplendorGame(); net = azg_b200.NNetWrapper(game, {'nn_version': 80}, seed=0)      # identical weights on every rank
    elif args.game == 'santorini':
        game = azg_b200.SantoriniGame(); net = azg_b200.SantoriniNNetWrapper(game, {'nn_version': 89}, seed=0)
    elif args.game == 'azul':
        game = azg_b200.AzulGame(); net = azg_b200.AzulNNetWrapper(game, {'nn_version': 84}, seed=0)
    else:
        game = azg_b200.AbaloneGame(); net = azg_b200.AbaloneNNetWrapper(game, {'nn_version': 21}, seed=0)
    S_BYTES, N_ACT = gm['S'], gm['A']
    a = mcts_args(args.sims, args.game)
    node_cap = args.node_cap or (gm.get('node_cap_per_sim', 6) * args.sims + 320)      # sized so that the tier-2 GC (gc_sweeps) never runs in the timed region
    from azg_b200.dist import gather_examples, shard_games
    first_game, _ = shard_games(args.games * world, rank, world)               # global slot ids: the games do not depend on the world size
    eng = Engine(game, net, a, n_games=args.games, dirichlet_noise=True, seed=1000, node_cap=node_cap, first_game=first_game)
    K, W = args.steps, args.warmup
    stream = torch.cuda.current_stream()

    # ---- device-resident arm ------------------------------------------------------------------------
    eng.selfplay(max_moves=W)                                                  # warm-up plies (untimed)
    s0 = eng.stats()
    PROF_EVERY = 8                                                             # CUDA events around every 8th lock-step simulation of the timed region
    eng.profile(PROF_EVERY)                                                    # (around EVERY launch they cost ~4 % of the loop: 6 event records per simulation)
    clocks = ClockSampler(local)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.selfplay(max_moves=K)
    e1.record(stream)
    barrier()
    ms_local = e0.elapsed_time(e1)
    clk = clocks.stop()
    kt = eng.kernel_times(); eng.profile(False)
    s1 = eng.stats()
    d = {k: s1[k] - s0[k] for k in s1 if k not in ('max_nodes', 'node_cap', 'edge_cap')}
    ms = max_over_ranks(ms_local)
    sims_total = sum_over_ranks(d['sims'])
    value = sims_total / (ms * 1e-3)

    # ---- roofline of the dominant kernel (rank 0's launches; all ranks run the same kernels) --------------
    pk = peaks()
    visits, exps, evals = d['node_visits'], d['expansions'], d['nn_evals']
    samp = max(int(kt['select_launches']), 1) / max(d['sims'] / args.games, 1)   # fraction of the simulations whose launches were timed
    Lbar_vis = d['sum_legal_visited'] / max(visits, 1); Lbar_exp = d['sum_legal'] / max(exps, 1); Dbar = visits / max(d['sims'], 1)
    n_launch = max(int(kt['select_launches']), 1) / samp                       # launches in the timed region (bytes / flops below are totals of the region)
    # Algorithmic bytes PER KERNEL: what each kernel itself has to touch in this engine's data layout, counted from
    # the engine's own counters of the same run. SP = board slot padded to 16 B, MW = mask words, U = child links per edge.
    #   k_select : per visit the 32 B node header + 4 B cached link + 16 B path record; the root is scanned in full (16 B edge + 4 B link
    #              per legal action); the ONE new child of a simulation is materialised here: parent board read (SP), action id, hash-table
    #              probe (32 slots x 8 B) + key check (16 B), child link write, and the new leaf's hand-over to the net: board (SP),
    #              legal mask, key / round / link (32 B).
    #   k_backup : per expansion the net's outputs (4 A + 4 np), the mask, L new edges (16 B + action id + 4 U link bytes each), the
    #              board slot (read SP + write SP), key (16 B), header (32 B), hash-table insert (256 B probe + 8 B); per visited level
    #              the path record (16 B), edge read-modify-write (32 B), header read + partial write (48 B); the cached-choice
    #              refresh re-reads 16 B per edge of every non-root node on the path and rewrites its U links + best (2 B).
    # 'tree_path' below is the SURVEY's own figure (B_sel + B_bak + B_exp + B_nn of a dense re-scan design) over the time of both kernels.
    U = max(GAMES[args.game]['universes'], 1)
    SPAD = (S_BYTES + 15) // 16 * 16; MWB = 4 * ((N_ACT + 31) // 32); ACTB = 1 if N_ACT <= 256 else 2
    refreshed = max(visits - d['sims'], 0)                                       # non-root visits: one cached-choice refresh each
    new_leaves = exps + d['terminal_hits']
    sel_bytes = (52.0 * visits + 20.0 * d['sum_legal_root_scans'] + (SPAD + ACTB + 256 + 16 + 4 * U) * new_leaves + (SPAD + MWB + 32) * exps)
    bak_bytes = ((4.0 * N_ACT + 4 * N_PL + MWB + 2 * SPAD + 16 + 32 + 264) * exps + (16.0 + ACTB + 4 * U) * d['sum_legal']
                 + (16.0 + 32 + 48) * visits + 16.0 * d['sum_legal_refreshed'] + (2.0 + 4.0 * U) * refreshed)
    survey_bytes = (16.0 * visits + 14.0 * d['sum_legal_visited'] + 32.0 * visits + (2.0 * S_BYTES + 32.0) * exps + 14.0 * d['sum_legal']
                    + (S_BYTES + 11 + 4 * N_ACT + 4 * N_PL) * evals)
    survey_sel_bytes = 16.0 * visits + 14.0 * d['sum_legal_visited']             # SURVEY 8d B_sel = 16 + 14 L per select step (dense re-scan of every visited node)
    net_flops = float(gm['flops']) * evals
    tree_ms = kt['select_ms'] + kt['backup_ms']
    kern = {
        'select': {'ms': kt['select_ms'], 'bound': 'hbm', 'achieved': sel_bytes * samp / max(kt['select_ms'], 1e-9) / 1e6, 'peak': pk['hbm'], 'unit': 'GB/s',
                   'per_launch_bytes': sel_bytes / n_launch,
                   'survey_B_sel_frac': survey_sel_bytes * samp / max(kt['select_ms'], 1e-9) / 1e6 / pk['hbm'],
                   'survey_B_sel_note': 'SURVEY 8d books 16 + 14 L bytes per select step (a design that re-scans every visited node); this engine follows a cached choice (52 B per non-root visit), so it moves fewer bytes than that model'},
        'net': {'ms': kt['net_ms'], 'bound': 'tensor', 'achieved': net_flops * samp / max(kt['net_ms'], 1e-9) / 1e9, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
                    'per_launch_flops': net_flops / n_launch},
        'expand_backup': {'ms': kt['backup_ms'], 'bound': 'hbm', 'achieved': bak_bytes * samp / max(kt['backup_ms'], 1e-9) / 1e6, 'peak': pk['hbm'], 'unit': 'GB/s',
                          'per_launch_bytes': bak_bytes / n_launch},
    }
    tree_path = {'ms': tree_ms, 'bound': 'hbm', 'achieved': survey_bytes * samp / max(tree_ms, 1e-9) / 1e6, 'peak': pk['hbm'], 'unit': 'GB/s',
                 'per_step_bytes': survey_bytes / n_launch, 'frac': survey_bytes * samp / max(tree_ms, 1e-9) / 1e6 / pk['hbm'],
                 'note': 'SURVEY.md 8d bytes of select + expand + backup (B_sel + B_bak + B_exp + B_nn) over the time of k_select + k_backup'}
    tot_ms = kt['select_ms'] + kt['net_ms'] + kt['backup_ms'] + kt['other_ms'] / PROF_EVERY      # the per-move kernels are timed at every move, the simulations sampled
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')                          # dram bytes per launch from the committed ncu --set full captures
    tj = (json.load(open(tp)).get(args.game) or {}) if os.path.exists(tp) else {}  # per game: only captures of the same workload count
    for name, v in kern.items():
        v['frac'] = v['achieved'] / v['peak']; v['share'] = v['ms'] / max(tot_ms, 1e-9); v['avg_launch_us'] = 1e3 * v['ms'] / max(int(kt['select_launches']), 1)
        v['timed_launches'] = int(kt['select_launches'])
        v['dram_bytes_per_launch_ncu'] = tj.get(name)
        if tj.get(name) and 'per_launch_bytes' in v:
            v['dram_over_algorithmic'] = tj[name] / v['per_launch_bytes']
    dom = max(kern, key=lambda k: kern[k]['ms'])
    traffic = tj.get(dom)
    roofline = {'kernel': dom, 'bound': kern[dom]['bound'], 'achieved': kern[dom]['achieved'], 'peak': kern[dom]['peak'], 'unit': kern[dom]['unit'],
                'frac': kern[dom]['frac'], 'traffic': traffic, 'peak_source': pk['src'] + (' sustained bf16' if kern[dom]['bound'] == 'tensor' else ''),
                'avg_launch_us': kern[dom]['avg_launch_us'], 'share_of_step': kern[dom]['share']}

    # ---- e2e arm: same plies through the plugin calls with host buffers ---------------------------------------
    e2e = None
    if not args.no_e2e:
        eng.reset()
        hl = HostLoop(torch, game, eng, args.games, seed=7 + rank)
        for _ in range(W):
            hl.step()
        hl.h2d = hl.d2h = 0
        t0 = eng.stats()
        barrier()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(K):
            hl.step()
        f1.record(stream)
        barrier()
        ms2 = max_over_ranks(f0.elapsed_time(f1))
        t1 = eng.stats()
        sims2 = sum_over_ranks(t1['sims'] - t0['sims'])
        e2e = {'value': sims2 / (ms2 * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': hl.h2d // K, 'd2h_bytes_per_step': hl.d2h // K,
               'ms_per_step': ms2 / K, 'api': 'azg_engine_search + azg_game_next/ended/canonical with pinned host buffers'}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        r = cpu_sample(args.sims, args.cpu_plies, threads, game=args.game)
        cpu = {'value': r['sims'] / r['seconds'], 'unit': UNIT, 'cores': threads, 'kind': 'port', 'seconds': r['seconds'],
               'sample': f'{threads} host threads x 1 {args.game} self-play game x {args.cpu_plies} plies x {args.sims} sims (oracle/azg_oracle.c, same MCTS args and net weights)'}
        cpu.update(port_vs_reference(cpu['value']))

    eng.close()
    # ---- secondary line: the reference's default playout-cap randomisation (main.py:130-131: prob_fullMCTS 0.25, ratio_fullMCTS 5) ------
    pcr = None
    if not args.no_pcr:
        a2 = dict(a, prob_fullMCTS=0.25, ratio_fullMCTS=5)
        eng2 = Engine(game, net, a2, n_games=args.games, dirichlet_noise=True, seed=1000, node_cap=node_cap, first_game=first_game)
        eng2.selfplay(max_moves=2)
        t0 = eng2.stats(); barrier()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        g0.record(stream); eng2.selfplay(max_moves=2 * K); g1.record(stream); barrier()
        t1 = eng2.stats(); ms3 = max_over_ranks(g0.elapsed_time(g1)); sims3 = sum_over_ranks(t1['sims'] - t0['sims'])
        pcr = {'prob_fullMCTS': 0.25, 'ratio_fullMCTS': 5, 'sims_per_sec': sims3 / (ms3 * 1e-3), 'vs_full_search_value': sims3 / (ms3 * 1e-3) / value,
               'moves_played': t1['moves_played'] - t0['moves_played'], 'mean_sims_per_move': (t1['sims'] - t0['sims']) / max(t1['moves_played'] - t0['moves_played'], 1),
               'schedule': 'ragged: per-slot move boundaries (k_sp_turn), every launch works on all trees', 'ms': ms3}
        eng2.close()
    # ---- whole-iteration leg: complete games -> drain -> gather (NCCL) -> symmetries --------------------------------------
    iteration = None
    if not args.no_iteration:
        clocks2 = ClockSampler(local) if rank == 0 else None                     # this leg runs last, after ~1 minute of load: its own clocks line
        iteration = run_iteration(torch, dist, args, game, net, rank, world, dev, barrier, max_over_ranks, sum_over_ranks, gather_examples, shard_games)
        if clocks2 is not None:
            iteration['clocks'] = clocks2.stop()

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 net (token GEMMs 3xTF32 on tcgen05, fp32 accumulate) / f64 PUCT / i8 boards', 'data': 'synthetic',
                'config': workload_cfg(args), 'e2e': e2e, 'gpu_launches': int(d['kernels_launched']), 'roofline': roofline, 'cpu_baseline': cpu,
                'clocks': clk, 'kernels': kern, 'tree_path': tree_path, 'pcr_default': pcr, 'iteration': iteration,
                'counters': {'sims': d['sims'], 'node_visits': visits, 'expansions': exps, 'nn_evals': evals, 'terminal_hits': d['terminal_hits'],
                             'arena_overflows': d['arena_overflows'], 'gc_runs': d['gc_runs'], 'gc_sweeps': d['gc_sweeps'], 'examples_dropped': s1['examples_dropped'],
                             'moves_played': d['moves_played'],
                             'episodes_finished': d['episodes_finished'], 'mean_depth': Dbar, 'mean_legal_visited': Lbar_vis, 'mean_legal_expanded': Lbar_exp,
                             'expansions_per_sec': sum_over_ranks(exps) / (ms * 1e-3) if world == 1 else None, 'node_visits_per_sec': visits / (ms * 1e-3) if world == 1 else None,
                             'node_cap': s1['node_cap'], 'max_nodes': s1['max_nodes']}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

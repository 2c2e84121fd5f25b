#!/usr/bin/env python
"""Golden fixtures for Coach.executeEpisode (Coach.py:37-84) by RUNNING THE REFERENCE'S OWN executeEpisode (test infrastructure).

    python oracle/gen_golden_selfplay.py [--out tests/golden] [--only splendor,santorini,abalone,azul]

The reference's `Coach.executeEpisode(my_mcts, my_game)` runs UNMODIFIED. Every random input it consumes is recorded so that the
oracle and the CUDA engine can replay the same episode and must return the same training examples, example for example:
  u_full      the playout-cap coin  MCTS.rng.random()                     (MCTS.py:58)       -> recorded through MCTS.rng
  noise       the Dirichlet draw at the root of a full search             (MCTS.py:187-197)  -> recorded through MCTS.rng
  u_move      the uniform np.random.choice consumes in random_pick        (Coach.py:289-292) -> replayed from the global RandomState
  chance_seed the random_seed of the real move (Coach.py:71 passes 0 = true random; the harness' Game subclass forwards a recorded
              non-zero seed instead, so the draw is the deterministic `(4594591*(seed+...)) % n` one both sides reproduce bit for bit)
  init_board  what getInitBoard returned (numba RNG)
Outputs per game: the returned example list (board, pi, z, valids, q) AFTER getSymmetries, in order, plus the per-ply trace
(is_full, action, canonical root) for debugging. The net is the deterministic hash-net (oracle/hashnet.py).
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
os.environ.setdefault('OMP_NUM_THREADS', '1')

CONFIGS = {
    # prob_fullMCTS < 1 so that both budgets and un-recorded plies occur; noise on; forced playouts on for splendor (shipped args)
    'splendor': dict(numMCTSSims=60, cpuct=0.8, fpu=0.0593, universes=3, dirichletAlpha=0.3, temperature=[1.25, 0.8, 1.1], tempThreshold=10,
                     forced_playouts=True, prob_fullMCTS=0.6, ratio_fullMCTS=5, no_mem_optim=False, no_compression=True),
    'santorini': dict(numMCTSSims=50, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1], tempThreshold=10,
                      forced_playouts=False, prob_fullMCTS=0.5, ratio_fullMCTS=5, no_mem_optim=False, no_compression=True),
    'abalone': dict(numMCTSSims=40, cpuct=1.25, fpu=0.2, universes=1, dirichletAlpha=-1.0, temperature=[1.0, 0.3, 1.0], tempThreshold=-20,
                    forced_playouts=False, prob_fullMCTS=0.3, ratio_fullMCTS=4, no_mem_optim=False, no_compression=True),
}
CONFIGS['azul'] = dict(numMCTSSims=50, cpuct=1.0, fpu=0.1, universes=2, dirichletAlpha=0.5, temperature=[1.0, 0.4, 1.2], tempThreshold=6,
                       forced_playouts=True, prob_fullMCTS=0.5, ratio_fullMCTS=5, no_mem_optim=False, no_compression=True)
N_GAMES = {'splendor': 3, 'santorini': 4, 'abalone': 2, 'azul': 3}


def setup_paths(game):
    if game == 'santorini':
        import gen_golden_santorini  # noqa: F401  (copies santorini/ to a scratch dir with NB_GODS = 1 and puts it on sys.path)
    else:
        sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference', HERE]


def run(game, out):
    sys.path.insert(0, HERE)
    setup_paths(game)
    import numpy as np
    from numba import njit
    from hashnet import HashNet
    from gen_golden import RecordingRng, dotdict
    if game == 'santorini':
        sys.path.insert(0, '/tmp/azg_ref_santorini_nogods')
    from MCTS import MCTS
    import Coach as coach_mod
    if game == 'splendor':
        from splendor.SplendorGame import SplendorGame as Game
    elif game == 'santorini':
        from santorini.SantoriniGame import SantoriniGame as Game
    elif game == 'azul':
        from azul.AzulGame import AzulGame as Game
    else:
        from abalone.AbaloneGame import AbaloneGame as Game

    @njit
    def seed_numba(s):
        np.random.seed(s)

    cfg = CONFIGS[game]
    args = dotdict(cfg)
    save = {'n_games': np.array(N_GAMES[game])}
    for k, v in cfg.items():
        save['cfg_' + k] = np.array(v)

    for gi in range(N_GAMES[game]):
        trace = dict(u_full=[], noise=[], u_move=[], chance_seed=[], is_full=[], action=[], root=[])

        class HarnessGame(Game):
            """Game.py facade, unmodified behaviour except that the real move's random_seed=0 is replaced by a recorded seed."""
            def getInitBoard(self):
                b = Game.getInitBoard(self)
                trace['init'] = np.array(b, copy=True)
                return b

            def getNextState(self, board, player, action, random_seed=0):
                assert random_seed == 0                                   # Coach.py:71
                seed = int(crng.integers(1, 2 ** 31 - 1))
                trace['chance_seed'].append(seed); trace['action'].append(int(action))
                return Game.getNextState(self, board, player, action, random_seed=seed)

        class HarnessMCTS(MCTS):
            def getActionProb(self, cb, temp=1, force_full_search=False):
                nd, nr = len(self.rng.dirichlets), len(self.rng.randoms)
                trace['root'].append(np.array(cb, copy=True))
                res = MCTS.getActionProb(self, cb, temp=temp, force_full_search=force_full_search)
                assert len(self.rng.randoms) == nr + 1
                trace['u_full'].append(self.rng.randoms[nr])
                trace['noise'].append(self.rng.dirichlets[nd] if len(self.rng.dirichlets) > nd else np.zeros(0))
                trace['is_full'].append(bool(res[2]))
                rs = np.random.RandomState(); rs.set_state(np.random.get_state())
                trace['u_move'].append(float(rs.random_sample()))              # the uniform random_pick's np.random.choice draws next
                return res

        crng = np.random.default_rng(4242 + gi)
        seed_numba(500 + gi); np.random.seed(600 + gi)
        g = HarnessGame()
        g.getInitBoard()
        m = HarnessMCTS(g, HashNet(g), args, dirichlet_noise=(cfg['dirichletAlpha'] != 0))
        m.rng = RecordingRng(700 + gi)
        coach = coach_mod.Coach.__new__(coach_mod.Coach)                        # __init__ would build two torch nets; not needed here
        coach.game = g; coach.nnet = None; coach.args = args; coach.mcts = m; coach.nb_threads = 1
        seed_numba(500 + gi)
        examples = coach.executeEpisode(m, g)                                   # <- the reference's own code, unmodified
        P = len(trace['u_full'])
        A = g.getActionSize()
        L = max(max((len(x) for x in trace['noise']), default=0), 1)
        noise = np.zeros((P, L)); noise_len = np.zeros(P, np.int64)
        for i, x in enumerate(trace['noise']):
            noise[i, :len(x)] = x; noise_len[i] = len(x)
        p = f'g{gi}_'
        save[p + 'init'] = trace['init']; save[p + 'u_full'] = np.array(trace['u_full']); save[p + 'u_move'] = np.array(trace['u_move'])
        save[p + 'chance_seed'] = np.array(trace['chance_seed'], np.int64); save[p + 'noise'] = noise; save[p + 'noise_len'] = noise_len
        save[p + 'is_full'] = np.array(trace['is_full']); save[p + 'action'] = np.array(trace['action'], np.int32)
        save[p + 'root'] = np.array(trace['root'], np.int8)
        save[p + 'ex_board'] = np.array([e[0] for e in examples], np.int8)
        pis = np.array([e[1] for e in examples], np.float32)
        if A > 1000:                                                             # sparse policies: (example, action, value) triplets
            nz = np.nonzero(pis)
            save[p + 'ex_pi_idx'] = np.stack(nz).astype(np.int32); save[p + 'ex_pi_val'] = pis[nz]
            save[p + 'ex_valids_bits'] = np.packbits(np.array([e[3] for e in examples], np.bool_), axis=1)
        else:
            save[p + 'ex_pi'] = pis
            save[p + 'ex_valids'] = np.array([e[3] for e in examples], np.bool_)
        save[p + 'ex_z'] = np.array([e[2] for e in examples], np.float32)
        save[p + 'ex_q'] = np.array([e[4] for e in examples], np.float32)
        assert all(np.asarray(e[1]).dtype == np.float32 for e in examples[:3]) or True
        print(f'{game} game {gi}: {P} plies, {int(np.sum(trace["is_full"]))} full, {len(examples)} examples, z0={examples[0][2] if examples else None}')
    np.savez_compressed(os.path.join(out, f'{game}_selfplay.npz'), **save)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    ap.add_argument('--only', default='splendor,santorini,abalone,azul')
    a = ap.parse_args()
    games = a.only.split(',')
    if len(games) > 1:                                                           # one process per game: the patched santorini import must not leak
        import subprocess
        for g in games:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), '--out', a.out, '--only', g])
        return
    run(games[0], a.out)


if __name__ == '__main__':
    main()

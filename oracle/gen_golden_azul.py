#!/usr/bin/env python
"""Golden step vectors for Azul (2 players) by RUNNING THE UNMODIFIED REFERENCE (test infrastructure; round-2 groundwork).

    python oracle/gen_golden_azul.py [--out tests/golden]

Imports azul/AzulGame.py (-> AzulLogicNumba.Board jitclass) from /root/reference. Random games are played with
`random_seed != 0`, so every tile draw is the reference's deterministic one and the recorded next states are exact.
Only the vectors are committed.
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference', HERE]

import numpy as np  # noqa: E402

A = 180


def gen_kat(out, n_games=12):
    from azul.AzulGame import AzulGame
    g = AzulGame()
    assert g.getActionSize() == A and tuple(g.getBoardSize()) == (23, 6) and g.getNumberOfPlayers() == 2
    rng = np.random.default_rng(2029)
    keys = ('board', 'player', 'canonical', 'valids', 'action', 'seed', 'next_board', 'next_player', 'ended', 'round', 'score', 'game', 'next_canonical')
    rec = {k: [] for k in keys}
    sym = {k: [] for k in ('board', 'pi', 'valids', 'out_boards', 'out_pi', 'out_valids')}
    for ep in range(n_games):
        board = np.array(g.getInitBoard(), copy=True)
        player, ply = 0, 0
        while True:
            cb = np.array(g.getCanonicalForm(board, player), copy=True)
            valids = np.array(g.getValidMoves(cb, 0), copy=True)
            assert (valids == np.array(g.getValidMoves(board, player))).all()
            legal = np.flatnonzero(valids)
            w = np.ones(len(legal))
            w[legal % 6 == 5] = 0.15 if ep % 3 else 1.0          # mostly avoid the floor so that walls fill up and games end
            action = int(rng.choice(legal, p=w / w.sum()))
            seed = int(rng.integers(1, 2 ** 31))
            nb, nplayer = g.getNextState(board, player, action, random_seed=seed)
            nb = np.array(nb, copy=True)
            ended = np.array(g.getGameEnded(nb, nplayer), copy=True)
            rec['board'].append(board.copy()); rec['player'].append(player); rec['canonical'].append(cb); rec['valids'].append(valids.copy())
            rec['action'].append(action); rec['seed'].append(seed); rec['next_board'].append(nb); rec['next_player'].append(nplayer)
            rec['ended'].append(ended); rec['round'].append(int(g.getRound(nb))); rec['score'].append([int(g.getScore(nb, 0)), int(g.getScore(nb, 1))])
            rec['game'].append(ep); rec['next_canonical'].append(np.array(g.getCanonicalForm(nb, nplayer), copy=True))
            if ply % 23 == 5:
                pi = rng.random(A).astype(np.float32)
                s = g.getSymmetries(cb, pi, valids)
                assert len(s) == 120
                sym['board'].append(cb); sym['pi'].append(pi); sym['valids'].append(valids.copy())
                sym['out_boards'].append(np.array([x[0] for x in s], dtype=np.int8))
                sym['out_pi'].append(np.array([x[1] for x in s], dtype=np.float32))
                sym['out_valids'].append(np.array([x[2] for x in s], dtype=np.bool_))
            board, player, ply = nb, nplayer, ply + 1
            if ended.any() or ply > 400:
                break
    arrs = {k: np.array(v) for k, v in rec.items()}
    arrs['board'] = arrs['board'].astype(np.int8); arrs['seed'] = arrs['seed'].astype(np.int64)
    for k, v in sym.items():
        arrs['sym_' + k] = np.array(v)
    np.savez_compressed(os.path.join(out, 'azul_kat.npz'), **arrs)
    e = arrs['ended']
    print(f'azul kat: {len(arrs["action"])} plies, {n_games} games, legal mean {arrs["valids"].sum(1).mean():.1f} max {arrs["valids"].sum(1).max()}, '
          f'rounds {[int(arrs["round"][np.flatnonzero(arrs["game"] == ep)[-1]]) for ep in range(n_games)]}, '
          f'final scores {[arrs["score"][np.flatnonzero(arrs["game"] == ep)[-1]].tolist() for ep in range(n_games)]}, '
          f'results {e[np.abs(e).sum(1) > 0].tolist()}, sym={len(sym["pi"])}')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    a = ap.parse_args()
    gen_kat(a.out)

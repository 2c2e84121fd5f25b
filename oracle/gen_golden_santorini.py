#!/usr/bin/env python
"""Golden fixtures for Santorini WITHOUT gods (BASELINE.json configs[1]) by RUNNING THE REFERENCE (test infrastructure).

The no-gods game is a source-level constant in the reference (`NB_GODS = 1`, santorini/SantoriniConstants.py:19,
baked at import into action_size and the permutation tables), so this script copies the reference's `santorini/`
package to a scratch directory OUTSIDE the repo, edits that one constant there, and imports the patched copy next
to the unmodified top-level modules (Game.py, MCTS.py). Nothing of the reference is written into the repo; only the
vectors it produces are committed (tests/golden/santorini_*.npz) together with this script.

    python oracle/gen_golden_santorini.py [--out tests/golden]
"""
import argparse
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache_santorini')
os.environ.setdefault('OMP_NUM_THREADS', '1')
SCRATCH = '/tmp/azg_ref_santorini_nogods'


def patched_reference():
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    os.makedirs(SCRATCH)
    shutil.copytree('/root/reference/santorini', os.path.join(SCRATCH, 'santorini'),
                    ignore=shutil.ignore_patterns('*.pt', '*.gif', '*.mp4', '*.png', '__pycache__'))
    p = os.path.join(SCRATCH, 'santorini', 'SantoriniConstants.py')
    src = open(p).read()
    new, n = re.subn(r'^NB_GODS = 11\b', 'NB_GODS = 1', src, flags=re.M)
    assert n == 1, 'NB_GODS constant not found'
    open(p, 'w').write(new)
    sys.path[:0] = [os.path.join(HERE, 'ref_shim'), SCRATCH, '/root/reference', HERE]


patched_reference()
import numpy as np  # noqa: E402
from numba import njit  # noqa: E402

from hashnet import HashNet  # noqa: E402
from gen_golden import MCTS_CONFIGS, RecordingRng, dotdict, tree_summary  # noqa: E402

sys.path.insert(0, SCRATCH)          # gen_golden put /root/reference first: the patched `santorini` package must win


@njit
def _seed_numba(s):
    np.random.seed(s)


def gen_kat(out, n_games=12):
    from santorini.SantoriniGame import SantoriniGame
    g = SantoriniGame()
    assert g.getActionSize() == 162 and tuple(g.getBoardSize()) == (5, 5, 3)
    rng = np.random.default_rng(2027)
    rec = {k: [] for k in ('board', 'player', 'canonical', 'valids', 'action', 'next_board', 'next_player', 'ended',
                           'ended_canonical0', 'round', 'score', 'game', 'next_canonical')}
    sym = {k: [] for k in ('board', 'pi', 'valids', 'out_boards', 'out_pi', 'out_valids')}
    inits = []
    for ep in range(n_games):
        _seed_numba(3000 + ep)
        board = g.getInitBoard().copy()
        inits.append(board.copy())
        player, ply = 0, 0
        while True:
            cb = np.array(g.getCanonicalForm(board, player), copy=True)
            valids = np.array(g.getValidMoves(cb, 0), copy=True)
            assert (valids == np.array(g.getValidMoves(board, player))).all()
            legal = np.flatnonzero(valids)
            # mostly random, sometimes climb-greedy so that games also end by reaching level 3
            action = int(rng.choice(legal))
            nb, nplayer = g.getNextState(board, player, action, random_seed=31416)
            nb = np.array(nb, copy=True)
            ended = np.array(g.getGameEnded(nb, nplayer), copy=True)
            ncb = np.array(g.getCanonicalForm(nb, nplayer), copy=True)
            rec['board'].append(board.copy()); rec['player'].append(player); rec['canonical'].append(cb); rec['valids'].append(valids)
            rec['action'].append(action); rec['next_board'].append(nb); rec['next_player'].append(nplayer); rec['ended'].append(ended)
            rec['ended_canonical0'].append(np.array(g.getGameEnded(ncb, 0), copy=True))      # the call MCTS.search makes (MCTS.py:131)
            rec['round'].append(int(g.getRound(nb))); rec['score'].append([int(g.getScore(nb, 0)), int(g.getScore(nb, 1))])
            rec['game'].append(ep); rec['next_canonical'].append(ncb)
            if ply % 2 == 0:
                pi = rng.random(162).astype(np.float32)
                s = g.getSymmetries(cb, pi, valids)
                assert len(s) == 8
                sym['board'].append(cb); sym['pi'].append(pi); sym['valids'].append(valids)
                sym['out_boards'].append(np.array([x[0] for x in s], dtype=np.int8))
                sym['out_pi'].append(np.array([x[1] for x in s], dtype=np.float32))
                sym['out_valids'].append(np.array([x[2] for x in s], dtype=np.bool_))
            board, player, ply = nb, nplayer, ply + 1
            if ended.any():
                break
    arrs = {k: np.array(v) for k, v in rec.items()}
    arrs['board'] = arrs['board'].astype(np.int8)
    arrs['init_boards'] = np.array(inits, dtype=np.int8)
    for k, v in sym.items():
        arrs['sym_' + k] = np.array(v)
    np.savez_compressed(os.path.join(out, 'santorini_kat.npz'), **arrs)
    e = arrs['ended']
    print(f'santorini kat: {len(arrs["action"])} plies, {n_games} games, mean legal {arrs["valids"].sum(1).mean():.1f} '
          f'max {arrs["valids"].sum(1).max()}, results {e[np.abs(e).sum(1) > 0].tolist()[:4]}..., sym={len(sym["pi"])}')


def gen_mcts(out):
    from santorini.SantoriniGame import SantoriniGame
    from MCTS import MCTS
    kat = np.load(os.path.join(out, 'santorini_kat.npz'))
    g = SantoriniGame()
    net = HashNet(g)
    game_ids = kat['game']
    idx_by_game = {ep: np.flatnonzero(game_ids == ep) for ep in np.unique(game_ids)}
    picks = []
    for ep in (0, 1, 2):
        idx = idx_by_game[ep]
        for frac in (0.0, 0.4, 0.8, 0.97):
            picks.append(int(idx[min(len(idx) - 1, int(frac * len(idx)))]))
    cases = []
    for ci, (name, cfg) in enumerate(MCTS_CONFIGS.items()):
        for pi_, p in enumerate(picks):
            if name != 'default' and pi_ % 3 != 0:
                continue
            n_sims = 800 if (name == 'default' and pi_ in (0, 5)) else 200
            args = dotdict(cfg, numMCTSSims=n_sims)
            m = MCTS(g, net, args, dirichlet_noise=cfg['noise'])
            rr = RecordingRng(500 + 100 * ci + pi_)
            m.rng = rr
            root = np.array(kat['canonical'][p], copy=True)
            probs, q, full = m.getActionProb(root, temp=1, force_full_search=True)
            s = g.stringRepresentation(root)
            A = g.getActionSize()
            raw = np.array([m.nodes_data[s][5][a] for a in range(A)], dtype=np.int64)
            cases.append(dict(cfg=name, root=root, n_sims=n_sims, probs=np.array(probs, dtype=np.float64), q=np.array(q, dtype=np.float32),
                              raw_counts=raw, root_P=np.array(m.nodes_data[s][2], dtype=np.float32),
                              root_Qsa=np.array(m.nodes_data[s][4], dtype=np.float64),
                              noise=(rr.dirichlets[0] if rr.dirichlets else np.zeros(0)), summary=tree_summary(m)))
            print(f'santorini mcts {name} root#{p} n={n_sims} nodes={cases[-1]["summary"]} top={int(np.argmax(raw))}:{int(raw.max())}')
    save = {'n_cases': np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            save[f'c{i}_{k}'] = np.array(v)
    np.savez_compressed(os.path.join(out, 'santorini_mcts.npz'), **save)


def gen_episode(out):
    """Tree reuse across moves: one self-play game with the hash-net (Coach.py:55-84 loop)."""
    from santorini.SantoriniGame import SantoriniGame
    from MCTS import MCTS
    g = SantoriniGame()
    net = HashNet(g)
    cfg = MCTS_CONFIGS['default']
    nsims = 150
    args = dotdict(cfg, numMCTSSims=nsims)
    _seed_numba(91); np.random.seed(6)
    m = MCTS(g, net, args, dirichlet_noise=False)
    m.rng = RecordingRng(10)
    board = g.getInitBoard().copy(); player = 0
    roots, counts, qs, actions, players, summaries, next_boards, enders = [], [], [], [], [], [], [], []
    while True:
        cb = np.array(g.getCanonicalForm(board, player), copy=True)
        probs, q, full = m.getActionProb(cb, temp=1, force_full_search=True)
        s = g.stringRepresentation(cb)
        raw = np.array([m.nodes_data[s][5][a] for a in range(162)], dtype=np.int64)
        action = int(np.random.choice(162, p=np.array(probs) / np.sum(probs)))
        nb, nplayer = g.getNextState(board, player, action)
        nb = np.array(nb, copy=True)
        r = np.array(g.getGameEnded(nb, nplayer), copy=True)
        roots.append(cb); counts.append(raw); qs.append(np.array(q, dtype=np.float32)); actions.append(action); players.append(player)
        summaries.append(tree_summary(m)); next_boards.append(nb); enders.append(r)
        board, player = nb, nplayer
        if r.any():
            break
    np.savez_compressed(os.path.join(out, 'santorini_episode.npz'), n_sims=np.array(nsims), roots=np.array(roots), raw_counts=np.array(counts),
                        q=np.array(qs), actions=np.array(actions), players=np.array(players), summaries=np.array(summaries),
                        next_boards=np.array(next_boards), ended=np.array(enders))
    print(f'santorini episode: {len(roots)} plies, final r={enders[-1]}, last summary={summaries[-1]}')


def gen_net(out):
    """SantoriniNNet V89 forward (santorini/SantoriniNNet.py:194-217,273-279) through the reference's torch branch of predict."""
    import torch
    torch.set_num_threads(1)
    from santorini.SantoriniGame import SantoriniGame
    from santorini.NNet import NNetWrapper
    kat = np.load(os.path.join(out, 'santorini_kat.npz'))
    g = SantoriniGame()
    nn_args = dict(nn_version=89, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32, no_compression=True, q_weight=0.5)
    sel = np.linspace(0, len(kat['canonical']) - 1, 64).astype(int)
    boards, valids = kat['canonical'][sel], kat['valids'][sel]
    for tag in ('rand', 'shipped'):
        torch.manual_seed(0)
        w = NNetWrapper(g, nn_args)
        w.device['inference'] = 'cpu'
        if tag == 'rand':
            gen = torch.Generator().manual_seed(2)
            with torch.no_grad():
                for mod in w.nnet.modules():
                    if isinstance(mod, torch.nn.BatchNorm2d):
                        mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.3)
                        mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) * 1.5 + 0.25)
                        mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=gen) + 0.5)
                        mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=gen) * 0.2)
                for name, p in w.nnet.named_parameters():
                    if name.endswith('.bias') and 'bn' not in name and 'first_layer.1' not in name:
                        p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        else:
            ck = torch.load('/root/reference/santorini/pretrained.pt', map_location='cpu', weights_only=False)
            w.nnet.load_state_dict(ck['state_dict'])
        w.nnet.eval()
        pis, vs = [], []
        for b, v in zip(boards, valids):
            pi, val = w.predict(b, v)
            pis.append(pi); vs.append(val)
        sd = {k: t.detach().cpu().numpy() for k, t in w.nnet.state_dict().items()}
        save = {'sd__' + k: v for k, v in sd.items()}
        save.update(boards=boards, valids=valids, pi=np.array(pis, dtype=np.float32), v=np.array(vs, dtype=np.float32))
        np.savez_compressed(os.path.join(out, f'santorini_v89_{tag}.npz'), **save)
        print(f'santorini net {tag}: {len(sd)} tensors, {sum(v.size for v in sd.values())} values, pi[0] max={pis[0].max():.4f} v[0]={vs[0]}')
        if tag == 'rand':
            for k, v in sd.items():
                print('   ', k, v.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    ap.add_argument('--only', default='kat,mcts,episode,net')
    a = ap.parse_args()
    only = a.only.split(',')
    if 'kat' in only:
        gen_kat(a.out)
    if 'mcts' in only:
        gen_mcts(a.out)
    if 'episode' in only:
        gen_episode(a.out)
    if 'net' in only:
        gen_net(a.out)


if __name__ == '__main__':
    main()

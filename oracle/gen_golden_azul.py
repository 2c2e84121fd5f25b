#!/usr/bin/env python
"""Golden step vectors for Azul (2 players) by RUNNING THE UNMODIFIED REFERENCE (test infrastructure; round-2 groundwork).

    python oracle/gen_golden_azul.py [--out tests/golden] [--only kat,mcts,episode,net]

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

from hashnet import HashNet  # noqa: E402
from gen_golden import MCTS_CONFIGS, RecordingRng, dotdict, tree_summary  # noqa: E402

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


def gen_mcts(out):
    """Root visit counts of the reference MCTS (MCTS.py:49-184) on Azul positions, driven by the hash-net (chance draws at round ends
    follow the universes' seeds, MCTS.py:63)."""
    from azul.AzulGame import AzulGame
    from MCTS import MCTS
    kat = np.load(os.path.join(out, 'azul_kat.npz'))
    g = AzulGame()
    net = HashNet(g)
    idx0 = np.flatnonzero(kat['game'] == 0); idx1 = np.flatnonzero(kat['game'] == 6)
    picks = [int(idx0[0]), int(idx0[len(idx0) // 3]), int(idx0[2 * len(idx0) // 3]), int(idx0[-4]), int(idx1[len(idx1) // 2]), int(idx1[-6])]
    cases = []
    for ci, (name, cfg) in enumerate(MCTS_CONFIGS.items()):
        for pi_, p in enumerate(picks):
            if name != 'default' and pi_ % 2 == 1:
                continue
            n_sims = 800 if (name == 'default' and pi_ in (0, 3)) else 200
            args = dotdict(cfg, numMCTSSims=n_sims)
            m = MCTS(g, net, args, dirichlet_noise=cfg['noise'])
            rr = RecordingRng(1300 + 100 * ci + pi_)
            m.rng = rr
            root = np.array(kat['canonical'][p], copy=True)
            probs, q, full = m.getActionProb(root, temp=1, force_full_search=True)
            s = g.stringRepresentation(root)
            raw = np.array([m.nodes_data[s][5][a] for a in range(A)], dtype=np.int64)
            cases.append(dict(cfg=name, root=root, n_sims=n_sims, probs=np.array(probs, dtype=np.float64), q=np.array(q, dtype=np.float32), raw_counts=raw,
                              noise=(rr.dirichlets[0] if rr.dirichlets else np.zeros(0)), summary=tree_summary(m)))
            print(f'azul mcts {name} root#{p} n={n_sims} nodes={cases[-1]["summary"]} top={int(np.argmax(raw))}:{int(raw.max())}')
    save = {'n_cases': np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            save[f'c{i}_{k}'] = np.array(v)
    np.savez_compressed(os.path.join(out, 'azul_mcts.npz'), **save)


def gen_episode(out):
    """Tree reuse across the plies of one game (MCTS.py:86-91 cleaning included): same MCTS object for 40 plies."""
    from azul.AzulGame import AzulGame
    from MCTS import MCTS
    kat = np.load(os.path.join(out, 'azul_kat.npz'))
    g = AzulGame()
    net = HashNet(g)
    cfg = MCTS_CONFIGS['default']; nsims = 120
    args = dotdict(cfg, numMCTSSims=nsims)
    np.random.seed(9)
    m = MCTS(g, net, args, dirichlet_noise=False)
    m.rng = RecordingRng(12)
    board = np.array(kat['board'][0], copy=True); player = int(kat['player'][0])
    roots, cnts, qs, actions, seeds, summaries = [], [], [], [], [], []
    for ply in range(40):
        cb = np.array(g.getCanonicalForm(board, player), copy=True)
        probs, q, full = m.getActionProb(cb, temp=1, force_full_search=True)
        raw = np.array([m.nodes_data[g.stringRepresentation(cb)][5][a] for a in range(A)], dtype=np.int64)
        action = int(np.random.choice(A, p=np.array(probs) / np.sum(probs)))
        seed = int(np.random.randint(1, 2 ** 31 - 1))
        nb, nplayer = g.getNextState(board, player, action, random_seed=seed)
        roots.append(cb); cnts.append(raw); qs.append(np.array(q, dtype=np.float32)); actions.append(action); seeds.append(seed); summaries.append(tree_summary(m))
        board, player = np.array(nb, copy=True), nplayer
        if np.array(g.getGameEnded(board, player)).any():
            break
    np.savez_compressed(os.path.join(out, 'azul_episode.npz'), n_sims=np.array(nsims), roots=np.array(roots), raw_counts=np.array(cnts), q=np.array(qs),
                        actions=np.array(actions), seeds=np.array(seeds, dtype=np.int64), summaries=np.array(summaries))
    print(f'azul episode: {len(roots)} plies, rounds seen {sorted(set(int(r[0, 2]) for r in roots))}, last summary={summaries[-1]}')


def gen_net(out):
    """AzulNNet V84 forward (azul/AzulNNet.py:84-137) through the reference's torch branch of GenericNNetWrapper.predict
    (GenericNNetWrapper.py:111-120): random-init with perturbed BatchNorm statistics / biases, and the shipped azul/pretrained.pt."""
    import torch
    torch.set_num_threads(1)
    from azul.AzulGame import AzulGame
    from azul.NNet import NNetWrapper
    kat = np.load(os.path.join(out, 'azul_kat.npz'))
    g = AzulGame()
    nn_args = dict(nn_version=84, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32, no_compression=True, q_weight=0.5)
    sel = np.linspace(0, len(kat['canonical']) - 1, 64).astype(int)
    boards = kat['canonical'][sel]; valids = kat['valids'][sel]
    for tag in ('rand', 'shipped'):
        torch.manual_seed(0)
        w = NNetWrapper(g, nn_args)
        w.device['inference'] = 'cpu'
        if tag == 'rand':
            gen = torch.Generator().manual_seed(4)
            with torch.no_grad():
                for mod in w.nnet.modules():
                    if isinstance(mod, torch.nn.BatchNorm1d):
                        mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.3)
                        mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) * 1.5 + 0.25)
                        mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=gen) + 0.5)
                        mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=gen) * 0.2)
        else:
            ck = torch.load('/root/reference/azul/pretrained.pt', map_location='cpu', weights_only=False)
            print('azul pretrained args:', {k: ck['args'][k] for k in ck.get('args', {}) if k in ('nn_version', 'cpuct', 'fpu', 'universes', 'numMCTSSims')} if isinstance(ck.get('args'), dict) else type(ck.get('args')))
            w.nnet.load_state_dict(ck['state_dict'])
        w.nnet.eval()
        pis, vs = [], []
        for b, v in zip(boards, valids):
            pi, val = w.predict(b, v)
            pis.append(pi); vs.append(val)
        sd = {k: t.detach().cpu().numpy() for k, t in w.nnet.state_dict().items()}
        save = {'sd__' + k: v for k, v in sd.items()}
        save.update(boards=boards, valids=valids, pi=np.array(pis, dtype=np.float32), v=np.array(vs, dtype=np.float32))
        np.savez_compressed(os.path.join(out, f'azul_v84_{tag}.npz'), **save)
        print(f'azul net {tag}: {len(sd)} tensors, {sum(v.size for v in sd.values())} values, pi[0] max={pis[0].max():.4f} v[0]={vs[0]}')


if __name__ == '__main__':
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

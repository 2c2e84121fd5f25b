#!/usr/bin/env python
"""Golden fixtures for Abalone (Belgian daisy, BASELINE.json configs[4]) by RUNNING THE UNMODIFIED REFERENCE (test infrastructure).

    python oracle/gen_golden_abalone.py [--out tests/golden] [--only kat,mcts,episode,net]

Imports abalone/AbaloneGame.py (-> AbaloneLogicNumba.Board jitclass, shipped constants INITIAL_LAYOUT = 1,
ENABLE_DYNAMIC_KOMI = False), MCTS.py and abalone/NNet.py (torch CPU branch) from /root/reference. Only the vectors are committed.
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

A = 3402


def gen_kat(out, n_games=6):
    from abalone.AbaloneGame import AbaloneGame
    g = AbaloneGame()
    assert g.getActionSize() == A and tuple(g.getBoardSize()) == (9, 9, 4)
    rng = np.random.default_rng(2028)
    rec = {k: [] for k in ('board', 'player', 'canonical', 'valids', 'action', 'next_board', 'next_player', 'ended', 'round', 'score',
                           'game', 'next_canonical')}
    sym = {k: [] for k in ('board', 'pi', 'valids', 'out_boards', 'out_pi', 'out_valids')}
    init = g.getInitBoard().copy()
    for ep in range(n_games):
        board = g.getInitBoard().copy()
        player, ply = 0, 0
        while True:
            cb = np.array(g.getCanonicalForm(board, player), copy=True)
            valids = np.array(g.getValidMoves(cb, 0), copy=True)
            assert (valids == np.array(g.getValidMoves(board, player))).all()
            legal = np.flatnonzero(valids)
            # prefer pushes (in-line moves of 2-3 marbles against an opponent marble) so that marbles get ejected and games can end by score
            w = np.ones(len(legal))
            planes = legal % 42
            w[planes >= 6] = 3.0
            if ep % 2 == 0:
                scores_before = int(g.getScore(board, player))
                for j, a in enumerate(legal[:0]):
                    pass
            action = int(rng.choice(legal, p=w / w.sum()))
            # greedy ejection every few plies: pick a move that increases the mover's score if one exists
            if ply % 3 == 0:
                for a in legal[rng.permutation(len(legal))][:40]:
                    nb_, _ = g.getNextState(board, player, int(a))
                    if int(g.getScore(nb_, player)) > int(g.getScore(board, player)):
                        action = int(a); break
            nb, nplayer = g.getNextState(board, player, action, random_seed=31416)
            nb = np.array(nb, copy=True)
            ended = np.array(g.getGameEnded(nb, nplayer), copy=True)
            rec['board'].append(board.copy()); rec['player'].append(player); rec['canonical'].append(cb); rec['valids'].append(np.packbits(valids))
            rec['action'].append(action); rec['next_board'].append(nb); rec['next_player'].append(nplayer); rec['ended'].append(ended)
            rec['round'].append(int(g.getRound(nb))); rec['score'].append([int(g.getScore(nb, 0)), int(g.getScore(nb, 1))]); rec['game'].append(ep)
            rec['next_canonical'].append(np.array(g.getCanonicalForm(nb, nplayer), copy=True))
            if ply % 16 == 3:
                pi = rng.random(A).astype(np.float32)
                s = g.getSymmetries(cb, pi, valids)
                assert len(s) == 12
                sym['board'].append(cb); sym['pi'].append(pi); sym['valids'].append(np.packbits(valids))
                sym['out_boards'].append(np.array([x[0] for x in s], dtype=np.int8))
                sym['out_pi'].append(np.array([x[1] for x in s], dtype=np.float32))
                sym['out_valids'].append(np.packbits(np.array([x[2] for x in s], dtype=np.bool_), axis=1))
            board, player, ply = nb, nplayer, ply + 1
            if ended.any():
                break
    arrs = {k: np.array(v) for k, v in rec.items()}
    arrs['board'] = arrs['board'].astype(np.int8)
    arrs['init_board'] = init.astype(np.int8)
    for k, v in sym.items():
        arrs['sym_' + k] = np.array(v)
    np.savez_compressed(os.path.join(out, 'abalone_kat.npz'), **arrs)
    nv = np.unpackbits(arrs['valids'], axis=1)[:, :A].sum(1)
    e = arrs['ended']
    print(f'abalone kat: {len(arrs["action"])} plies, {n_games} games, legal mean {nv.mean():.1f} max {nv.max()}, final scores '
          f'{[arrs["score"][np.flatnonzero(arrs["game"] == ep)[-1]].tolist() for ep in range(n_games)]}, results {e[np.abs(e).sum(1) > 0].tolist()}, sym={len(sym["pi"])}')


def gen_mcts(out):
    from abalone.AbaloneGame import AbaloneGame
    from MCTS import MCTS
    kat = np.load(os.path.join(out, 'abalone_kat.npz'))
    g = AbaloneGame()
    net = HashNet(g)
    idx0 = np.flatnonzero(kat['game'] == 0); idx1 = np.flatnonzero(kat['game'] == 1)
    picks = [int(idx0[0]), int(idx0[len(idx0) // 3]), int(idx0[2 * len(idx0) // 3]), int(idx0[-2]), int(idx1[len(idx1) // 2]), int(idx1[-3])]
    cases = []
    for ci, (name, cfg) in enumerate(MCTS_CONFIGS.items()):
        for pi_, p in enumerate(picks):
            if name != 'default' and pi_ % 3 != 1:
                continue
            n_sims = 400 if (name == 'default' and pi_ == 0) else 160
            args = dotdict(cfg, numMCTSSims=n_sims)
            m = MCTS(g, net, args, dirichlet_noise=cfg['noise'])
            rr = RecordingRng(900 + 100 * ci + pi_)
            m.rng = rr
            root = np.array(kat['canonical'][p], copy=True)
            probs, q, full = m.getActionProb(root, temp=1, force_full_search=True)
            s = g.stringRepresentation(root)
            raw = np.array(m.nodes_data[s][5], dtype=np.int64)
            nz = np.flatnonzero(raw)
            cases.append(dict(cfg=name, root=root, n_sims=n_sims, q=np.array(q, dtype=np.float32), raw_idx=nz.astype(np.int32), raw_cnt=raw[nz],
                              probs_nz=np.array(probs, dtype=np.float64)[np.flatnonzero(np.array(probs))], probs_idx=np.flatnonzero(np.array(probs)).astype(np.int32),
                              noise=(rr.dirichlets[0] if rr.dirichlets else np.zeros(0)), summary=tree_summary(m)))
            print(f'abalone mcts {name} root#{p} n={n_sims} nodes={cases[-1]["summary"]} top={int(np.argmax(raw))}:{int(raw.max())}')
    save = {'n_cases': np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            save[f'c{i}_{k}'] = np.array(v)
    np.savez_compressed(os.path.join(out, 'abalone_mcts.npz'), **save)


def gen_mcts1600(out):
    """BASELINE.json configs[4] runs numMCTSSims = 1600: two reference searches of that length (opening position with the main.py
    defaults; a mid-game position with the shipped-style arguments incl. injected root noise), same format as abalone_mcts.npz."""
    from abalone.AbaloneGame import AbaloneGame
    from MCTS import MCTS
    kat = np.load(os.path.join(out, 'abalone_kat.npz'))
    g = AbaloneGame(); net = HashNet(g)
    idx0 = np.flatnonzero(kat['game'] == 0)
    cases = []
    for ci, (name, p) in enumerate((('default', int(idx0[0])), ('shipped', int(idx0[len(idx0) // 2])))):
        cfg = MCTS_CONFIGS[name]; args = dotdict(cfg, numMCTSSims=1600)
        m = MCTS(g, net, args, dirichlet_noise=cfg['noise']); rr = RecordingRng(1600 + ci); m.rng = rr
        root = np.array(kat['canonical'][p], copy=True)
        probs, q, full = m.getActionProb(root, temp=1, force_full_search=True)
        raw = np.array(m.nodes_data[g.stringRepresentation(root)][5], dtype=np.int64); nz = np.flatnonzero(raw)
        cases.append(dict(cfg=name, root=root, n_sims=1600, q=np.array(q, dtype=np.float32), raw_idx=nz.astype(np.int32), raw_cnt=raw[nz],
                          probs_nz=np.array(probs, dtype=np.float64)[np.flatnonzero(np.array(probs))], probs_idx=np.flatnonzero(np.array(probs)).astype(np.int32),
                          noise=(rr.dirichlets[0] if rr.dirichlets else np.zeros(0)), summary=tree_summary(m)))
        print(f'abalone mcts1600 {name} root#{p} nodes={cases[-1]["summary"]} top={int(np.argmax(raw))}:{int(raw.max())}')
    save = {'n_cases': np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            save[f'c{i}_{k}'] = np.array(v)
    np.savez_compressed(os.path.join(out, 'abalone_mcts1600.npz'), **save)


def gen_episode(out):
    from abalone.AbaloneGame import AbaloneGame
    from MCTS import MCTS
    g = AbaloneGame()
    net = HashNet(g)
    cfg = MCTS_CONFIGS['default']; nsims = 100
    args = dotdict(cfg, numMCTSSims=nsims)
    np.random.seed(8)
    m = MCTS(g, net, args, dirichlet_noise=False)
    m.rng = RecordingRng(11)
    board = g.getInitBoard().copy(); player = 0
    roots, idxs, cnts, qs, actions, summaries = [], [], [], [], [], []
    for ply in range(30):                                     # 30 plies of tree reuse (a full game is 127 plies)
        cb = np.array(g.getCanonicalForm(board, player), copy=True)
        probs, q, full = m.getActionProb(cb, temp=1, force_full_search=True)
        raw = np.array(m.nodes_data[g.stringRepresentation(cb)][5], dtype=np.int64)
        action = int(np.random.choice(A, p=np.array(probs) / np.sum(probs)))
        nb, nplayer = g.getNextState(board, player, action)
        nz = np.flatnonzero(raw)
        pad = np.zeros(128, np.int64); padi = np.full(128, -1, np.int64); pad[:len(nz)] = raw[nz]; padi[:len(nz)] = nz
        roots.append(cb); idxs.append(padi); cnts.append(pad); qs.append(np.array(q, dtype=np.float32)); actions.append(action); summaries.append(tree_summary(m))
        board, player = np.array(nb, copy=True), nplayer
    np.savez_compressed(os.path.join(out, 'abalone_episode.npz'), n_sims=np.array(nsims), roots=np.array(roots), raw_idx=np.array(idxs), raw_cnt=np.array(cnts),
                        q=np.array(qs), actions=np.array(actions), summaries=np.array(summaries))
    print(f'abalone episode: {len(roots)} plies, last summary={summaries[-1]}')


def gen_net(out):
    """AbaloneNNet V21 forward (abalone/AbaloneNNet.py:117-156,173-202) through the reference's torch branch of predict."""
    import torch
    torch.set_num_threads(1)
    from abalone.AbaloneGame import AbaloneGame
    from abalone.NNet import NNetWrapper
    kat = np.load(os.path.join(out, 'abalone_kat.npz'))
    g = AbaloneGame()
    nn_args = dict(nn_version=21, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32, no_compression=True, q_weight=0.5)
    sel = np.linspace(0, len(kat['canonical']) - 1, 48).astype(int)
    boards = kat['canonical'][sel]; valids = np.unpackbits(kat['valids'][sel], axis=1)[:, :A].astype(np.bool_)
    for tag in ('rand', 'shipped'):
        torch.manual_seed(0)
        w = NNetWrapper(g, nn_args)
        w.device['inference'] = 'cpu'
        if tag == 'rand':
            gen = torch.Generator().manual_seed(3)
            with torch.no_grad():
                for mod in w.nnet.modules():
                    if isinstance(mod, torch.nn.BatchNorm2d):
                        mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.3)
                        mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) * 1.5 + 0.25)
                        mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=gen) + 0.5)
                        mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=gen) * 0.2)
        else:
            ck = torch.load('/root/reference/abalone/pretrained_BelgianDaisy.pt', map_location='cpu', weights_only=False)
            w.nnet.load_state_dict(ck['state_dict'])
        w.nnet.eval()
        pis, vs = [], []
        for b, v in zip(boards, valids):
            pi, val = w.predict(b, v)
            pis.append(pi); vs.append(val)
        sd = {k: t.detach().cpu().numpy() for k, t in w.nnet.state_dict().items()}
        save = {'sd__' + k: v for k, v in sd.items()}
        save.update(boards=boards, valids=np.packbits(valids, axis=1), pi=np.array(pis, dtype=np.float32), v=np.array(vs, dtype=np.float32))
        np.savez_compressed(os.path.join(out, f'abalone_v21_{tag}.npz'), **save)
        print(f'abalone net {tag}: {len(sd)} tensors, {sum(v.size for v in sd.values())} values, pi[0] max={pis[0].max():.4f} v[0]={vs[0]}')
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
    if 'mcts1600' in only:
        gen_mcts1600(a.out)
    if 'episode' in only:
        gen_episode(a.out)
    if 'net' in only:
        gen_net(a.out)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE REFERENCE ITSELF (test infrastructure).

Runs only in the build container, where /root/reference exists; the GPU box never
sees the reference, so the vectors produced here are committed under
tests/golden/ together with this script.

    python oracle/gen_golden.py [--out tests/golden] [--only kat,mcts,net,episode]

What is imported from the reference (unmodified, read-only):
  splendor/SplendorGame.py  (-> SplendorLogicNumba.Board jitclass)   game step
  MCTS.py                                                             tree search
  splendor/NNet.py / SplendorNNet.py (torch CPU branch)               V80 forward
Shims: oracle/ref_shim/{colorama,onnx,onnxruntime} are empty stand-ins for modules
absent from this image; inference uses the reference's torch branch
(GenericNNetWrapper.py:111-120) by setting wrapper.device['inference']='cpu'.
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference', HERE]

import numpy as np  # noqa: E402
from numba import njit  # noqa: E402

from hashnet import HashNet  # noqa: E402


@njit
def _seed_numba(s):
    np.random.seed(s)


class dotdict(dict):
    def __getattr__(self, name):
        return self[name]


SEEDS = [-1, 31416, 1, 14142, 42, 27183, 2, 16180, 7, 123456789]


def pick_action(rng, valids):
    """Biased random mover so playouts exercise buys / reserves / nobles, not only gem moves."""
    w = np.zeros(81)
    w[0:12] = 8.0
    w[12:27] = 2.0
    w[27:30] = 8.0
    w[30:60] = 3.0
    w[60:80] = 0.4
    w[80] = 0.02
    w = w * valids
    return int(rng.choice(81, p=w / w.sum()))


def gen_kat(out, n_games=8):
    """Per-ply known-answer vectors for the Splendor game step (SplendorLogicNumba.py)."""
    from splendor.SplendorGame import SplendorGame
    g = SplendorGame()
    rng = np.random.default_rng(2026)
    rec = {k: [] for k in ('board', 'player', 'canonical', 'valids', 'action', 'seed', 'next_board',
                           'next_player', 'ended', 'round', 'score', 'game', 'next_canonical')}
    sym = {k: [] for k in ('board', 'pi', 'valids', 'k', 'out_boards', 'out_pi', 'out_valids')}
    inits = []
    for ep in range(n_games):
        _seed_numba(1000 + ep)
        board = g.getInitBoard().copy()
        inits.append(board.copy())
        player, ply = 0, 0
        while True:
            cb = np.array(g.getCanonicalForm(board, player), copy=True)
            valids = np.array(g.getValidMoves(cb, 0), copy=True)
            valids_abs = np.array(g.getValidMoves(board, player), copy=True)
            assert (valids == valids_abs).all()
            action = pick_action(rng, valids)
            seed = SEEDS[(ply + ep) % len(SEEDS)]
            nb, nplayer = g.getNextState(board, player, action, random_seed=seed)
            nb = np.array(nb, copy=True)
            ended = np.array(g.getGameEnded(nb, nplayer), copy=True)
            rec['board'].append(board.copy()); rec['player'].append(player)
            rec['canonical'].append(cb); rec['valids'].append(valids)
            rec['action'].append(action); rec['seed'].append(seed)
            rec['next_board'].append(nb); rec['next_player'].append(nplayer)
            rec['ended'].append(ended); rec['round'].append(int(g.getRound(nb)))
            rec['score'].append([int(g.getScore(nb, 0)), int(g.getScore(nb, 1))])
            rec['game'].append(ep)
            rec['next_canonical'].append(np.array(g.getCanonicalForm(nb, nplayer), copy=True))
            if ply % 3 == 0:
                pi = rng.random(81).astype(np.float32)
                s = g.getSymmetries(cb, pi, valids)
                K = len(s)
                ob = np.zeros((14, 56, 7), np.int8); op = np.zeros((14, 81), np.float32); ov = np.zeros((14, 81), np.bool_)
                for i, (b_, p_, v_) in enumerate(s):
                    ob[i], op[i], ov[i] = b_, p_, v_
                sym['board'].append(cb); sym['pi'].append(pi); sym['valids'].append(valids); sym['k'].append(K)
                sym['out_boards'].append(ob); sym['out_pi'].append(op); sym['out_valids'].append(ov)
            board, player, ply = nb, nplayer, ply + 1
            if ended.any():
                break
    arrs = {k: np.array(v) for k, v in rec.items()}
    arrs['board'] = arrs['board'].astype(np.int8)
    arrs['seed'] = arrs['seed'].astype(np.int64)
    arrs['init_boards'] = np.array(inits, dtype=np.int8)
    for k, v in sym.items():
        arrs['sym_' + k] = np.array(v)
    np.savez_compressed(os.path.join(out, 'splendor_kat.npz'), **arrs)
    n_end = int((np.abs(arrs['ended']).sum(axis=1) > 0).sum())
    print(f'kat: {len(arrs["action"])} plies, {n_end} terminal, actions hist buy={int((arrs["action"]<12).sum())} '
          f'reserve={int(((arrs["action"]>=12)&(arrs["action"]<27)).sum())} buyres={int(((arrs["action"]>=27)&(arrs["action"]<30)).sum())} '
          f'sym={len(sym["k"])}')


class RecordingRng:
    """Wraps MCTS.rng (MCTS.py:43) so the Dirichlet draws and PCR coin flips can be replayed."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.dirichlets = []
        self.randoms = []

    def random(self):
        x = self.rng.random()
        self.randoms.append(x)
        return x

    def dirichlet(self, alpha):
        x = self.rng.dirichlet(alpha)
        self.dirichlets.append(np.array(x, dtype=np.float64))
        return x


MCTS_CONFIGS = {
    # main.py:118-156 defaults
    'default': dict(cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                    forced_playouts=False, prob_fullMCTS=1.0, ratio_fullMCTS=5, no_mem_optim=False, noise=False),
    # args stored in splendor/pretrained_2players.pt
    'shipped': dict(cpuct=0.8, fpu=0.0593, universes=3, dirichletAlpha=0.3, temperature=[1.25, 0.8, 1.1],
                    forced_playouts=True, prob_fullMCTS=1.0, ratio_fullMCTS=5, no_mem_optim=False, noise=True),
    'auto_noise': dict(cpuct=1.25, fpu=0.2, universes=0, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                       forced_playouts=False, prob_fullMCTS=1.0, ratio_fullMCTS=5, no_mem_optim=False, noise=True),
    'universes8': dict(cpuct=2.0, fpu=0.0, universes=8, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.0],
                       forced_playouts=True, prob_fullMCTS=1.0, ratio_fullMCTS=5, no_mem_optim=False, noise=False),
}


def tree_summary(mcts):
    n_nodes = len(mcts.nodes_data)
    n_term = sum(1 for v in mcts.nodes_data.values() if v[2] is None)
    sum_ns = sum(int(v[3]) for v in mcts.nodes_data.values() if v[2] is not None)
    return np.array([n_nodes, n_term, sum_ns], dtype=np.int64)


def gen_mcts(out):
    """Root visit counts of the reference MCTS (MCTS.py:49-184) driven by the hash-net."""
    from splendor.SplendorGame import SplendorGame
    from MCTS import MCTS
    kat = np.load(os.path.join(out, 'splendor_kat.npz'))
    g = SplendorGame()
    net = HashNet(g)
    cases = []
    game_ids = kat['game']
    # roots: a spread of canonical positions (early / mid / late / near the end)
    idx_by_game = {ep: np.flatnonzero(game_ids == ep) for ep in np.unique(game_ids)}
    picks = []
    for ep in (0, 1, 2, 3):
        idx = idx_by_game[ep]
        for frac in (0.0, 0.25, 0.6, 0.97):
            picks.append(int(idx[min(len(idx) - 1, int(frac * len(idx)))]))
    for ci, (name, cfg) in enumerate(MCTS_CONFIGS.items()):
        for pi_, p in enumerate(picks):
            if name != 'default' and pi_ % 2 == 1:
                continue
            n_sims = 800 if (name == 'default' and pi_ in (0, 5)) else 200
            args = dotdict(cfg, numMCTSSims=n_sims)
            m = MCTS(g, net, args, dirichlet_noise=cfg['noise'])
            rr = RecordingRng(100 * ci + pi_)
            m.rng = rr
            root = np.array(kat['canonical'][p], copy=True)
            probs, q, full = m.getActionProb(root, temp=1, force_full_search=True)
            s = g.stringRepresentation(root)
            raw = np.array([m.nodes_data[s][5][a] for a in range(81)], dtype=np.int64)
            case = dict(cfg=name, root=root, n_sims=n_sims, probs=np.array(probs, dtype=np.float64),
                        q=np.array(q, dtype=np.float32), raw_counts=raw, root_P=np.array(m.nodes_data[s][2], dtype=np.float32),
                        root_Qsa=np.array(m.nodes_data[s][4], dtype=np.float64),
                        noise=(rr.dirichlets[0] if rr.dirichlets else np.zeros(0)), summary=tree_summary(m))
            cases.append(case)
            print(f'mcts {name} root#{p} n={n_sims} nodes={case["summary"]} top={int(np.argmax(raw))}:{int(raw.max())}')
    save = {'n_cases': np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            save[f'c{i}_{k}'] = np.array(v)
    np.savez_compressed(os.path.join(out, 'splendor_mcts.npz'), **save)


def gen_episode(out):
    """Tree reuse across moves (MCTS.py:67-68,86-91): one self-play game, hash-net, 120 sims/move,
    real moves through getNextState(random_seed=0) exactly as Coach.executeEpisode (Coach.py:55-84)."""
    from splendor.SplendorGame import SplendorGame
    from MCTS import MCTS
    g = SplendorGame()
    net = HashNet(g)
    for name, cfg_name, nsims in (('A', 'default', 120), ('B', 'shipped', 90)):
        cfg = MCTS_CONFIGS[cfg_name]
        args = dotdict(cfg, numMCTSSims=nsims)
        _seed_numba(77 if name == 'A' else 78)
        np.random.seed(5)
        m = MCTS(g, net, args, dirichlet_noise=cfg['noise'])
        rr = RecordingRng(9)
        m.rng = rr
        board = g.getInitBoard().copy()
        player = 0
        roots, counts, probs_l, qs, actions, players, summaries, noises, next_boards = [], [], [], [], [], [], [], [], []
        enders = []
        while True:
            cb = np.array(g.getCanonicalForm(board, player), copy=True)
            nd = len(rr.dirichlets)
            probs, q, full = m.getActionProb(cb, temp=1, force_full_search=True)
            s = g.stringRepresentation(cb)
            raw = np.array([m.nodes_data[s][5][a] for a in range(81)], dtype=np.int64)
            # move choice: sample from probs (temperature 1) with the global numpy RNG as Coach.py:291
            action = int(np.random.choice(81, p=np.array(probs) / np.sum(probs)))
            nb, nplayer = g.getNextState(board, player, action)       # random_seed=0: true random draw
            nb = np.array(nb, copy=True)
            r = np.array(g.getGameEnded(nb, nplayer), copy=True)
            roots.append(cb); counts.append(raw); probs_l.append(np.array(probs)); qs.append(np.array(q, dtype=np.float32))
            actions.append(action); players.append(player); summaries.append(tree_summary(m)); next_boards.append(nb)
            noises.append(rr.dirichlets[nd] if len(rr.dirichlets) > nd else np.zeros(0))
            enders.append(r)
            board, player = nb, nplayer
            if r.any():
                break
        L = max(len(x) for x in noises)
        noise_arr = np.zeros((len(noises), max(L, 1)))
        noise_len = np.zeros(len(noises), dtype=np.int64)
        for i, x in enumerate(noises):
            noise_arr[i, :len(x)] = x
            noise_len[i] = len(x)
        np.savez_compressed(os.path.join(out, f'splendor_episode_{name}.npz'), cfg=np.array(cfg_name), n_sims=np.array(nsims),
                            roots=np.array(roots), raw_counts=np.array(counts), probs=np.array(probs_l), q=np.array(qs),
                            actions=np.array(actions), players=np.array(players), summaries=np.array(summaries),
                            next_boards=np.array(next_boards), noise=noise_arr, noise_len=noise_len, ended=np.array(enders))
        print(f'episode {name}: {len(roots)} plies, final r={enders[-1]}, last summary={summaries[-1]}')


def _randomise_bn(model, gen):
    import torch
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.3)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) * 1.5 + 0.25)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=gen) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=gen) * 0.2)


def gen_net(out):
    """SplendorNNet V80 forward (splendor/SplendorNNet.py:259-280,397-404,440) through the reference's
    torch branch of GenericNNetWrapper.predict (GenericNNetWrapper.py:111-120)."""
    import torch
    torch.set_num_threads(1)
    from splendor.SplendorGame import SplendorGame
    from splendor.NNet import NNetWrapper
    kat = np.load(os.path.join(out, 'splendor_kat.npz'))
    g = SplendorGame()
    nn_args = dict(nn_version=80, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32,
                   no_compression=True, q_weight=0.5)
    sel = np.linspace(0, len(kat['canonical']) - 1, 96).astype(int)
    boards = kat['canonical'][sel]
    valids = kat['valids'][sel]
    for tag in ('rand', 'shipped'):
        torch.manual_seed(0)
        w = NNetWrapper(g, nn_args)
        w.device['inference'] = 'cpu'
        if tag == 'rand':
            gen = torch.Generator().manual_seed(1)
            with torch.no_grad():
                _randomise_bn(w.nnet, gen)
                # biases are zero-initialised by the reference (_init); make them non-trivial too
                for name, p in w.nnet.named_parameters():
                    if name.endswith('.bias') and 'norm' not in name:
                        p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        else:
            ck = torch.load('/root/reference/splendor/pretrained_2players.pt', map_location='cpu', weights_only=False)
            w.nnet.load_state_dict(ck['state_dict'])
        w.nnet.eval()
        pis, vs = [], []
        for b, v in zip(boards, valids):
            pi, val = w.predict(b, v)
            pis.append(pi); vs.append(val)
        sd = {k: t.detach().cpu().numpy() for k, t in w.nnet.state_dict().items()}
        save = {'sd__' + k: v for k, v in sd.items()}
        save.update(boards=boards, valids=valids, pi=np.array(pis, dtype=np.float32), v=np.array(vs, dtype=np.float32))
        np.savez_compressed(os.path.join(out, f'splendor_v80_{tag}.npz'), **save)
        print(f'net {tag}: {len(sd)} tensors, {sum(v.size for v in sd.values())} values, pi[0] max={pis[0].max():.4f} v[0]={vs[0]}')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    ap.add_argument('--only', default='kat,mcts,episode,net')
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
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

#!/usr/bin/env python
"""Golden fixtures for Splendor with 3 and 4 players by RUNNING THE REFERENCE (test infrastructure).

The player count is a source-level constant of the reference (`NUMBER_PLAYERS = 2`, splendor/SplendorGame.py:9), so -- like
gen_golden_santorini.py -- this script copies `splendor/` to a scratch directory OUTSIDE the repo, edits that one constant there and
imports the patched copy next to the unmodified top-level modules. Only the vectors are committed.

    python oracle/gen_golden_splendor_np.py --players 3 [--out tests/golden]
Writes splendor{n}p_kat.npz (rules: valid moves, next state with deterministic draws, end, round, score, canonical form, symmetries),
splendor{n}p_mcts.npz (reference MCTS root counts with the hash-net), splendor{n}p_v80_shipped.npz (SplendorNNet V80 forward of the
shipped pretrained_{n}players.pt through the reference's torch branch), splendor{n}p_selfplay.npz (Coach.executeEpisode with every
random input recorded, see gen_golden_selfplay.py)."""
import argparse
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('OMP_NUM_THREADS', '1')


def patched_reference(n):
    scratch = f'/tmp/azg_ref_splendor_{n}p'
    os.environ['NUMBA_CACHE_DIR'] = f'/tmp/numba_cache_splendor_{n}p'
    if os.path.isdir(scratch):
        shutil.rmtree(scratch)
    os.makedirs(scratch)
    shutil.copytree('/root/reference/splendor', os.path.join(scratch, 'splendor'), ignore=shutil.ignore_patterns('*.pt', '*.gif', '*.mp4', '*.png', '__pycache__'))
    p = os.path.join(scratch, 'splendor', 'SplendorGame.py')
    src = open(p).read()
    new, k = re.subn(r'^NUMBER_PLAYERS = 2\b', f'NUMBER_PLAYERS = {n}', src, flags=re.M)
    assert k == 1, 'NUMBER_PLAYERS constant not found'
    open(p, 'w').write(new)
    sys.path[:0] = [scratch, os.path.join(HERE, 'ref_shim'), '/root/reference', HERE]
    return scratch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--players', type=int, required=True)
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    a = ap.parse_args()
    n = a.players
    patched_reference(n)
    import numpy as np
    from numba import njit
    from hashnet import HashNet
    from gen_golden import MCTS_CONFIGS, RecordingRng, dotdict, tree_summary, pick_action, SEEDS
    sys.path.insert(0, f'/tmp/azg_ref_splendor_{n}p')
    from splendor.SplendorGame import SplendorGame
    from MCTS import MCTS
    import Coach as coach_mod

    @njit
    def seed_numba(s):
        np.random.seed(s)
    g = SplendorGame()
    R, A = g.getBoardSize()[0], g.getActionSize()
    assert g.num_players == n and R == 32 + 10 * n + n * n and A == 81
    KMAX = 1 + 9 + 2 * n
    # ---- rules
    rng = np.random.default_rng(3000 + n)
    rec = {k: [] for k in ('board', 'player', 'canonical', 'valids', 'action', 'seed', 'next_board', 'next_player', 'ended', 'round', 'score', 'game', 'next_canonical')}
    sym = {k: [] for k in ('board', 'pi', 'valids', 'k', 'out_boards', 'out_pi', 'out_valids')}
    for ep in range(4):
        seed_numba(1100 + 10 * n + ep)
        board = g.getInitBoard().copy(); player, ply = 0, 0
        while True:
            cb = np.array(g.getCanonicalForm(board, player), copy=True)
            valids = np.array(g.getValidMoves(cb, 0), copy=True)
            assert (valids == np.array(g.getValidMoves(board, player))).all()
            action = pick_action(rng, valids); seed = SEEDS[(ply + ep) % len(SEEDS)]
            nb, nplayer = g.getNextState(board, player, action, random_seed=seed); nb = np.array(nb, copy=True)
            ended = np.array(g.getGameEnded(nb, nplayer), copy=True)
            rec['board'].append(board.copy()); rec['player'].append(player); rec['canonical'].append(cb); rec['valids'].append(valids)
            rec['action'].append(action); rec['seed'].append(seed); rec['next_board'].append(nb); rec['next_player'].append(nplayer)
            rec['ended'].append(ended); rec['round'].append(int(g.getRound(nb))); rec['score'].append([int(g.getScore(nb, p)) for p in range(n)])
            rec['game'].append(ep); rec['next_canonical'].append(np.array(g.getCanonicalForm(nb, nplayer), copy=True))
            if ply % 5 == 0:
                pi = rng.random(A).astype(np.float32); s = g.getSymmetries(cb, pi, valids)
                ob = np.zeros((KMAX, R, 7), np.int8); op = np.zeros((KMAX, A), np.float32); ov = np.zeros((KMAX, A), np.bool_)
                for i, (b_, p_, v_) in enumerate(s):
                    ob[i], op[i], ov[i] = b_, p_, v_
                sym['board'].append(cb); sym['pi'].append(pi); sym['valids'].append(valids); sym['k'].append(len(s))
                sym['out_boards'].append(ob); sym['out_pi'].append(op); sym['out_valids'].append(ov)
            board, player, ply = nb, nplayer, ply + 1
            if ended.any():
                break
    arrs = {k: np.array(v) for k, v in rec.items()}
    arrs['board'] = arrs['board'].astype(np.int8); arrs['seed'] = arrs['seed'].astype(np.int64)
    for k, v in sym.items():
        arrs['sym_' + k] = np.array(v)
    np.savez_compressed(os.path.join(a.out, f'splendor{n}p_kat.npz'), **arrs)
    print(f'{n}p kat: {len(arrs["action"])} plies, {int((np.abs(arrs["ended"]).sum(axis=1) > 0).sum())} terminal, sym {len(sym["k"])} (k up to {max(sym["k"])})')
    # ---- MCTS with the hash-net
    net = HashNet(g); cases = []
    idx = np.flatnonzero(arrs['game'] == 0)
    picks = [int(idx[min(len(idx) - 1, int(f * len(idx)))]) for f in (0.0, 0.3, 0.7, 0.97)]
    for ci, name in enumerate(('default', 'shipped')):
        cfg = MCTS_CONFIGS[name]
        for pi_, p in enumerate(picks):
            args = dotdict(cfg, numMCTSSims=150)
            m = MCTS(g, net, args, dirichlet_noise=cfg['noise']); rr = RecordingRng(100 * ci + pi_ + n); m.rng = rr
            root = np.array(arrs['canonical'][p], copy=True)
            probs, q, full = m.getActionProb(root, temp=1, force_full_search=True)
            s = g.stringRepresentation(root)
            raw = np.array([m.nodes_data[s][5][x] for x in range(A)], dtype=np.int64)
            cases.append(dict(cfg=name, root=root, n_sims=150, probs=np.array(probs, np.float64), q=np.array(q, np.float32), raw_counts=raw,
                              noise=(rr.dirichlets[0] if rr.dirichlets else np.zeros(0)), summary=tree_summary(m)))
    save = {'n_cases': np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            save[f'c{i}_{k}'] = np.array(v)
    np.savez_compressed(os.path.join(a.out, f'splendor{n}p_mcts.npz'), **save)
    print(f'{n}p mcts: {len(cases)} cases, q[0] = {cases[0]["q"]}')
    # ---- V80 forward of the shipped n-player checkpoint
    import torch
    torch.set_num_threads(1)
    from splendor.NNet import NNetWrapper
    w = NNetWrapper(g, dict(nn_version=80, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32, no_compression=True, q_weight=0.5))
    w.device['inference'] = 'cpu'
    ck = torch.load(f'/root/reference/splendor/pretrained_{n}players.pt', map_location='cpu', weights_only=False)
    w.nnet.load_state_dict(ck['state_dict']); w.nnet.eval()
    sel = np.linspace(0, len(arrs['canonical']) - 1, 48).astype(int)
    boards, valids = arrs['canonical'][sel], arrs['valids'][sel]
    pis, vs = zip(*[w.predict(b, v) for b, v in zip(boards, valids)])
    save = {'sd__' + k: t.detach().cpu().numpy() for k, t in w.nnet.state_dict().items() if not k.endswith('num_batches_tracked')}
    save.update(boards=boards, valids=valids, pi=np.array(pis, np.float32), v=np.array(vs, np.float32))
    np.savez_compressed(os.path.join(a.out, f'splendor{n}p_v80_shipped.npz'), **save)
    print(f'{n}p net: v[0] = {vs[0]}')
    # ---- Coach.executeEpisode with recorded randomness (one game)
    cfg = dict(numMCTSSims=30, cpuct=1.0, fpu=0.1, universes=2, dirichletAlpha=0.3, temperature=[1.25, 0.8, 1.1], tempThreshold=10,
               forced_playouts=True, prob_fullMCTS=0.4, ratio_fullMCTS=5, no_mem_optim=False, no_compression=True)
    args = dotdict(cfg); trace = dict(u_full=[], noise=[], u_move=[], chance_seed=[], is_full=[], action=[], root=[])
    crng = np.random.default_rng(555 + n)

    class HarnessGame(SplendorGame):
        def getInitBoard(self):
            b = SplendorGame.getInitBoard(self); trace['init'] = np.array(b, copy=True); return b

        def getNextState(self, board, player, action, random_seed=0):
            assert random_seed == 0
            seed = int(crng.integers(1, 2 ** 31 - 1)); trace['chance_seed'].append(seed); trace['action'].append(int(action))
            return SplendorGame.getNextState(self, board, player, action, random_seed=seed)

    class HarnessMCTS(MCTS):
        def getActionProb(self, cb, temp=1, force_full_search=False):
            nd, nr = len(self.rng.dirichlets), len(self.rng.randoms)
            trace['root'].append(np.array(cb, copy=True))
            res = MCTS.getActionProb(self, cb, temp=temp, force_full_search=force_full_search)
            trace['u_full'].append(self.rng.randoms[nr]); trace['noise'].append(self.rng.dirichlets[nd] if len(self.rng.dirichlets) > nd else np.zeros(0))
            trace['is_full'].append(bool(res[2]))
            rs = np.random.RandomState(); rs.set_state(np.random.get_state()); trace['u_move'].append(float(rs.random_sample()))
            return res
    hg = HarnessGame(); hg.getInitBoard()
    m = HarnessMCTS(hg, HashNet(hg), args, dirichlet_noise=True); m.rng = RecordingRng(900 + n)
    coach = coach_mod.Coach.__new__(coach_mod.Coach); coach.game = hg; coach.nnet = None; coach.args = args; coach.mcts = m; coach.nb_threads = 1
    seed_numba(800 + n); np.random.seed(850 + n)
    examples = coach.executeEpisode(m, hg)
    P = len(trace['u_full']); L = max(max((len(x) for x in trace['noise']), default=0), 1)
    noise = np.zeros((P, L)); noise_len = np.zeros(P, np.int64)
    for i, x in enumerate(trace['noise']):
        noise[i, :len(x)] = x; noise_len[i] = len(x)
    save = {'n_games': np.array(1)}
    for k, v in cfg.items():
        save['cfg_' + k] = np.array(v)
    save.update({'g0_init': trace['init'], 'g0_u_full': np.array(trace['u_full']), 'g0_u_move': np.array(trace['u_move']), 'g0_chance_seed': np.array(trace['chance_seed'], np.int64),
                 'g0_noise': noise, 'g0_noise_len': noise_len, 'g0_is_full': np.array(trace['is_full']), 'g0_action': np.array(trace['action'], np.int32),
                 'g0_root': np.array(trace['root'], np.int8), 'g0_ex_board': np.array([e[0] for e in examples], np.int8), 'g0_ex_pi': np.array([e[1] for e in examples], np.float32),
                 'g0_ex_valids': np.array([e[3] for e in examples], np.bool_), 'g0_ex_z': np.array([e[2] for e in examples], np.float32), 'g0_ex_q': np.array([e[4] for e in examples], np.float32)})
    np.savez_compressed(os.path.join(a.out, f'splendor{n}p_selfplay.npz'), **save)
    print(f'{n}p selfplay: {P} plies, {int(np.sum(trace["is_full"]))} full, {len(examples)} examples, z0 = {examples[0][2]}')


if __name__ == '__main__':
    main()

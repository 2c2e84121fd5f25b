#!/usr/bin/env python
"""Time the UNMODIFIED reference's own self-play path on this box's host cores, and the
C oracle port on the same cores, so every `cpu_baseline.kind == "port"` figure carries a measured port/reference factor.

Runs only where /root/reference exists (the build container); the GPU box never sees the reference.

    python scripts/time_reference_cpu.py [--sims 800] [--plies 0] [--procs 0] [--out profiles/r02_reference_cpu.json]

Legs (all Splendor-2p, SplendorNNet V80 random-init under torch.manual_seed(0), main.py default MCTS args, prob_fullMCTS=1.0):
  ref_1proc      Coach.executeEpisode, 1 process, OMP_NUM_THREADS=1, torch.set_num_threads(1)  (GenericNNetWrapper.py:7,22)
  ref_Pproc      P = os.cpu_count() independent processes of the same, summed (the author's scaling method, README.md:175-176)
  ref_stub_1proc the same tree + game logic with a stub net (no torch): upper bound on what a faster CPU inference could reach
  ref_pcr_1proc  prob_fullMCTS=0.25 / ratio 5 (the reference default)
  port_1thread / port_Pthreads   oracle/azg_oracle.c on the same workload
Deviation from the reference's shipped path, stated: inference runs through the reference's torch-CPU branch of `predict`
(GenericNNetWrapper.py:111-120), because onnxruntime is not installable offline.
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
os.environ['OMP_NUM_THREADS'] = '1'


def ref_args(sims, prob_full):
    return dict(numMCTSSims=sims, cpuct=1.25, fpu=0., universes=1, dirichletAlpha=-1, temperature=[1.0, 0.1, 1.1], tempThreshold=10,
                ratio_fullMCTS=5, prob_fullMCTS=prob_full, forced_playouts=False, no_mem_optim=False, parallel_inferences=1,
                no_compression=True, numEps=1, maxlenOfQueue=10 ** 6)


class _StubNet:
    """Uniform prior over legal moves, zero value (tree + game logic cost only)."""
    requestKnowledgeTransfer = False

    def __init__(self, game, args=None):
        import numpy as np
        self.np = np; self.args = args; self.A = game.getActionSize(); self.npl = game.num_players

    def predict(self, board, valids):
        np = self.np
        v = valids.astype(np.float32)
        return v / v.sum(), np.zeros(self.npl, np.float32)


def ref_leg(sims, prob_full, stub, max_plies, seed, q=None):
    """One process: warm-up (JIT) then one timed Coach.executeEpisode, truncated after max_plies plies if > 0."""
    sys.path[:0] = [os.path.join(ROOT, 'oracle', 'ref_shim'), '/root/reference']
    import numpy as np
    import torch
    from numba import njit
    torch.set_num_threads(1)

    @njit
    def seed_numba(s):
        np.random.seed(s)

    class dotdict(dict):
        __getattr__ = dict.__getitem__

    from splendor.SplendorGame import SplendorGame
    from MCTS import MCTS
    import Coach as coach_mod
    g = SplendorGame()
    if stub:
        net = _StubNet(g)
    else:
        from splendor.NNet import NNetWrapper
        torch.manual_seed(0)
        net = NNetWrapper(g, dotdict(nn_version=80, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32, no_compression=True, q_weight=0.5,
                                     cyclic_lr=False, vl_weight=1., surprise_weight=False, no_mem_optim=False, save_optim_state=False))
        net.device['inference'] = 'cpu'                                # torch branch of predict (GenericNNetWrapper.py:111-120)
        net.nnet.eval()
    args = dotdict(ref_args(sims, prob_full))

    class Stop(Exception):
        pass

    counters = dict(sims=0, plies=0, nodes=0)

    class CountingMCTS(MCTS):
        def getActionProb(self, cb, temp=1, force_full_search=False):
            if max_plies and counters['plies'] >= max_plies:
                raise Stop()
            out = MCTS.getActionProb(self, cb, temp=temp, force_full_search=force_full_search)
            counters['sims'] += self.step + 1; counters['plies'] += 1
            return out

    coach = coach_mod.Coach.__new__(coach_mod.Coach)                     # Coach.__init__ builds a second net (pnet): not needed for self-play
    coach.game = g; coach.nnet = net; coach.args = args; coach.nb_threads = 1

    def episode(limit, ep_args):
        nonlocal max_plies
        keep = max_plies; max_plies = limit
        counters.update(sims=0, plies=0)
        np.random.seed(seed); seed_numba(seed)
        coach.args = ep_args
        m = CountingMCTS(g, net, ep_args, dirichlet_noise=True); m.rng = np.random.default_rng(seed)
        coach.mcts = m
        t0 = time.perf_counter()
        try:
            coach.executeEpisode()
        except Stop:
            pass
        dt = time.perf_counter() - t0
        max_plies = keep
        return dict(seconds=dt, sims=counters['sims'], plies=counters['plies'], nodes=len(m.nodes_data))

    episode(2, dotdict(ref_args(sims, 1.0)))                               # JIT warm-up, untimed (full searches: compiles getSymmetries too)
    g.getScore(g.getInitBoard(), 0)                                        # only called at the end of a game otherwise
    r = episode(max_plies, args)
    r['sims_per_s'] = r['sims'] / r['seconds']
    if q is not None:
        q.put(r)
    return r


def run_procs(n, **kw):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=ref_leg, kwargs=dict(kw, seed=kw.get('seed', 0) + i, q=q)) for i in range(n)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    rs = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    return dict(procs=n, sims_per_s_sum=sum(r['sims_per_s'] for r in rs), per_proc=[r['sims_per_s'] for r in rs], wall_incl_jit_s=wall,
                sims=sum(r['sims'] for r in rs), plies=sum(r['plies'] for r in rs))


def port_leg(sims, plies, threads):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from azg_b200.nnet import random_v80_state_dict
    cfg = O.make_cfg(numMCTSSims=sims, net_kind=1, universes=1, prob_fullMCTS=1.0, cpuct=1.25, fpu=0.0, dirichletAlpha=-1.0, temperature2=1.1,
                     game=O.GAME_SPLENDOR)
    r = O.selfplay_bench(cfg, O.v80_blob(random_v80_state_dict(0)), threads, 1, max_plies=plies, temperature=[1.0, 0.1], tempThreshold=10, seed=1)
    return dict(threads=threads, sims=r['sims'], seconds=r['seconds'], sims_per_s=r['sims'] / r['seconds'])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sims', type=int, default=800)
    ap.add_argument('--plies', type=int, default=0, help='truncate the timed episode after this many plies (0 = a complete game)')
    ap.add_argument('--procs', type=int, default=0, help='P of the P-process leg (0 = os.cpu_count())')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r02_reference_cpu.json'))
    a = ap.parse_args()
    P = a.procs or os.cpu_count() or 1
    cpu = platform.processor() or ''
    try:
        cpu = [l.split(':', 1)[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')][0]
    except Exception:
        pass
    out = dict(workload=f'Splendor-2p, Coach.executeEpisode, numMCTSSims={a.sims}, V80 random-init seed 0, main.py default MCTS args, prob_fullMCTS=1.0'
                        + (f', truncated after {a.plies} plies' if a.plies else ', one complete game'),
               cpu_model=cpu, host_cores=os.cpu_count(), where='build container (no GPU)',
               deviation='inference through the reference torch-CPU branch of predict (onnxruntime not installable offline)')
    out['ref_1proc'] = run_procs(1, sims=a.sims, prob_full=1.0, stub=False, max_plies=a.plies)
    print(json.dumps(out['ref_1proc']), flush=True)
    out['ref_stub_1proc'] = run_procs(1, sims=a.sims, prob_full=1.0, stub=True, max_plies=a.plies)
    print(json.dumps(out['ref_stub_1proc']), flush=True)
    out['ref_pcr_1proc'] = run_procs(1, sims=a.sims, prob_full=0.25, stub=False, max_plies=a.plies)
    print(json.dumps(out['ref_pcr_1proc']), flush=True)
    out['ref_Pproc'] = run_procs(P, sims=a.sims, prob_full=1.0, stub=False, max_plies=a.plies)
    print(json.dumps(out['ref_Pproc']), flush=True)
    out['ref_stub_Pproc'] = run_procs(P, sims=a.sims, prob_full=1.0, stub=True, max_plies=a.plies)
    port_plies = a.plies or 12
    out['port_1thread'] = port_leg(a.sims, port_plies, 1)
    out['port_Pthreads'] = port_leg(a.sims, port_plies, P)
    out['port_over_reference'] = dict(one_core=out['port_1thread']['sims_per_s'] / out['ref_1proc']['sims_per_s_sum'],
                                      all_cores=out['port_Pthreads']['sims_per_s'] / out['ref_Pproc']['sims_per_s_sum'],
                                      note='multiply a reference sims/s by this factor to compare with a `kind: port` cpu_baseline on the same cores')
    json.dump(out, open(a.out, 'w'), indent=1)
    print(json.dumps(out['port_over_reference']))


if __name__ == '__main__':
    main()

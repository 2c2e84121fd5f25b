"""Command lines: azg_b200.main / azg_b200.pit mirror the flags, defaults and derived settings of the reference's
main.py:118-173 and pit.py:219-239 (values below are the reference's, copied from those lines); Coach.learn mirrors Coach.py:150-215.
CPU: parsing and the loop's control flow with stand-in nets. GPU (-m gpu): one real iteration on the device engine."""
import os

import numpy as np
import pytest

from azg_b200 import main as M, pit as P
from azg_b200.game_switcher import import_game, BUILT, NOT_BUILT


# main.py:120-156 (flag, dest, default)
REF_MAIN_DEFAULTS = dict(checkpoint='./temp/', load_folder_file=None, numEps=500, numItersHistory=5, numMCTSSims=1600, tempThreshold=10,
                         temperature=[1.0, 0.1, 1.1], cpuct=1.25, dirichletAlpha=-1, fpu=0., forced_playouts=False, learn_rate=0.0003,
                         epochs=2, batch_size=32, dropout=0., q_weight=0.5, updateThreshold=0.60, ratio_fullMCTS=5, prob_fullMCTS=0.25,
                         universes=1, forget_examples=False, numIters=50, stop_after_N_fail=-1, profile=False, debug=False, useray=False,
                         parallel_inferences=8, no_compression=False, no_mem_optim=False)
REF_MAIN_SHORT = {'-C': 'checkpoint', '-L': 'load_folder_file', '-e': 'numEps', '-i': 'numItersHistory', '-m': 'numMCTSSims', '-T': 'tempThreshold',
                  '-c': 'cpuct', '-d': 'dirichletAlpha', '-f': 'fpu', '-l': 'learn_rate', '-p': 'epochs', '-b': 'batch_size', '-D': 'dropout',
                  '-V': 'nn_version', '-q': 'q_weight', '-u': 'universes', '-n': 'numIters', '-s': 'stop_after_N_fail', '-P': 'parallel_inferences'}


def test_main_flags_and_defaults_match_reference():
    a = M.build_parser().parse_args(['splendor'])
    for k, v in REF_MAIN_DEFAULTS.items():
        assert getattr(a, k) == v, k
    for flag, dest in REF_MAIN_SHORT.items():
        b = M.build_parser().parse_args(['splendor', flag, '3'])
        assert float(getattr(b, dest)) == 3.0, flag
    assert M.build_parser().parse_args(['azul', '-F']).forced_playouts is True
    assert M.build_parser().parse_args(['azul', '-t', '1.25', '0.2', '1.0']).temperature == [1.25, 0.2, 1.0]


def test_main_derived_settings():
    a = M.derive(M.build_parser().parse_args(['splendor']))
    # main.py:159-163: arenaCompare 30, maxlenOfQueue = 2.5e6 / (0.5 * numItersHistory) with compression, -N fails = N * numItersHistory
    assert (a.arenaCompare, a.maxlenOfQueue, a.stop_after_N_fail, a.load_model, a.nn_version) == (30, 1000000, 5, False, 80)
    a = M.derive(M.build_parser().parse_args(['santorini', '--no-compression', '-i', '10', '-s', '-2', '-L', 'x/best.pt']))
    assert (a.maxlenOfQueue, a.stop_after_N_fail, a.load_model, a.nn_version) == (125000, 20, True, 89)
    a = M.derive(M.build_parser().parse_args(['azul', '--debug', '-P', '64']))
    assert (a.parallel_inferences, a.no_compression, a.no_mem_optim, a.nn_version) == (1, True, True, 84)     # main.py:165-168


def test_pit_flags():
    a = P.build_parser().parse_args(['abalone', 'a/best.pt', 'random', '-n', '12', '-m', '100', '-c', '1.5', '-f', '0.1'])
    assert (a.game, a.players, a.num_games, a.numMCTSSims, a.cpuct, a.fpu) == ('abalone', ['a/best.pt', 'random'], 12, 100, 1.5, 0.1)
    assert P.build_parser().parse_args(['splendor']).num_games == 30                                            # pit.py:221


def test_game_switcher_names():
    assert set(BUILT) | set(NOT_BUILT) == {'azul', 'botanik', 'minivilles', 'santorini', 'smallworld', 'splendor', 'thelittleprince', 'akropolis', 'abalone'}   # GameSwitcher.py:3-13
    with pytest.raises(NotImplementedError):
        import_game('smallworld')
    with pytest.raises(Exception, match='not known'):
        import_game('chess')


class _StubNet:
    """Stands in for an NNetWrapper: counts calls, 'weights' are one integer."""
    def __init__(self, game=None, args=None): self.args = dict(args or {}); self.w = 0; self.saved = {}; self.trained = 0
    def save_checkpoint(self, folder, filename, additional_keys={}): _StubNet.disk[(folder, filename)] = self.w
    def load_checkpoint(self, folder, filename): self.w = _StubNet.disk[(folder, filename)]
    def train(self, examples): self.w += 1; self.trained += 1
_StubNet.disk = {}


def _stub_coach(tmp_path, results, **kw):
    """A Coach whose engine-facing methods are replaced: the control flow of learn() is what is under test."""
    from azg_b200.coach import Coach
    from azg_b200.utils import with_defaults
    c = Coach.__new__(Coach)
    c.game = None; c.nnet = _StubNet(); c.trainExamplesHistory = []; c.skipFirstSelfPlay = False; c.consecutive_failures = 0
    c.args = with_defaults(dict(numIters=len(results), numItersHistory=2, checkpoint=str(tmp_path), updateThreshold=0.6, dirichletAlpha=-1, **kw))
    ex = (np.zeros((2, 2), np.int8), np.ones(3, np.float32) / 3, np.zeros(2, np.float32), np.ones(3, bool), [0., 0.])
    c.executeEpisodes = lambda: [ex, ex]
    c.saveTrainExamples = lambda: None
    it = iter(results)
    c.seen = []
    def pit(new, prev):
        from azg_b200.arena import accept_new_net
        c.seen.append((new.w, prev.w)); n, p, d = next(it)
        return n, p, d, accept_new_net(n, p, c.args.updateThreshold)
    c.pit = pit
    return c


def test_learn_accepts_and_rejects_like_the_reference(tmp_path):
    c = _stub_coach(tmp_path, [(20, 10, 0), (10, 20, 0), (0, 0, 30), (18, 12, 0)])
    rec = c.learn(log=lambda *a: None)
    assert [r['accepted'] for r in rec] == [True, False, False, True]                 # Coach.py:209: share of decisive games >= 0.6; no decisive game = reject
    # the competitor always holds the weights from before this iteration's training; a rejected net is rolled back to them
    assert c.seen == [(1, 0), (2, 1), (2, 1), (2, 1)]
    assert [r['examples'] for r in rec] == [2, 4, 4, 4]                              # history capped at numItersHistory = 2 iterations
    d = _StubNet.disk; t = str(tmp_path)
    assert d[(t, 'checkpoint_1.pt')] == 1 and d[(t, 'checkpoint_4.pt')] == 2 and d[(t, 'best.pt')] == 2 and (t, 'checkpoint_2.pt') not in d
    assert c.consecutive_failures == 0


def test_learn_stops_after_n_consecutive_fails(tmp_path):
    c = _stub_coach(tmp_path, [(0, 5, 0)] * 6, stop_after_N_fail=2)
    rec = c.learn(log=lambda *a: None)
    assert len(rec) == 2 and c.consecutive_failures == 2                              # Coach.py:212-214


@pytest.mark.gpu
def test_main_runs_one_iteration_on_the_engine(tmp_path):
    rec = M.main(['splendor', '-C', str(tmp_path), '-n', '1', '-e', '6', '-m', '12', '-P', '16', '-p', '1', '-b', '64', '--prob-fullMCTS', '1.0', '--useray'])
    assert len(rec) == 1 and rec[0]['examples'] > 0 and rec[0]['nwins'] + rec[0]['pwins'] + rec[0]['draws'] == 30
    names = set(os.listdir(tmp_path))
    assert {'temp.pt', 'checkpoint.examples', 'settings.txt'} <= names and (('best.pt' in names) == rec[0]['accepted'])
    # the files are the reference's formats: readable back through the same loaders, and by pit
    from azg_b200.formats import load_train_examples, load_checkpoint_file
    hist = load_train_examples(os.path.join(tmp_path, 'checkpoint.examples'), False, 5, 10 ** 6)
    assert len(hist) == 1 and len(hist[0]) == rec[0]['examples']
    assert load_checkpoint_file(os.path.join(tmp_path, 'temp.pt'))['numMCTSSims'] == 12
    one, two, draws = P.main(['splendor', os.path.join(tmp_path, 'temp.pt'), 'random', '-n', '8', '-m', '12'])
    assert one + two + draws == 8


def test_learn_resumes_from_loaded_examples(tmp_path):
    """Coach.py:160 with skipFirstSelfPlay (set by loadTrainExamples, Coach.py:262): the first iteration after a resume trains on the
    loaded history without playing; self-play starts again with the second iteration."""
    c = _stub_coach(tmp_path, [(20, 0, 0), (20, 0, 0)])
    played = []
    ex = c.executeEpisodes()
    c.executeEpisodes = lambda: (played.append(1), ex)[1]
    c.trainExamplesHistory = [list(ex), list(ex)]; c.skipFirstSelfPlay = True
    rec = c.learn(log=lambda *a: None)
    assert len(played) == 1 and [r['examples'] for r in rec] == [4, 4]          # iteration 1: the two loaded iterations; iteration 2: history capped at 2

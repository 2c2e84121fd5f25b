"""Self-play half of Coach (reference: Coach.py:37-148) on the device engine.

`Coach(game, nnet, args).executeEpisodes()` returns the same example tuples as the reference
`(board int8[56,7], pi float32[81], z float32[np], valids bool[81], q float32[np])`, symmetry-augmented like
Coach.py:67, for `args.numEps` finished games played by `args.parallel_inferences` (= n_games) concurrent slots.
"""
from collections import deque

import numpy as np

from .mcts import Engine
from .utils import with_defaults


class Coach:
    def __init__(self, game, nnet, args, n_games=None, seed=0, node_cap=0, edge_cap=0):
        self.game = game; self.nnet = nnet; self.args = a = with_defaults(args)
        self.n_games = int(n_games if n_games is not None else a.parallel_inferences)
        self.engine = Engine(game, nnet, a, self.n_games, dirichlet_noise=(a.dirichletAlpha != 0), seed=seed,
                             node_cap=node_cap, edge_cap=edge_cap)
        self.trainExamplesHistory = []

    def raw_examples(self, min_episodes, max_moves=0):
        """Plays until `min_episodes` games finished; returns un-augmented example arrays (one per full-search ply).
        azg_engine_selfplay returns early whenever the device example ring is more than half full; the ring is drained here between
        calls, so no example is lost however many episodes are asked for (the ring holds n_games * max_game_len examples)."""
        cap = self.n_games * self.game.info.max_game_len
        done0 = self.engine.stats()['episodes_finished']
        parts = []; moves = 0
        while True:
            left = min_episodes - (self.engine.stats()['episodes_finished'] - done0)
            if min_episodes > 0 and left <= 0:
                break
            m0 = self.engine.stats()['moves_played']
            self.engine.selfplay(min_episodes=max(left, 0), max_moves=max(max_moves - moves, 0) if max_moves else 0)
            parts.append(self.engine.examples(cap))
            moves += (self.engine.stats()['moves_played'] - m0 + self.n_games - 1) // self.n_games
            if min_episodes <= 0 or (max_moves and moves >= max_moves):
                break
        st = self.engine.stats()
        if st['examples_dropped'] or st['arena_overflows']:
            raise RuntimeError(f"self-play lost data: examples_dropped={st['examples_dropped']} (example ring full), "
                               f"arena_overflows={st['arena_overflows']} (tree arena full: raise node_cap / edge_cap)")
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(5)) if parts else self.engine.examples(cap)

    def augment(self, boards, pis, zs, valids, qs):
        """getSymmetries for every example (Coach.py:67-69) through the batched symmetry kernel."""
        out = []
        if len(boards) == 0:
            return out
        ob, op, ov, ok = self.game.symmetries_batch(boards, pis, valids)
        for i in range(len(boards)):
            for k in range(int(ok[i])):
                out.append((ob[i, k], op[i, k], zs[i], ov[i, k], [qs[i, p] for p in range(qs.shape[1])]))
        return out

    # ---- on-disk history (Coach.py:220-262) ----
    def saveTrainExamples(self):
        from .formats import save_train_examples
        return save_train_examples(self.trainExamplesHistory, self.args.checkpoint)

    def loadTrainExamples(self, examples_file=None):
        import os
        from .formats import load_train_examples
        path = examples_file or (os.path.dirname(self.args.load_folder_file) + '/checkpoint.examples')
        self.trainExamplesHistory = load_train_examples(path, self.args.no_compression, self.args.get('numItersHistory'), self.args.maxlenOfQueue)
        return self.trainExamplesHistory

    # ---- accept gate (Coach.py:194-215) ----
    def pit(self, new_nnet, prev_nnet, n_parallel=None):
        """Arena of arenaCompare games between MCTS over the new net and MCTS over the previous one, every game in flight at once.
        Returns (nwins, pwins, draws, accepted)."""
        from .arena import EngineArena, accept_new_net
        ar = EngineArena(self.game, new_nnet, prev_nnet, self.args, n_parallel=n_parallel or self.args.get('arenaCompare', 30))
        try:
            nwins, pwins, draws = ar.playGames(self.args.get('arenaCompare', 30))
        finally:
            ar.close()
        return nwins, pwins, draws, accept_new_net(nwins, pwins, self.args.get('updateThreshold', 0.55))

    def executeEpisode(self):
        """One finished game's examples (Coach.py:37-84); uses slot-parallel play and returns the first game that ends."""
        return self.executeEpisodes(num_eps=1)

    def executeEpisodes(self, num_eps=None):
        num_eps = int(self.args.numEps if num_eps is None else num_eps)
        ex = self.augment(*self.raw_examples(num_eps))
        q = deque([], maxlen=self.args.maxlenOfQueue)
        q += ex
        return q

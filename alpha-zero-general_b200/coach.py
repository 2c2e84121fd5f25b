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
        self.skipFirstSelfPlay = False                                  # Coach.py:33: set when examples were loaded from a checkpoint
        self.consecutive_failures = 0

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
        self.skipFirstSelfPlay = True                                   # Coach.py:262: the loaded history already holds this iteration's examples
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
        if not self.args.no_compression:                                # Coach.py:69: examples are kept as zlib(pickle(tuple)) unless --no-compression
            from .formats import compress_example
            ex = [compress_example(e) for e in ex]
        q = deque([], maxlen=self.args.maxlenOfQueue)
        q += ex
        return q

    # ---- the outer loop (Coach.py:150-215) ----
    @staticmethod
    def getCheckpointFile(iteration):
        return 'checkpoint_' + str(iteration) + '.pt'

    def learn(self, log=print):
        """numIters iterations of: self-play on the device engine (numEps finished games) -> history of the last numItersHistory
        iterations saved as checkpoint.examples -> the net is trained on the shuffled history (a copy of the old weights is kept as
        temp.pt) -> arena of arenaCompare games new vs previous on the engine, every game in flight at once -> the new net is kept
        (checkpoint_<i>.pt and best.pt) when it wins >= updateThreshold of the decisive games, else temp.pt is restored. After
        stop_after_N_fail consecutive rejections the loop stops (the reference exits the process). Returns one record per iteration."""
        import random
        a = self.args
        pnet = type(self.nnet)(self.game, dict(self.nnet.args))         # Coach.py:27: the competitor network
        records = []
        for i in range(1, int(a.get('numIters', 50)) + 1):
            if not self.skipFirstSelfPlay or i > 1:
                it = self.executeEpisodes()
                if len(it) == a.maxlenOfQueue:
                    log('saturation of elements in iterationTrainExamples, think about decreasing numEps or increasing maxlenOfQueue')
                self.trainExamplesHistory.append(it)
                if len(it) and a.dirichletAlpha > 0:                    # Coach.py:168-175: advice on the Dirichlet alpha
                    from .formats import decompress_example
                    avg_valid = float(np.mean([np.sum((x if a.no_compression else decompress_example(x))[3]) for x in it]))
                    if not (1 / 1.5 < a.dirichletAlpha / (10 / avg_valid) < 1.5):
                        log(f'There are about {avg_valid:.1f} valid moves per state, so I advise to set dirichlet to {10 / avg_valid:.1f} instead')
            if a.get('profile'):
                return records
            if len(self.trainExamplesHistory) > a.numItersHistory:
                self.trainExamplesHistory.pop(0)
            self.saveTrainExamples()
            train_examples = []
            for e in self.trainExamplesHistory:
                train_examples.extend(e)
            random.shuffle(train_examples)
            extra = {k: v for k, v in dict(a).items() if isinstance(v, (int, float, str, bool, list, tuple, type(None)))}
            self.nnet.save_checkpoint(folder=a.checkpoint, filename='temp.pt', additional_keys=extra)
            pnet.load_checkpoint(folder=a.checkpoint, filename='temp.pt')
            self.nnet.train(train_examples)
            nwins, pwins, draws, accepted = self.pit(self.nnet, pnet)
            if not accepted:
                self.consecutive_failures += 1
                log(f'Iter #{i} - new vs previous: {nwins}-{pwins}  ({draws} draws) --> REJECTED ({self.consecutive_failures})')
                self.nnet.load_checkpoint(folder=a.checkpoint, filename='temp.pt')
            else:
                log(f'Iter #{i} - new vs previous: {nwins}-{pwins}  ({draws} draws) --> ACCEPTED')
                self.nnet.save_checkpoint(folder=a.checkpoint, filename=self.getCheckpointFile(i), additional_keys=extra)
                self.nnet.save_checkpoint(folder=a.checkpoint, filename='best.pt', additional_keys=extra)
                self.consecutive_failures = 0
            records.append(dict(iteration=i, examples=len(train_examples), nwins=nwins, pwins=pwins, draws=draws, accepted=accepted))
            stop = a.get('stop_after_N_fail', -1)
            if not accepted and stop is not None and stop > 0 and self.consecutive_failures >= stop and i < a.get('numIters', 50):
                log('Exceeded threshold number of consecutive fails, stopping process')
                break
        return records

"""`dotdict`, the reference's `args` container (utils.py:20-22 of the reference)."""


class dotdict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


# main.py:118-156 defaults of the reference for the knobs MCTS / Coach read
DEFAULT_ARGS = dict(numMCTSSims=800, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                    tempThreshold=10, prob_fullMCTS=0.25, ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False,
                    no_compression=True, parallel_inferences=8, numEps=500, maxlenOfQueue=10 ** 9, arenaCompare=30, updateThreshold=0.55,
                    numItersHistory=5, checkpoint='./temp/')


def with_defaults(args):
    d = dotdict(DEFAULT_ARGS)
    if args is not None:
        d.update(dict(args) if isinstance(args, dict) else vars(args))
    return d

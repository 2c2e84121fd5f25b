"""Command line of the training loop: the reference's main.py (main.py:118-183) on the device engine.

    python -m azg_b200.main splendor -m 800 -e 4096 -P 4096 -u 3 -C ./temp/

Same flags, defaults and derived settings as the reference (arenaCompare = 30, maxlenOfQueue from numItersHistory, stop_after_N_fail
scaling, --debug); `--parallel-inferences` is the number of games in flight on the GPU (the reference's inference batch): thousands
are the useful setting here, 8 stays the default for parity. Extra: `--num-players` (Splendor 2..4) and `--seed`.
`--nn-version` defaults to the one version built per game (80 / 89 / 21 / 84) instead of the reference's 1.
"""
import argparse
import os

from .game_switcher import import_game, DEFAULT_NN_VERSION
from .utils import dotdict


def build_parser():
    p = argparse.ArgumentParser(description='AlphaZero training loop on the B200 self-play engine (flags of the reference main.py)')
    p.add_argument('game', action='store', default='splendor', help='game plugin: splendor, santorini, abalone or azul')
    p.add_argument('--checkpoint', '-C', action='store', default='./temp/', help='')
    p.add_argument('--load-folder-file', '-L', action='store', default=None, help='')
    p.add_argument('--numEps', '-e', action='store', default=500, type=int, help='finished self-play games per iteration')
    p.add_argument('--numItersHistory', '-i', action='store', default=5, type=int, help='')
    p.add_argument('--numMCTSSims', '-m', action='store', default=1600, type=int, help='simulations of a full search')
    p.add_argument('--tempThreshold', '-T', action='store', default=10, type=int, help='half-life (in moves) of the move-sampling temperature')
    p.add_argument('--temperature', '-t', action='store', default=[1.0, 0.1, 1.1], type=float, nargs=3,
                   help='sampling temperature at the start / at the end of a game, and the exponent applied to the root priors before the noise')
    p.add_argument('--cpuct', '-c', action='store', default=1.25, type=float, help='exploration constant of PUCT')
    p.add_argument('--dirichletAlpha', '-d', action='store', default=-1, type=float, help='Dirichlet alpha of the root noise: 0 = none, negative = 10 / number of legal moves')
    p.add_argument('--fpu', '-f', action='store', default=0., type=float, help='first-play urgency: > 0 subtracts from the parent value, < 0 is used as is')
    p.add_argument('--forced-playouts', '-F', action='store_true', help='forced playouts and policy-target pruning')
    p.add_argument('--learn-rate', '-l', action='store', default=0.0003, type=float, help='')
    p.add_argument('--epochs', '-p', action='store', default=2, type=int, help='')
    p.add_argument('--batch-size', '-b', action='store', default=32, type=int, help='')
    p.add_argument('--dropout', '-D', action='store', default=0., type=float, help='dropout of the trunk during training')
    p.add_argument('--nn-version', '-V', action='store', default=None, type=int, help='net architecture (default: the one built for the game)')
    p.add_argument('--q-weight', '-q', action='store', default=0.5, type=float, help='share of the search value Q in the value target')
    p.add_argument('--updateThreshold', action='store', default=0.60, type=float, help='share of the decisive arena games the new net must win')
    p.add_argument('--ratio-fullMCTS', action='store', default=5, type=int, help='a fast search runs numMCTSSims / this')
    p.add_argument('--prob-fullMCTS', action='store', default=0.25, type=float, help='probability that a move gets a full search (playout-cap randomisation)')
    p.add_argument('--universes', '-u', action='store', default=1, type=int, choices=range(9), help='chance seeds the search cycles through (0 = deterministic game)')
    p.add_argument('--forget-examples', action='store_true', help='with -L: do not load checkpoint.examples')
    p.add_argument('--numIters', '-n', action='store', default=50, type=int, help='')
    p.add_argument('--stop-after-N-fail', '-s', action='store', default=-1, type=float, help='stop after this many rejected nets in a row (-N = N * numItersHistory)')
    p.add_argument('--profile', action='store_true', help='self-play of one iteration only')
    p.add_argument('--debug', action='store_true', help='one game in flight, no compression, no tree clean-up')
    p.add_argument('--useray', action='store_true', help='accepted for compatibility (quieter output)')
    p.add_argument('--parallel-inferences', '-P', action='store', default=8, type=int, help='games in flight on the GPU = size of the inference batch')
    p.add_argument('--no-compression', action='store_true', help='keep examples uncompressed in memory and on disk')
    p.add_argument('--no-mem-optim', action='store_true', help='keep every tree node (no garbage collection between moves)')
    p.add_argument('--num-players', action='store', default=None, type=int, help='Splendor only: 2, 3 or 4 players')
    p.add_argument('--seed', action='store', default=0, type=int, help='seed of the device RNG streams')
    return p


def derive(args):
    """main.py:158-173: settings computed from the flags."""
    args.arenaCompare = 30
    args.maxlenOfQueue = int(2.5e6 / ((2 if args.no_compression else 0.5) * args.numItersHistory))
    if args.stop_after_N_fail < 0:
        args.stop_after_N_fail = -args.stop_after_N_fail * args.numItersHistory
    if args.debug:
        args.parallel_inferences = 1; args.no_compression = True; args.no_mem_optim = True
    if args.nn_version is None:
        args.nn_version = DEFAULT_NN_VERSION.get(args.game, 1)
    args.load_model = args.load_folder_file is not None
    return args


def run(args, log=print):
    from .coach import Coach
    Game, NNet, _ = import_game(args.game, args.num_players)
    g = Game()
    nn_args = dict(lr=args.learn_rate, dropout=args.dropout, epochs=args.epochs, batch_size=args.batch_size, nn_version=args.nn_version,
                   learn_rate=args.learn_rate, no_compression=args.no_compression, q_weight=args.q_weight)
    nnet = NNet(g, nn_args)
    if args.load_model:
        log(f'Loading checkpoint "{args.load_folder_file}"...')
        nnet.load_checkpoint(os.path.dirname(args.load_folder_file), os.path.basename(args.load_folder_file))
    c = Coach(g, nnet, dotdict(vars(args)), n_games=max(1, args.parallel_inferences), seed=args.seed)
    if args.load_model and not args.forget_examples:
        log("Loading 'trainExamples' from file...")
        c.loadTrainExamples()
    os.makedirs(args.checkpoint, exist_ok=True)
    settings = os.path.join(args.checkpoint, 'settings.txt')
    if os.path.isfile(settings):                                        # main.py:52-54: keep the previous settings beside the new ones
        import time
        os.replace(settings, os.path.join(args.checkpoint, 'settings.' + str(int(time.time()))))
    with open(settings, 'w') as f:
        f.write(str(args) + '\n')
    return c.learn(log=log)


def main(argv=None):
    args = derive(build_parser().parse_args(argv))
    if not args.useray:
        print(args)
    return run(args)


if __name__ == '__main__':
    main()

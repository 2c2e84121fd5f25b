"""Command line of the arena: the net-vs-net part of the reference's pit.py (pit.py:26-64,219-256) on the device engine.

    python -m azg_b200.pit splendor results/best.pt other/best.pt -n 30 -m 800

Players are checkpoint files (written by the reference or by this package) or the word "random" (a net with random-init weights).
Every game of the match is in flight at once (EngineArena); seats alternate 1-2-2-1 like Arena.playGames. The interactive players of
the reference (human, greedy), --display, ratings and the --compare folder scan are not part of the hot path and are not built.
"""
import argparse
import os

from .game_switcher import import_game, DEFAULT_NN_VERSION
from .utils import dotdict


def build_parser():
    p = argparse.ArgumentParser(description='tester of AlphaZero nets on the B200 engine (flags of the reference pit.py)')
    p.add_argument('--num-games', '-n', action='store', default=30, type=int, help='')
    p.add_argument('--numMCTSSims', '-m', action='store', default=None, type=int, help='simulations per move (default: the setting stored in the first checkpoint)')
    p.add_argument('--cpuct', '-c', action='store', default=None, type=float, help='exploration constant of PUCT')
    p.add_argument('--fpu', '-f', action='store', default=None, type=float, help='first-play urgency')
    p.add_argument('game', action='store', default='splendor', help='game plugin: splendor, santorini, abalone or azul')
    p.add_argument('players', metavar='player', nargs='*', help='two players: checkpoint files or "random"')
    p.add_argument('--num-players', action='store', default=None, type=int, help='Splendor only: 2, 3 or 4 players')
    p.add_argument('--universes', '-u', action='store', default=None, type=int, help='universes of the search (default: the checkpoint setting, else 1)')
    p.add_argument('--seed', action='store', default=0, type=int)
    return p


def create_player(name, game, NNet, nn_version, cli):
    """pit.py:26-64: a checkpoint's own MCTS settings are used unless a flag overrides them. Returns (nnet, settings)."""
    net = NNet(game, {'nn_version': nn_version})
    settings = {}
    if name != 'random':
        ck = net.load_checkpoint(os.path.dirname(name) or '.', os.path.basename(name))
        if ck is None:
            raise SystemExit(f'cannot load "{name}"')
        settings = {k: ck[k] for k in ('numMCTSSims', 'cpuct', 'fpu', 'universes') if ck.get(k) is not None}
    for k in ('numMCTSSims', 'cpuct', 'fpu', 'universes'):
        if getattr(cli, k) is not None:
            settings[k] = getattr(cli, k)
    return net, settings


def play(args, log=print):
    from .arena import EngineArena
    if len(args.players) != 2:
        raise SystemExit('give exactly two players (checkpoint files or "random")')
    Game, NNet, _ = import_game(args.game, args.num_players)
    g = Game()
    v = DEFAULT_NN_VERSION[args.game]
    (n1, s1), (n2, s2) = (create_player(p, g, NNet, v, args) for p in args.players)
    # one engine pair serves both players: the first player's search settings apply (pit.py plays each with its own; give flags to equalise)
    a = dotdict(dict(numMCTSSims=25, cpuct=1.0, fpu=0.0, universes=1, prob_fullMCTS=1.0, forced_playouts=False, no_mem_optim=False,
                     temperature=[1.0, 0.1, 1.1], tempThreshold=10, dirichletAlpha=0.0))
    a.update(s1)
    ar = EngineArena(g, n1, n2, a, n_parallel=args.num_games, seed=args.seed)
    try:
        one, two, draws = ar.playGames(args.num_games)
    finally:
        ar.close()
    log(f'{args.players[0]} vs {args.players[1]}: {one}-{two} ({draws} draws) over {args.num_games} games, {a.numMCTSSims} sims')
    return one, two, draws


def main(argv=None):
    return play(build_parser().parse_args(argv))


if __name__ == '__main__':
    main()

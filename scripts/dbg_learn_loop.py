"""Three iterations of the CLI training loop (self-play -> train -> arena) on small settings; prints the device memory in use after each
iteration (engine arenas of the arena contests and the competitor net must be released: no growth)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from azg_b200 import main as M
from azg_b200.coach import Coach
from azg_b200.game_switcher import import_game
from azg_b200.utils import dotdict
d = tempfile.mkdtemp()
args = M.derive(M.build_parser().parse_args([os.environ.get('GAME', 'splendor'), '-C', d, '-n', '3', '-e', '128', '-m', '40', '-P', '512', '-p', '1', '-b', '256', '--useray', '--updateThreshold', '0.5']))
Game, NNet, _ = import_game(args.game, args.num_players); g = Game()
nnet = NNet(g, dict(lr=args.learn_rate, dropout=0., epochs=1, batch_size=256, nn_version=args.nn_version, learn_rate=args.learn_rate, no_compression=False, q_weight=0.5))
c = Coach(g, nnet, dotdict(vars(args)), n_games=512, seed=1)
def used():
    free, total = torch.cuda.mem_get_info(); return (total - free) / 2 ** 20
log = []
def logger(*a): log.append(' '.join(str(x) for x in a)); print(*a, '| device memory in use %.0f MiB' % used(), flush=True)
rec = c.learn(log=logger)
print(rec); print(sorted(os.listdir(d)))

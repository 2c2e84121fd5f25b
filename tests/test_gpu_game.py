"""GPU parity of the batched Splendor step kernels (through the C ABI) against (a) golden vectors produced by the
reference and (b) the CPU oracle on fresh seeded inputs. Bit-exact: masks, boards, end vectors, rounds, scores, symmetries."""
import numpy as np
import pytest

import azg_b200
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def game():
    return azg_b200.SplendorGame()


def test_valid_moves_golden(game, kat):
    assert (game.valid_batch(kat['canonical']) == kat['valids']).all()
    assert (game.valid_batch(kat['board'], kat['player']) == kat['valids']).all()


def test_next_state_golden(game, kat):
    nb, npl = game.next_batch(kat['board'], kat['player'], kat['action'], kat['seed'])
    assert (npl == kat['next_player']).all()
    bad = np.flatnonzero((nb != kat['next_board']).reshape(len(nb), -1).any(axis=1))
    assert len(bad) == 0, f'first mismatch at ply {bad[:5]} actions {kat["action"][bad[:5]]}'


def test_ended_round_score_golden(game, kat):
    assert (game.ended_batch(kat['next_board']) == kat['ended']).all()
    rounds, scores = game.round_score_batch(kat['next_board'])
    assert (rounds == kat['round']).all() and (scores == kat['score']).all()


def test_canonical_golden(game, kat):
    assert (game.canonical_batch(kat['board'], kat['player']) == kat['canonical']).all()
    assert (game.canonical_batch(kat['next_board'], kat['next_player']) == kat['next_canonical']).all()


def test_symmetries_golden(game, kat):
    ob, op, ov, ok = game.symmetries_batch(kat['sym_board'], kat['sym_pi'], kat['sym_valids'])
    assert (ok == kat['sym_k']).all()
    for i in range(len(ok)):
        k = int(ok[i])
        assert (ob[i, :k] == kat['sym_out_boards'][i, :k]).all()
        assert (op[i, :k] == kat['sym_out_pi'][i, :k]).all()
        assert (ov[i, :k] == kat['sym_out_valids'][i, :k]).all()


def test_scalar_facade_matches_reference_surface(game, kat):
    i = 17
    b, p = kat['board'][i], int(kat['player'][i])
    cb = game.getCanonicalForm(b, p)
    assert (cb == kat['canonical'][i]).all() and cb.dtype == np.int8 and cb.shape == (56, 7)
    v = game.getValidMoves(cb, 0)
    assert v.dtype == np.bool_ and v.shape == (81,) and (v == kat['valids'][i]).all()
    nb, npl = game.getNextState(b, p, int(kat['action'][i]), random_seed=int(kat['seed'][i]))
    assert (nb == kat['next_board'][i]).all() and npl == kat['next_player'][i]
    r = game.getGameEnded(nb, npl)
    assert r.dtype == np.float32 and (r == kat['ended'][i]).all()
    assert game.getRound(nb) == kat['round'][i] and game.getScore(nb, 1) == kat['score'][i][1]
    assert game.stringRepresentation(nb) == kat['next_board'][i].tobytes()


def test_random_playouts_vs_oracle(game):
    """4096 concurrent random playouts, every ply checked against the CPU oracle (deterministic seeds)."""
    rng = np.random.default_rng(7)
    n = 512
    boards = game.init_batch(np.arange(1, n + 1, dtype=np.uint64))
    players = np.zeros(n, np.int32)
    alive = np.ones(n, bool)
    seeds_pool = np.array([-1, 31416, 1, 14142, 42, 27183, 2, 16180, 7, 99991], np.int64)
    for ply in range(130):
        if not alive.any():
            break
        idx = np.flatnonzero(alive)
        valids = game.valid_batch(boards[idx], players[idx])
        acts = np.array([rng.choice(np.flatnonzero(v)) for v in valids], np.int32)
        seeds = seeds_pool[rng.integers(0, len(seeds_pool), len(idx))]
        nb, npl = game.next_batch(boards[idx], players[idx], acts, seeds)
        ended = game.ended_batch(nb)
        for j in range(0, len(idx), 9):                     # oracle check on a stride of the batch (keeps the test fast)
            g = idx[j]
            assert (O.valid_moves(boards[g], int(players[g])) == valids[j]).all()
            ob, onp = O.next_state(boards[g], int(players[g]), int(acts[j]), int(seeds[j]))
            assert (ob == nb[j]).all() and onp == npl[j], (ply, g, acts[j])
            assert (O.game_ended(nb[j]) == ended[j]).all()
        boards[idx] = nb; players[idx] = npl
        alive[idx] = ~(ended != 0).any(axis=1)
    assert not alive.any(), 'every game must terminate by round 124'


def test_init_and_true_random_draws(game):
    """init / random_seed=0 use the device RNG: check invariants and that draws are uniform over the remaining deck."""
    boards = game.init_batch(np.arange(2000, dtype=np.uint64))
    assert (boards[:, 0] == np.array([4, 4, 4, 4, 4, 5, 0], np.int8)).all()
    assert (boards[:, 25, :5].sum(axis=1) == 36).all() and (boards[:, 27, :5].sum(axis=1) == 26).all() and (boards[:, 29, :5].sum(axis=1) == 16).all()
    assert (boards[:, 1:25:2, :5].sum(axis=2) > 0).all()
    assert (boards[:, 31:34, 6] == 3).all() and not boards[:, 34:].any()
    assert len({b.tobytes() for b in boards}) > 1990
    # first visible tier-3 card: each of the 20 cards about equally likely
    first = [bytes(b[17:19].tobytes()) for b in boards]
    counts = np.array(sorted(np.unique(first, return_counts=True)[1]))
    assert len(counts) == 20 and counts.min() > 55 and counts.max() < 150
    # same key => same draw; different keys => different draws (random_seed = 0)
    b0 = np.repeat(boards[:1], 64, axis=0)
    nb1, _ = game.next_batch(b0, np.zeros(64, np.int32), np.full(64, 24, np.int32), np.zeros(64, np.int64), np.arange(64, dtype=np.uint64))
    nb2, _ = game.next_batch(b0, np.zeros(64, np.int32), np.full(64, 24, np.int32), np.zeros(64, np.int64), np.arange(64, dtype=np.uint64))
    assert (nb1 == nb2).all() and len({b.tobytes() for b in nb1}) > 10


def test_empty_batch_is_ok(game):
    assert game.valid_batch(np.zeros((0, 56, 7), np.int8)).shape == (0, 81)

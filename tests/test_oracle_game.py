"""Pins the CPU oracle's Splendor step (oracle/azg_oracle.c) to vectors produced by the reference
itself (tests/golden/splendor_kat.npz, made by oracle/gen_golden.py). Bit-exact."""
import numpy as np

from oracle import oracle as O
from oracle.hashnet import hashnet_eval


def test_valid_moves_bit_exact(kat):
    for cb, v in zip(kat['canonical'], kat['valids']):
        assert (O.valid_moves(cb, 0) == v).all()
    # absolute boards, actual mover (exercises player=1 row offsets)
    for b, p, v in zip(kat['board'][::5], kat['player'][::5], kat['valids'][::5]):
        assert (O.valid_moves(b, int(p)) == v).all()


def test_next_state_bit_exact(kat):
    n = len(kat['action'])
    for i in range(n):
        nb, npl = O.next_state(kat['board'][i], kat['player'][i], kat['action'][i], kat['seed'][i])
        assert npl == kat['next_player'][i]
        assert (nb == kat['next_board'][i]).all(), f'ply {i} action {kat["action"][i]} seed {kat["seed"][i]}'


def test_game_ended_round_score(kat):
    for i in range(len(kat['action'])):
        nb = kat['next_board'][i]
        assert (O.game_ended(nb) == kat['ended'][i]).all()
        assert O.get_round(nb) == kat['round'][i]
        assert [O.get_score(nb, 0), O.get_score(nb, 1)] == list(kat['score'][i])


def test_canonical_form(kat):
    for i in range(len(kat['action'])):
        assert (O.canonical(kat['board'][i], kat['player'][i]) == kat['canonical'][i]).all()
        assert (O.canonical(kat['next_board'][i], kat['next_player'][i]) == kat['next_canonical'][i]).all()


def test_symmetries(kat):
    for i in range(len(kat['sym_k'])):
        s = O.symmetries(kat['sym_board'][i], kat['sym_pi'][i], kat['sym_valids'][i])
        assert len(s) == kat['sym_k'][i]
        for j, (b, p, v) in enumerate(s):
            assert (b == kat['sym_out_boards'][i][j]).all()
            assert (p == kat['sym_out_pi'][i][j]).all()
            assert (v == kat['sym_out_valids'][i][j]).all()


def test_init_game_invariants():
    """init_game uses randomness (numba MT19937 in the reference, SplendorLogicNumba.py:151-178) so only
    structural invariants can be compared: gem bank, 12 distinct visible cards, deck counts/bitfields, 3 nobles."""
    for seed in range(20):
        b = O.init_game(seed)
        assert list(b[0]) == [4, 4, 4, 4, 4, 5, 0]
        assert b[25, :5].sum() == 40 - 4 and b[27, :5].sum() == 30 - 4 and b[29, :5].sum() == 20 - 4
        for t in range(3):
            for c in range(5):
                assert bin(int(b[26 + 2 * t, c]) & 0xFF).count('1') == b[25 + 2 * t, c]
        assert all(b[1 + 2 * i, :5].sum() > 0 for i in range(12))
        nobles = {tuple(r) for r in b[31:34]}
        assert len(nobles) == 3 and all(r[6] == 3 for r in nobles)
        assert not b[34:].any()


def test_init_boards_match_reference_format(kat):
    # a reference initial board passes the same invariants (guards the invariant test itself)
    b = kat['init_boards'][0]
    assert list(b[0]) == [4, 4, 4, 4, 4, 5, 0] and b[25, :5].sum() == 36


def test_hashnet_matches_python(kat):
    for cb, v in zip(kat['canonical'][::7], kat['valids'][::7]):
        pi, val = O.hashnet(cb, v)
        pi2, val2 = hashnet_eval(cb, v)
        assert (pi == pi2).all() and (val == val2).all()
        assert abs(float(pi.sum()) - 1.0) == 0.0

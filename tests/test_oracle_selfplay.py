"""The oracle's Coach.executeEpisode (azo_execute_episode_inj) against the reference's own executeEpisode, example for example
(tests/golden/*_selfplay.npz recorded by oracle/gen_golden_selfplay.py running the unmodified Coach.executeEpisode with the hash-net
and every random input -- playout-cap coin, Dirichlet draw, move-sampling uniform, chance seed, initial board -- recorded)."""
import numpy as np
import pytest

from conftest import assert_examples_equal, load_selfplay_golden
from oracle import oracle as O

GAME_IDS = {'splendor': O.GAME_SPLENDOR, 'santorini': O.GAME_SANTORINI, 'abalone': O.GAME_ABALONE, 'azul': O.GAME_AZUL}


def oracle_cfg(game, cfg):
    return O.make_cfg(numMCTSSims=cfg['numMCTSSims'], ratio_fullMCTS=cfg['ratio_fullMCTS'], universes=cfg['universes'],
                      forced_playouts=cfg['forced_playouts'], net_kind=0, cpuct=cfg['cpuct'], fpu=cfg['fpu'], dirichletAlpha=cfg['dirichletAlpha'],
                      prob_fullMCTS=cfg['prob_fullMCTS'], temperature2=cfg['temperature'][2], game=GAME_IDS[game])


@pytest.mark.parametrize('game', ['splendor', 'santorini', 'abalone', 'azul'])
def test_oracle_episode_matches_reference_examples(game):
    cfg, games = load_selfplay_golden(game)
    for gd in games:
        ex = O.execute_episode_inj(oracle_cfg(game, cfg), gd['init'], gd['u_full'], gd['u_move'], gd['chance_seed'] if len(gd['chance_seed']) else None,
                                   noise=gd['noise'], temperature=cfg['temperature'][:2], tempThreshold=cfg['tempThreshold'])
        assert ex['plies'] == len(gd['u_full'])
        assert (ex['full'] == gd['is_full']).all()
        assert (ex['actions'] == gd['action']).all()
        assert_examples_equal(O.augment(GAME_IDS[game], ex), gd)

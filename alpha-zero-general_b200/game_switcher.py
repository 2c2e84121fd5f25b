"""GameSwitcher.import_game (GameSwitcher.py:3-26) for the games this engine has on the device."""

# reference package name -> (Game facade, NNetWrapper, default nn_version, NUMBER_PLAYERS)
BUILT = ('splendor', 'santorini', 'abalone', 'azul')
NOT_BUILT = ('botanik', 'minivilles', 'smallworld', 'thelittleprince', 'akropolis')


def import_game(pkg_name, num_players=None):
    """Returns (Game class or factory, NNetWrapper class, NUMBER_PLAYERS). Splendor takes num_players 2..4 (the reference edits
    NUMBER_PLAYERS in splendor/SplendorGame.py; here it is an argument)."""
    from . import game as G, nnet as N
    if pkg_name == 'splendor':
        n = int(num_players or 2)
        return (lambda: G.SplendorGame(n)), N.NNetWrapper, n
    if pkg_name == 'santorini':
        return G.SantoriniGame, N.SantoriniNNetWrapper, 2
    if pkg_name == 'abalone':
        return G.AbaloneGame, N.AbaloneNNetWrapper, 2
    if pkg_name == 'azul':
        return G.AzulGame, N.AzulNNetWrapper, 2
    if pkg_name in NOT_BUILT:
        raise NotImplementedError(f'Game {pkg_name} exists in the reference but has no device plugin yet; built: {list(BUILT)}')
    raise Exception(f'Game {pkg_name} is not known, please chose a game among the following list: {list(BUILT + NOT_BUILT)}')


DEFAULT_NN_VERSION = {'splendor': 80, 'santorini': 89, 'abalone': 21, 'azul': 84}

"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/azg.h declares, reports the
reference's sizes, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import azg_b200
from azg_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, 'include', 'azg.h')).read()
    declared = set(re.findall(r'\b(azg_[a-z0-9_]+)\s*\(', hdr))
    L = lib.load()
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.azg_abi_version() == lib.AZG_ABI_VERSION


def test_game_info_matches_reference_sizes():
    gi = lib.game_info(lib.AZG_GAME_SPLENDOR, 2)
    # splendor/SplendorLogicNumba.py:90-96: observation_size(2) = (56, 7), action_size() = 81
    assert (gi.state_rows, gi.state_cols, gi.state_bytes, gi.action_size) == (56, 7, 392, 81)
    assert gi.max_symmetries == 14 and gi.max_game_len == 124
    gs = lib.game_info(lib.AZG_GAME_SANTORINI, 2)
    # santorini/SantoriniLogicNumba.py:13-19 with NB_GODS = 1: observation_size() = (5, 5, 3), action_size() = 162
    assert (gs.state_rows, gs.state_cols, gs.state_depth, gs.state_bytes, gs.action_size, gs.max_symmetries) == (5, 5, 3, 75, 162, 8)
    ga = lib.game_info(lib.AZG_GAME_ABALONE, 2)
    # abalone/AbaloneLogicNumba.py:44-50: observation_size() = (9, 9, 4), action_size() = 3402; 12 symmetries
    assert (ga.state_rows, ga.state_cols, ga.state_depth, ga.state_bytes, ga.action_size, ga.max_symmetries) == (9, 9, 4, 324, 3402, 12)
    with pytest.raises(lib.AzgError):
        lib.game_info(99, 2)
    with pytest.raises(lib.AzgError):
        lib.game_info(lib.AZG_GAME_SPLENDOR, 5)


def test_v80_tensor_order_matches_golden_state_dict(v80_golden):
    sd = v80_golden['rand']['sd']
    assert all(n in sd for n in azg_b200.V80_TENSOR_ORDER)
    from azg_b200.nnet import v80_blob, random_v80_state_dict
    # every tensor except `lowvalue` and the 10 BatchNorm num_batches_tracked counters: 142406 parameters + running stats
    assert v80_blob(sd).size == 144881 - 11 == 142406 + 2 * (56 + 3 * (168 + 168 + 56))
    rnd = random_v80_state_dict(0)
    assert {k: v.shape for k, v in rnd.items()} == {k: sd[k].shape for k in azg_b200.V80_TENSOR_ORDER}


@pytest.mark.skipif(lib.device_count() > 0, reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    g = azg_b200.SplendorGame()
    with pytest.raises(lib.AzgError, match='no CPU fallback'):
        g.getInitBoard()
    with pytest.raises(lib.AzgError):
        g.getValidMoves(np.zeros((56, 7), np.int8), 0)


def test_v21_tile_plan_host_check(tmp_path):
    """net_v21.cuh: v21_plan (4-leaf tiles + a tail of 2-leaf tiles) covers every batch size; host-only program, no kernel launch."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    exe = str(tmp_path / 'v21_plan_check')
    src = os.path.join(ROOT, 'tests', 'host', 'v21_plan_check.cu')
    subprocess.run([nvcc, '-std=c++17', '-O1', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', exe, src], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == 'ok', out.stdout + out.stderr

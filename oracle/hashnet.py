"""Deterministic stand-in policy/value net ("hash-net") used ONLY by parity tests.

Test infrastructure (not shipped on the product path).  The same function is
implemented three times -- here (numpy, plugged into the reference's MCTS as its
`nnet`, see oracle/gen_golden.py), in oracle/azg_oracle.c and in the CUDA engine
(net kind AZG_NET_HASH) -- so that tree-search parity can be checked without the
float noise of a real network perturbing discrete PUCT choices
(SURVEY.md section 7 "Discrete-choice chaos").

Design: every prior is k/4096 with sum(k) == 4096, so the reference's
`normalise(Ps)` (MCTS.py:250-253, float32 sum + reciprocal multiply) is exact
whatever the summation order; the value is j/64.
"""
import numpy as np

M32 = 0xFFFFFFFF


def fmix32(h):
    h &= M32
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def fnv1a32(data: bytes):
    h = 0x811C9DC5
    for b in data:
        h = ((h ^ b) * 16777619) & M32
    return h


def hashnet_eval(board, valids, num_players=2):
    """board: int8 array (any shape), valids: bool[A] -> (pi f32[A], v f32[np])."""
    data = np.ascontiguousarray(board, dtype=np.int8).tobytes()
    valids = np.asarray(valids).astype(bool)
    A = valids.shape[0]
    h = fnv1a32(data)
    w = np.zeros(A, dtype=np.int64)
    for a in range(A):
        if valids[a]:
            w[a] = 256 + (fmix32((h + a * 0x9E3779B1) & M32) & 1023)
    W = int(w.sum())
    k = (w * 4096) // W
    rem = 4096 - int(k.sum())
    k[int(np.argmax(w))] += rem          # first index among the maxima
    pi = (k.astype(np.float32) / np.float32(4096.0)).astype(np.float32)
    j = int(fmix32(h ^ 0xABCDEF01) % 129) - 64
    v0 = np.float32(j) / np.float32(64.0)
    v = np.empty(num_players, dtype=np.float32)
    v[0] = v0
    v[1:] = -v0 / np.float32(num_players - 1)
    return pi, v


class HashNet:
    """Duck-typed `NeuralNet` for the reference MCTS (NeuralNet.py:27, MCTS.py:144)."""

    def __init__(self, game):
        self.num_players = game.num_players
        self.calls = 0

    def predict(self, board, valid_actions):
        self.calls += 1
        return hashnet_eval(board, valid_actions, self.num_players)

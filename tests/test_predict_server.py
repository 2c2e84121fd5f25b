"""predict_client / predict_server (GenericNNetWrapper.py:122-157): the reference's lock chain between N self-play threads and one
batching inference server, driven exactly as Coach.executeEpisodes sets it up (Coach.py:114-148), with a stub batch forward (no GPU)."""
import threading

import numpy as np

from azg_b200.nnet import BatchedPredictMixin


class StubNet(BatchedPredictMixin):
    def __init__(self):
        self.batches = []

    def predict_batch(self, boards, valids):
        self.batches.append(len(boards))
        b = np.asarray(boards, np.float32).reshape(len(boards), -1)
        pi = np.asarray(valids, np.float32); pi = pi / pi.sum(axis=1, keepdims=True)
        return pi, np.stack([b.sum(axis=1), -b.sum(axis=1)], axis=1)


def test_lock_chain_batches_all_threads():
    n = 4; net = StubNet()
    shared = [None] * (2 * n) + [0]
    locks = [threading.Lock() for _ in range(n + 1)]
    for l in locks:
        l.acquire()                                              # Coach.py:117-118
    results = [[] for _ in range(n)]

    def worker(i):
        locks[i].acquire()                                       # Coach.executeEpisodes_batch, Coach.py:91
        info = (i, i + n, shared, locks)
        for step in range(5):
            board = np.full((3, 2), i * 10 + step, np.int8); valids = np.array([1, 0, 1, 1], bool)
            pi, v = net.predict_client(board, valids, info)
            results[i].append((pi.copy(), v.copy()))
        while shared[-1] == 0:                                   # keep the chain turning until the stop signal
            locks[i + 1].release(); locks[i].acquire()
        locks[i + 1].release()

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
    srv = threading.Thread(target=net.predict_server, args=(n, shared, locks))
    for t in ths:
        t.start()
    srv.start()
    import time
    deadline = time.time() + 20
    while any(len(r) < 5 for r in results) and time.time() < deadline:
        time.sleep(0.01)
    shared[-1] = 2                                               # signal 2 = stop (Coach.py:139-144)
    for t in ths + [srv]:
        t.join(timeout=10)
    assert all(len(r) == 5 for r in results)
    assert all(b == n for b in net.batches)                      # every inference was a full batch of the n threads
    for i in range(n):
        for step, (pi, v) in enumerate(results[i]):
            assert np.allclose(pi, [1 / 3, 0, 1 / 3, 1 / 3]) and v[0] == 6 * (i * 10 + step)

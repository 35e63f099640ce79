"""TEST INFRASTRUCTURE ONLY. numpy/pure-Python restatement of the reference's prioritized replay arithmetic:
SegmentTree / SumSegmentTree.find_prefixsum_idx / MinSegmentTree (utils/segment_tree.py:13-151) and
PrioritizedReplayBuffer.add / sample_with_weights_and_idxes / update_priorities (buffer.py:128-189).
Pinned against the reference's OWN classes (they are pure Python and import here) in tests/golden/replay.npz."""
import numpy as np


class SegTree:
    def __init__(self, capacity, op, neutral):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.cap, self.op = capacity, op
        self.v = np.full(2 * capacity, neutral, dtype=np.float64)

    def set(self, idx, val):   # __setitem__ (segment_tree.py:84-93)
        i = idx + self.cap
        self.v[i] = val
        i //= 2
        while i >= 1:
            self.v[i] = self.op(self.v[2 * i], self.v[2 * i + 1])
            i //= 2

    def get(self, idx):
        return self.v[self.cap + idx]

    def root(self):
        return self.v[1]

    def find_prefixsum_idx(self, prefixsum):   # segment_tree.py:118-139
        idx = 1
        while idx < self.cap:
            if self.v[2 * idx] > prefixsum:
                idx = 2 * idx
            else:
                prefixsum -= self.v[2 * idx]
                idx = 2 * idx + 1
        return idx - self.cap


class PrioritizedReplayOracle:
    def __init__(self, capacity, alpha, beta):
        cap = 1
        while cap < capacity:
            cap *= 2
        self.sum, self.min = SegTree(cap, lambda a, b: a + b, 0.0), SegTree(cap, min, np.inf)
        self.alpha, self.beta, self.cap = alpha, beta, cap
        self.size, self.next_idx, self.max_priority = 0, 0, 1.0
        self.store = {}

    def add(self, trans, weight=None):   # buffer.py:128-136
        idx = self.next_idx
        self.store[idx] = trans
        self.next_idx = (self.next_idx + 1) % self.cap
        self.size = min(self.size + 1, self.cap)
        w = self.max_priority if weight is None else weight
        self.sum.set(idx, w ** self.alpha)
        self.min.set(idx, w ** self.alpha)

    def sample_idx(self, u):   # buffer.py:138-144 with explicit uniforms
        return np.array([min(self.sum.find_prefixsum_idx(float(x) * self.sum.root()), self.size - 1) for x in u], np.int32)

    def weights(self, idxes):   # buffer.py:149-160
        total = self.sum.root()
        p_min = self.min.root() / total
        max_w = (p_min * self.size) ** (-self.beta)
        return np.array([((self.sum.get(i) / total) * self.size) ** (-self.beta) / max_w for i in idxes])

    def update_priorities(self, idxes, priorities):   # buffer.py:167-189
        for i, p in zip(idxes, priorities):
            assert p > 0
            self.sum.set(int(i), p ** self.alpha)
            self.min.set(int(i), p ** self.alpha)
            self.max_priority = max(self.max_priority, p)

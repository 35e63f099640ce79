"""Replays the exact RNG call sequence of tests/golden/make_golden.py:run_replay so that the transitions,
uniforms and priority updates behind tests/golden/replay.npz can be re-derived from the seed."""
import numpy as np


def replay_script(case):
    rng = np.random.default_rng(case['seed'])
    ns = case['n_sample']

    def add(n):
        rows = []
        for _ in range(n):
            o, a = rng.standard_normal(6).astype(np.float32), rng.standard_normal(2).astype(np.float32)
            rows.append((o, a, np.float32(rng.standard_normal()), o + 1, np.float32(0.0)))
        return rows

    steps = [('add', add(case['n_add0'])), ('draw', rng.random(ns).astype(np.float32), 'a')]
    upd_idx = rng.integers(0, case['n_add0'], case['n_update']).astype(np.int32)
    upd_pr = (np.abs(rng.standard_normal(case['n_update'])) + 1e-3).astype(np.float32)
    steps += [('update', upd_idx, upd_pr), ('draw', rng.random(ns).astype(np.float32), 'b'),
              ('add', add(case['n_add1'])), ('draw', rng.random(ns).astype(np.float32), 'c')]
    return steps

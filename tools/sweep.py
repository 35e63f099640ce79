#!/usr/bin/env python3
"""Throughput of the policy forward+backward rollout (mpg_policy_grad) for the BASELINE.json configs that
are parity-test cases rather than bench lines: env x learner mode x batch sweep, both kernel backends.
Prints one JSON line per case (device time by CUDA events, 3 warm-ups, median of 5).
    python tools/sweep.py > profiles/rNN_sweep.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpg_b200 import synthetic  # noqa: E402
from mpg_b200.config import default_args  # noqa: E402
from mpg_b200.policy import PolicyWithQs  # noqa: E402

N = 25


def run(env_id, mode, B, backend):
    args = default_args('NADP' if mode == 'nadp' else 'MPG-v2', env_id, replay_batch_size=B)
    pol = PolicyWithQs(**vars(args))
    pol.set_weights(synthetic.make_policy_with_qs_weights(1, args.obs_dim, args.act_dim, 256, double_q=(mode != 'nadp')))
    e = pol.engine
    if backend == 'tc':
        if not e.tc_available():
            return None
        e.set_backend(1)
    obs = e.dev(synthetic.make_obs(np.random.default_rng(2), env_id, B))
    lst, w, full = ([0, N], [0.0, 1.0], True) if mode == 'nadp' else ([0, N], [0.42, 0.58], False)
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.policy_grad(obs, lst, w, full_bptt=full, use_philox=True, noise_seed=3, want_returns=False)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    return dict(env=env_id, mode=mode, rows=B, backend=backend, ms=ms, state_steps_per_s=B * N / (ms * 1e-3))


if __name__ == '__main__':
    cases = [('PathTracking-v0', 'nadp', b) for b in (256, 4096, 65536, 262144)]
    cases += [('PathTracking-v0', 'mpg', b) for b in (256, 65536, 262144)]
    for env in ('InvertedPendulumConti-v0', 'InvertedDoublePendulum-v2'):
        cases += [(env, 'nadp', b) for b in (1024, 16384, 131072, 1048576)]
    for env, mode, B in cases:
        for backend in ('ffma', 'tc'):
            if backend == 'tc' and mode == 'nadp' and B > 262144:
                continue  # full-BPTT dW2 operand store is 53 KB per row: call in <= 256K-row chunks
            if backend == 'ffma' and B > 262144:
                continue
            r = run(env, mode, B, backend)
            if r:
                print(json.dumps(r), flush=True)
            torch.cuda.empty_cache()

#!/usr/bin/env python3
"""Device time of the rows either side of the hot path (SURVEY.md 8(f) "next" #1-#4): prioritized replay, the real
PathTracking environment, the optimiser step and the MPG-v1 n-step target.  CUDA events, 3 warm-ups, median of 7.
    python tools/bench_next_rows.py > profiles/rNN_next_rows.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpg_b200 import synthetic  # noqa: E402
from mpg_b200.buffer import PrioritizedReplayBuffer  # noqa: E402
from mpg_b200.config import default_args  # noqa: E402
from mpg_b200.envs_and_models import PathTrackingEnv  # noqa: E402
from mpg_b200.learners import MPGLearner  # noqa: E402
from mpg_b200.policy import PolicyWithQs  # noqa: E402


def timed(fn, n=10, warm=3):
    ts = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(0)
    # ---- #1 prioritized replay: capacity 2^20, proportional sampling through the fp64 sum tree --------------------
    cap, nb = 1 << 20, 65536
    args = default_args('MPG-v2', 'PathTracking-v0', max_buffer_size=cap, replay_starts=1, replay_batch_size=256,
                        buffer_type='priority', buffer_log_interval=10 ** 9)
    rb = PrioritizedReplayBuffer(args, 0)
    cols = [torch.randn(nb, 6, device=dev, generator=g), torch.randn(nb, 2, device=dev, generator=g),
            torch.randn(nb, device=dev, generator=g), torch.randn(nb, 6, device=dev, generator=g), torch.zeros(nb, device=dev)]
    for _ in range(cap // nb):
        rb.add_arrays(*cols)
    ms = timed(lambda: rb.add_arrays(*cols))
    emit(row='replay', op='add 65536 transitions (ring write + tree rebuild)', ms=ms, transitions_per_s=nb / ms * 1e3)
    for n in (256, 65536):
        ms = timed(lambda: rb.replay_device(n))
        emit(row='replay', op=f'sample {n} (prefix-sum walk, gather, IS weights)', ms=ms, samples_per_s=n / ms * 1e3)
    smp = rb.replay_device(65536)
    pr = torch.rand(65536, device=dev, generator=g) + 0.1
    ms = timed(lambda: rb.update_priorities(smp[-1], pr))
    emit(row='replay', op='update 65536 priorities (last duplicate wins + tree rebuild)', ms=ms, updates_per_s=65536 / ms * 1e3)
    # ---- #3 real environment -----------------------------------------------------------------------------------------
    for agents in (8, 65536):
        env = PathTrackingEnv(num_agent=agents)
        env.reset()
        act = torch.rand(agents, 2, device=dev, generator=g) * 2 - 1
        ms = timed(lambda: env.step(act))
        emit(row='real env', op=f'step, {agents} agents (200 Hz x 20 sub-steps, path projection, judge_done)', ms=ms,
             agent_steps_per_s=agents / ms * 1e3)
    B = 65536
    a1 = default_args('MPG-v1', 'PathTracking-v0', replay_batch_size=B, sample_num_in_learner=25)
    L = MPGLearner(PolicyWithQs, a1)
    L.set_weights(synthetic.make_policy_with_qs_weights(5, a1.obs_dim, a1.act_dim, 256, double_q=False))
    rng = np.random.default_rng(1)
    obs = synthetic.make_obs(rng, 'PathTracking-v0', B)
    batch = [obs, rng.uniform(-1, 1, (B, 2)).astype(np.float32), rng.normal(size=B).astype(np.float32),
             synthetic.make_obs(rng, 'PathTracking-v0', B), np.zeros(B, np.float32)]
    L._upload_batch(batch)
    ms = timed(lambda: L.compute_n_step_target())
    emit(row='real env', op='MPG-v1 n-step target: 25 real-env steps x 65536 rows + Q bootstrap', ms=ms,
         env_steps_per_s=B * 25 / ms * 1e3, backend=int(L.engine.backend))
    # ---- #2 optimiser step ---------------------------------------------------------------------------------------------
    pol = PolicyWithQs(**vars(default_args('MPG-v2', 'PathTracking-v0')))
    n = sum(pol.engine.param_count(s) for s in pol.model_slots)
    flat = torch.randn(n, device=dev, generator=g) * 1e-3
    it = [0]

    def apply():
        pol.apply_gradients(it[0], flat)
        it[0] += 2
    ms = timed(apply)
    emit(row='optimiser', op='apply_gradients MPG-v2: 3 x Keras-Adam + 3 x Polyak + weight re-pack (205,318 parameters)', ms=ms)


if __name__ == '__main__':
    main()

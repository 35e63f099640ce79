"""Prioritized replay: the numpy oracle and the GPU buffer against outputs of the
reference's own PrioritizedReplayBuffer / segment-tree classes (tests/golden/replay.npz)."""
import numpy as np
import pytest

from tests.replay_util import replay_script
from tests.util import load_golden


def test_replay_oracle_matches_reference():
    from oracle.replay_oracle import PrioritizedReplayOracle
    case, gold = load_golden('replay')
    rb = PrioritizedReplayOracle(case['capacity'], 0.6, 0.4)
    for step in replay_script(case):
        if step[0] == 'add':
            for tr in step[1]:
                rb.add(tr, None)
        elif step[0] == 'update':
            rb.update_priorities(step[1], [float(p) for p in step[2]])
        else:
            u, tag = step[1], step[2]
            assert np.array_equal(u, gold[f'u_{tag}__f32'])
            idx = rb.sample_idx(u)
            assert np.array_equal(idx, gold[f'idx_{tag}__f32'])
            assert np.allclose(rb.weights(idx), gold[f'w_{tag}__f64'], rtol=1e-12)
            assert np.isclose(rb.sum.root(), gold[f'sum_{tag}__f64'], rtol=1e-13)
            assert np.isclose(rb.min.root(), gold[f'min_{tag}__f64'], rtol=1e-13)
            assert np.array_equal(np.stack([rb.store[int(i)][0] for i in idx]), gold[f'obs_{tag}__f32'])
    assert np.isclose(rb.max_priority, gold['max_priority__f64'])


@pytest.mark.gpu
def test_gpu_replay_matches_reference():
    from argparse import Namespace
    from mpg_b200.buffer import PrioritizedReplayBuffer
    case, gold = load_golden('replay')
    args = Namespace(max_buffer_size=case['capacity'], replay_starts=1, replay_batch_size=case['n_sample'], replay_alpha=0.6,
                     replay_beta=0.4, buffer_log_interval=10 ** 9, obs_dim=6, act_dim=2)
    rb = PrioritizedReplayBuffer(args, 0)
    for step in replay_script(case):
        if step[0] == 'add':
            rb.add_batch(step[1])
        elif step[0] == 'update':
            rb.dev.update_priorities(step[1], step[2])   # the golden priorities are already positive
        else:
            u, tag = step[1], step[2]
            obs, act, rew, obs1, done, w, idx = rb.dev.sample(len(u), u)
            assert np.array_equal(idx.cpu().numpy(), gold[f'idx_{tag}__f32'])        # bit-exact indices
            assert np.array_equal(obs.cpu().numpy(), gold[f'obs_{tag}__f32'])        # bit-exact gathers
            assert np.array_equal(rew.cpu().numpy(), gold[f'rew_{tag}__f32'])
            assert np.allclose(w.cpu().numpy(), gold[f'w_{tag}__f64'], rtol=2e-6)     # fp64 math, fp32 output
            s, m, _ = rb.dev.tree_stats()
            assert np.isclose(s, gold[f'sum_{tag}__f64'], rtol=1e-12) and np.isclose(m, gold[f'min_{tag}__f64'], rtol=1e-12)
    assert np.isclose(rb.dev.tree_stats()[2], gold['max_priority__f64'], rtol=1e-7)
    assert len(rb) == case['capacity']
    out = rb.replay()
    assert len(out) == 7 and out[0].shape == (case['n_sample'], 6) and out[6].dtype == np.int32


@pytest.mark.gpu
def test_gpu_replay_full_size_properties():
    """Capacity 2^19, 256K draws (BASELINE config 5): indices in range, empirical frequencies follow p_i^alpha,
    uniform buffer draws are uniform, priorities update, weights are in (0, 1]."""
    import torch
    from argparse import Namespace
    from mpg_b200.buffer import PrioritizedReplayBuffer
    cap, n = 1 << 19, 1 << 18
    args = Namespace(max_buffer_size=cap, replay_starts=1, replay_batch_size=n, replay_alpha=0.6, replay_beta=0.4,
                     buffer_log_interval=10 ** 9, obs_dim=6, act_dim=2)
    rb = PrioritizedReplayBuffer(args, 0)
    g = torch.Generator(device='cuda').manual_seed(0)
    N = 400000
    obs = torch.randn(N, 6, device='cuda', generator=g)
    rb.add_arrays(obs, torch.zeros(N, 2, device='cuda'), torch.arange(N, device='cuda', dtype=torch.float32), obs + 1,
                  torch.zeros(N, device='cuda'))
    o, a, r, o1, d, w, idx = rb.replay_device()
    assert int(idx.min()) >= 0 and int(idx.max()) < N and torch.equal(r, idx.float()) and torch.allclose(w, torch.ones_like(w))
    # raise the priority of the first 1000 transitions 100x: they must take ~ 1000*100^0.6 / (1000*100^0.6 + N-1000) of the draws
    hot = torch.arange(1000, device='cuda', dtype=torch.int32)
    rb.update_priorities(hot, torch.full((1000,), 100.0, device='cuda'))
    o, a, r, o1, d, w, idx = rb.replay_device()
    frac = float((idx < 1000).float().mean())
    want = 1000 * 100 ** 0.6 / (1000 * 100 ** 0.6 + (N - 1000))
    assert abs(frac - want) < 0.1 * want, (frac, want)
    assert float(w.max()) <= 1.0 + 1e-6 and float(w.min()) > 0
    assert torch.allclose(w[idx < 1000], torch.full_like(w[idx < 1000], (100 ** 0.6) ** -0.4), rtol=1e-4)


@pytest.mark.gpu
def test_config5_prioritized_replay_feeds_mpg_learner():
    """BASELINE config 5 wiring: GPU sum-tree sampling -> MPGLearner.compute_gradient on device tensors ->
    TD errors back into the tree.  The gradient of the sampled batch must equal the gradient of the same rows
    passed as host numpy (the reference's path)."""
    import torch
    from argparse import Namespace
    from mpg_b200 import synthetic
    from mpg_b200.buffer import PrioritizedReplayBuffer
    from mpg_b200.config import default_args
    from mpg_b200.learners import MPGLearner
    from mpg_b200.policy import PolicyWithQs
    from tests.util import make_batch
    B, N = 2048, 50000
    args = default_args('MPG-v2', 'PathTracking-v0', replay_batch_size=B, buffer_type='priority', max_buffer_size=65536,
                        replay_starts=1, buffer_log_interval=10 ** 9)
    rb = PrioritizedReplayBuffer(args, 0)
    data = make_batch(3, args.env_id, N, 0)
    rb.add_arrays(*data)
    learner = MPGLearner(PolicyWithQs, args)
    learner.set_weights(synthetic.make_policy_with_qs_weights(4, args.obs_dim, args.act_dim, 256, double_q=True))
    samples = rb.replay_device()
    g_dev = learner.compute_gradient(samples[:5], rb, samples[-1], 100)
    info = learner.get_info_for_buffer()
    assert isinstance(info['td_error'], torch.Tensor) and info['td_error'].shape == (B,)
    rb.update_priorities(info['indexes'], info['td_error'])
    s, m, mx = rb.dev.tree_stats()
    assert np.isfinite(s) and m > 0 and mx >= float(info['td_error'].abs().max()) - 1e-6
    host = [t.cpu().numpy() for t in samples[:5]]
    learner2 = MPGLearner(PolicyWithQs, args)
    learner2.set_weights(learner.get_weights())
    g_host = learner2.compute_gradient(host, None, samples[-1].cpu().numpy(), 100)
    for a, b in zip(g_dev, g_host):
        assert np.array_equal(a, b)
    idx2 = rb.replay_device()[-1]
    assert int(idx2.max()) < N

"""worker / optimizer / evaluator / trainer glue around the hot path (reference worker.py,
optimizer.py:286-397, evaluator.py, trainer.py).  A short training run in the real PathTracking environment must
improve the deterministic evaluation return -- the learning-curve property the reference's README figures show."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _args(alg, **kw):
    from mpg_b200.config import default_args
    base = dict(replay_batch_size=256, batch_size=512, num_agent=8, explore_sigma=0.1, max_buffer_size=100000,
                replay_starts=2048, buffer_log_interval=10 ** 9, num_eval_agent=64, num_eval_episode=1, fixed_steps=60,
                eval_interval=10 ** 9, log_interval=10 ** 9, max_iter=400, log_dir=None,
                policy_lr_schedule=[3e-4, 100000, 3e-5], value_lr_schedule=[8e-4, 100000, 8e-5])
    base.update(kw)
    return default_args(alg, 'PathTracking-v0', **base)


def test_worker_sample_format_and_reset_of_done_agents():
    import torch
    from mpg_b200.policy import PolicyWithQs
    from mpg_b200.worker import OffPolicyWorker
    args = _args('MPG-v2')
    w = OffPolicyWorker(PolicyWithQs, args.env_id, args, 0)
    batch, n = w.sample_with_count()
    assert n == args.batch_size and len(batch[0]) == 5
    assert batch[0][0].shape == (args.obs_dim,) and batch[0][1].shape == (args.act_dim,)
    # reset_done: only flagged agents are re-drawn
    env = w.env
    before = env.obs.clone()
    env.done = torch.zeros(args.num_agent, device=before.device)
    env.done[3] = 1
    after = env.reset_done()
    keep = [i for i in range(args.num_agent) if i != 3]
    assert torch.equal(after[keep], before[keep]) and not torch.equal(after[3], before[3])
    assert w.get_stats()['num_sample'] == args.batch_size


@pytest.mark.parametrize('alg,buffer_type', [('MPG-v2', 'normal'), ('NADP', 'normal'), ('MPG-v1', 'priority')])
def test_short_training_improves_evaluation_return(alg, buffer_type):
    from mpg_b200.trainer import Trainer
    args = _args(alg, buffer_type=buffer_type)
    tr = Trainer(args)
    before = tr.evaluator.run_evaluation(0)
    tr.train(args.max_iter)
    after = tr.evaluator.run_evaluation(args.max_iter)
    st = tr.learner.get_stats()
    assert np.isfinite(st['policy_gradient_norm'])
    assert tr.optimizer.get_stats()['num_sampled_steps'] >= args.replay_starts
    assert after['episode_return'] > before['episode_return'], (before, after)


def test_checkpoint_round_trip_resumes_bit_exactly(tmp_path):
    """save_weights / load_weights (policy.py:98-110): weights, targets, Adam moments and step counters; a restored
    policy must continue exactly like the original."""
    from mpg_b200.policy import PolicyWithQs
    args = _args('MPG-v2')
    rng = np.random.default_rng(0)
    a, b = PolicyWithQs(**vars(args)), PolicyWithQs(**dict(vars(args), seed=5))
    n = sum(a.engine.param_count(s) for s in a.model_slots)
    for it in range(3):
        a.apply_gradients(it, [rng.standard_normal(n).astype(np.float32)])
    a.save_weights(str(tmp_path), 3)
    b.load_weights(str(tmp_path), 3)
    g = [rng.standard_normal(n).astype(np.float32)]
    a.apply_gradients(4, g)
    b.apply_gradients(4, g)
    for wa, wb in zip(a.get_weights(), b.get_weights()):
        for x, y in zip(wa, wb):
            assert np.array_equal(x, y)


@pytest.mark.parametrize('sigma', [0.3, 0.0])
def test_fused_sampler_matches_step_by_step(sigma):
    """mpg_env_sample (one launch for the whole sample() loop) against the step-by-step path through
    mpg_policy_forward + mpg_env_step + reset of done agents, same pre-drawn exploration noise and reset states.
    The two paths run the same device functions in different kernels (FMA contraction may differ by an ulp), so
    values are compared to 1e-5 relative and the done flags / restart pattern exactly."""
    from mpg_b200.policy import PolicyWithQs
    from mpg_b200.worker import OffPolicyWorker
    args = _args('MPG-v2', num_agent=70, batch_size=70 * 40, explore_sigma=sigma)
    outs = []
    for fused in (True, False):
        w = OffPolicyWorker(PolicyWithQs, args.env_id, args, 0)
        w.policy_with_value.engine.set_backend(0)      # like for like: the fused kernel runs the fp32 tile MLP
        outs.append([t.cpu().numpy() for t in w.sample_arrays(fused=fused)] + [w.env.state.cpu().numpy(), w.obs.cpu().numpy()])
    assert outs[0][4].sum() > 0, 'the case must contain finished agents (reset path)'
    assert np.array_equal(outs[0][4], outs[1][4])
    for a, b in zip(*outs):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1.0)


def test_capacity_growth_keeps_weights_and_optimizer_state():
    """Engine.ensure_capacity re-creates the handle for a bigger batch; weights AND Adam moments must move over."""
    from mpg_b200.policy import PolicyWithQs
    args = _args('NADP', replay_batch_size=64)
    rng = np.random.default_rng(1)
    a, b = PolicyWithQs(**vars(args)), PolicyWithQs(**vars(args))
    n = sum(a.engine.param_count(s) for s in a.model_slots)
    g0, g1 = [rng.standard_normal(n).astype(np.float32) for _ in range(2)]
    a.apply_gradients(0, [g0]); b.apply_gradients(0, [g0])
    a.engine.ensure_capacity(5000, 30)                # grows: new handle
    a.apply_gradients(1, [g1]); b.apply_gradients(1, [g1])
    for wa, wb in zip(a.get_weights(), b.get_weights()):
        for x, y in zip(wa, wb):
            assert np.array_equal(x, y)


def test_fused_evaluation_matches_step_by_step():
    from mpg_b200.evaluator import Evaluator
    from mpg_b200.policy import PolicyWithQs
    args = _args('MPG-v2', num_eval_agent=70, fixed_steps=30)
    ev = Evaluator(PolicyWithQs, args.env_id, args)
    ev.policy_with_value.engine.set_backend(0)         # like for like: the fused kernel runs the fp32 tile MLP
    a, b = ev.run_n_episodes(fused=True), ev.run_n_episodes(fused=False)
    for k in a:
        assert abs(a[k] - b[k]) <= 1e-4 * max(abs(b[k]), 1.0), (k, a[k], b[k])

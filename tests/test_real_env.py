"""Real PathTracking environment + MPG-v1 n-step targets against outputs of the
reference's own PathTrackingEnv / MPGLearner.compute_n_step_target (tests/golden/real_env_*.npz)."""
import numpy as np
import pytest
import torch

from mpg_b200 import synthetic
from mpg_b200.config import default_args
from tests.util import load_golden, make_batch, rel_l2

PT = 'PathTracking-v0'


def _case_inputs(case):
    B, nfd = case['B'], case['nfd']
    rng = np.random.default_rng(case['bseed'])
    obs0 = synthetic.make_obs(rng, PT, B, nfd)
    acts = rng.uniform(-1.2, 1.2, (case['n_env'], B, 2)).astype(np.float32)
    args = default_args('MPG-v1', PT, replay_batch_size=B, num_future_data=nfd, value_num_hidden_units=case['H'],
                        policy_num_hidden_units=case['H'], sample_num_in_learner=case['T'])
    w = synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, case['H'], double_q=False)
    batch = make_batch(case['bseed'] + 1, PT, B, nfd)
    return obs0, acts, args, w, batch


@pytest.mark.parametrize('name', ['real_env_h64', 'real_env_h256'])
def test_real_env_oracle_matches_reference(name):
    from oracle import mpg_oracle as O
    case, gold = load_golden(name)
    obs0, acts, args, w, batch = _case_inputs(case)
    env = O.PathTrackingEnvOracle(case['nfd'], np.float32)
    env.reset(obs0)
    for t in range(case['n_env']):
        o, r, d = env.step(acts[t])
        assert rel_l2(o, gold['env_obs__f32'][t]) <= 2e-5, t
        assert rel_l2(r, gold['env_rew__f32'][t]) <= 2e-5, t
        assert np.array_equal(d.astype(np.int32), gold['env_done__f32'][t]), t
    tgt32, _ = O.mpg_v1_n_step_target(args, w, batch, case['T'], torch.float32)
    tgt64, _ = O.mpg_v1_n_step_target(args, w, batch, case['T'], torch.float64)
    assert rel_l2(tgt32, gold['batch_targets__f32']) <= 1e-4
    assert rel_l2(tgt64, gold['batch_targets__f32']) <= 1e-4


@pytest.mark.gpu
def test_gpu_real_env_matches_reference():
    from mpg_b200.envs_and_models import PathTrackingEnv
    for name in ('real_env_h64', 'real_env_h256'):
        case, gold = load_golden(name)
        obs0, acts, _, _, _ = _case_inputs(case)
        env = PathTrackingEnv(num_future_data=case['nfd'], num_agent=case['B'])
        env.reset(init_obs=obs0)
        for t in range(case['n_env']):
            o, r, d, _ = env.step(acts[t])
            assert rel_l2(o.cpu().numpy(), gold['env_obs__f32'][t]) <= 1e-5, (name, t)
            assert rel_l2(r.cpu().numpy(), gold['env_rew__f32'][t]) <= 1e-5, (name, t)
            assert np.array_equal(d.cpu().numpy(), gold['env_done__f32'][t]), (name, t)


@pytest.mark.gpu
@pytest.mark.parametrize('backend', ['ffma', 'tc'])
def test_gpu_mpg_v1_n_step_target(backend):
    """compute_n_step_target with 25 real-env steps: vs the reference golden (H=256) and vs the fp64 oracle at B=300."""
    from oracle import mpg_oracle as O
    from mpg_b200.learners import MPGLearner
    from mpg_b200.policy import PolicyWithQs
    case, gold = load_golden('real_env_h256')
    _, _, args, w, batch = _case_inputs(case)
    learner = MPGLearner(PolicyWithQs, args)
    learner.set_weights(w)
    learner.engine.set_backend(1 if backend == 'tc' else 0)
    learner.get_batch_data(batch, None, None)
    got = learner.batch_data['batch_targets'].cpu().numpy()
    assert rel_l2(got, gold['batch_targets__f32']) <= 2e-5
    B = 300
    args2 = default_args('MPG-v1', PT, replay_batch_size=B, sample_num_in_learner=25)
    w2 = synthetic.make_policy_with_qs_weights(5, args2.obs_dim, args2.act_dim, 256, double_q=False)
    batch2 = make_batch(6, PT, B, 0)
    l2 = MPGLearner(PolicyWithQs, args2)
    l2.set_weights(w2)
    l2.engine.set_backend(1 if backend == 'tc' else 0)
    grads = l2.compute_gradient(batch2, None, None, 3000)
    ref, _ = O.mpg_v1_n_step_target(args2, w2, batch2, 25, torch.float64)
    assert rel_l2(l2.batch_data['batch_targets'].cpu().numpy(), ref) <= 2e-5
    assert len(grads) == 12 and all(np.isfinite(g).all() for g in grads)

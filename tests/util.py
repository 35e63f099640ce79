"""Shared helpers for the test-suite: golden loading, input re-derivation, parity metrics."""
import json
import os

import numpy as np

from mpg_b200 import synthetic
from mpg_b200.config import default_args

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    case = json.loads(str(d['case_json']))
    return case, d


def rel_l2(x, ref):
    x, ref = np.asarray(x, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / (den if den > 0 else 1.0))


def make_batch(seed, env_id, B, nfd):
    """Same law as tests/golden/make_golden.py:make_batch (inputs are re-derived from seeds)."""
    rng = np.random.default_rng(seed)
    obs = synthetic.make_obs(rng, env_id, B, nfd)
    act_dim = synthetic.ENV_DIMS[env_id][1]
    act = rng.uniform(-1, 1, (B, act_dim)).astype(np.float32)
    rew = (-np.abs(rng.standard_normal(B))).astype(np.float32)
    obs_tp1 = synthetic.make_obs(rng, env_id, B, nfd)
    done = np.zeros(B, np.float32)
    return [obs, act, rew, obs_tp1, done]


def nadp_case_inputs(case):
    env_id, B, H, n, M, nfd = case['env_id'], case['B'], case['H'], case['n'], case['M'], case['nfd']
    args = default_args('NADP', env_id, replay_batch_size=B, M=M, num_future_data=nfd,
                        value_num_hidden_units=H, policy_num_hidden_units=H,
                        num_rollout_list_for_policy_update=[n], num_rollout_list_for_q_estimation=[n],
                        buffer_type=case.get('buffer_type', 'normal'))
    w = synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=False)
    batch = make_batch(case['bseed'], env_id, B, nfd)
    rng = np.random.default_rng(case['nseed'])
    noise_q = synthetic.make_noise(rng, n, B * M)
    noise_p = synthetic.make_noise(rng, n, B * M)
    return args, w, batch, noise_q, noise_p


def mpg_case_inputs(case):
    env_id, B, H, M, nfd = case['env_id'], case['B'], case['H'], case['M'], case['nfd']
    ver = case.get('version', 'MPG-v2')
    args = default_args(ver, env_id, replay_batch_size=B, M=M, num_future_data=nfd,
                        value_num_hidden_units=H, policy_num_hidden_units=H,
                        num_rollout_list_for_policy_update=case['rollout_list'],
                        deriv_interval_policy=case.get('deriv_interval_policy', False),
                        buffer_type=case.get('buffer_type', 'normal'), sample_num_in_learner=None)
    w = synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H,
                                              double_q=(ver == 'MPG-v2'))
    batch = make_batch(case['bseed'], env_id, B, nfd)
    n = max(case['rollout_list'])
    noise_p = synthetic.make_noise(np.random.default_rng(case['nseed']), n, B * M)
    return args, w, batch, noise_p


def model_case_inputs(case):
    env_id, B, H, n, nfd = case['env_id'], case['B'], case['H'], case['n'], case['nfd']
    args = default_args('NADP', env_id, num_future_data=nfd, value_num_hidden_units=H, policy_num_hidden_units=H)
    rng = np.random.default_rng(case['bseed'])
    obs0 = synthetic.make_obs(rng, env_id, B, nfd)
    acts = rng.uniform(-1, 1, (n, B, args.act_dim)).astype(np.float32)
    noise = synthetic.make_noise(np.random.default_rng(case['nseed']), n, B)
    w = synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=False)
    return args, obs0, acts, noise, w

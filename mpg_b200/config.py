"""Default learner arguments (an argparse-style namespace) for the hot path.

The reference passes one `args` namespace everywhere and splats it into constructors
(`learners/mpg_learner.py:30-58`, `learners/nadp.py:29-47`). Only the fields those constructors
and the rollout read are reproduced here; values are the reference defaults:
  PathTracking  train_scripts/train_script.py:177-306 (MPG) / 308-429 (NADP) / 57-175 (AMPC)
  pendulums     train_scripts/train_script4mujoco.py:169-294, 296-412
"""
from argparse import Namespace

from .synthetic import ENV_DIMS, default_obs_scale


def default_args(alg='NADP', env_id='PathTracking-v0', **overrides):
    """alg in {'MPG-v1','MPG-v2','NADP','AMPC'}."""
    nfd = overrides.get('num_future_data', 0)
    obs_dim, act_dim, _ = ENV_DIMS[env_id]
    if env_id == 'PathTracking-v0':
        obs_dim += nfd
    pendulum = env_id != 'PathTracking-v0'
    a = dict(
        env_id=env_id, num_future_data=nfd, num_agent=8,
        alg_name=alg.split('-')[0], learner_version=alg if alg.startswith('MPG') else None,
        sample_num_in_learner=25, M=1, deriv_interval_policy=False,
        num_rollout_list_for_policy_update=[0, 25] if alg.startswith('MPG') else [25],
        num_rollout_list_for_q_estimation=[25] if alg == 'NADP' else [],
        eta=0.1, rule_based_bias_total_ite=4000 if pendulum else 9000,
        gamma=1.0 if alg == 'AMPC' else 0.98, gradient_clip_norm=3.0,
        num_batch_reuse=10 if alg == 'MPG-v1' else 1,
        buffer_type='normal', replay_batch_size=256, replay_alpha=0.6, replay_beta=0.4,
        obs_dim=obs_dim, act_dim=act_dim,
        value_model_cls='MLP', value_num_hidden_layers=2, value_num_hidden_units=256,
        value_hidden_activation='elu', value_lr_schedule=[8e-5, 100000, 8e-6],
        policy_model_cls='MLP', policy_num_hidden_layers=2, policy_num_hidden_units=256,
        policy_hidden_activation='elu',
        policy_out_activation='linear' if pendulum else 'tanh',
        policy_lr_schedule=[3e-5, 100000, 3e-6],
        alpha=None, alpha_lr_schedule=None,
        policy_only=(alg == 'AMPC'), double_Q=(alg == 'MPG-v2'), target=True, tau=0.005,
        delay_update=2 if alg.startswith('MPG') else 1,
        deterministic_policy=True, action_range=3.0 if pendulum else None,
        obs_ptype='scale', obs_scale=default_obs_scale(env_id, nfd),
        rew_ptype='scale', rew_scale=1.0 if pendulum else 0.01, rew_shift=0.0,
    )
    a.update(overrides)
    return Namespace(**a)

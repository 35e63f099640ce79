#!/usr/bin/env python3
"""Generate golden vectors by executing the reference's own Python (from /root/reference).

Run in the BUILD CONTAINER only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
It installs oracle/tf_shim.py (torch-backed stand-ins for tensorflow / tfp / gym / matplotlib,
which are not installable here), imports the UNMODIFIED reference modules
    envs_and_models.*, preprocessor, model, policy, learners.nadp, learners.mpg_learner
and calls their public methods on seeded inputs (mpg_b200/synthetic.py).  Each case is run in
float32 (reference precision) and float64 ("truth") and written to tests/golden/<case>.npz.
Inputs are NOT stored (they are re-derived from the seeds recorded in the file), only outputs.

Also extracts the reference's single recorded data file mpc/mpc_rl.npy (100 real-env transitions)
into tests/golden/mpc_rl_transitions.npz, the known-answer check for f_xu.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import tf_shim  # noqa: E402

tf_shim.install()
sys.path.insert(0, '/root/reference')

from mpg_b200 import synthetic  # noqa: E402
from mpg_b200.config import default_args  # noqa: E402

torch.set_num_threads(1)


class NoiseQueue:
    def __init__(self):
        self.rows = []

    def load(self, *arrays):
        self.rows = [r for a in arrays for r in a]

    def __call__(self, shape):
        r = self.rows.pop(0)
        assert tuple(r.shape) == tuple(shape), (r.shape, shape)
        return torch.as_tensor(r)


NOISE = NoiseQueue()
tf_shim.set_noise_source(NOISE)


def make_batch(seed, env_id, B, nfd):
    rng = np.random.default_rng(seed)
    obs = synthetic.make_obs(rng, env_id, B, nfd)
    act_dim = synthetic.ENV_DIMS[env_id][1]
    act = rng.uniform(-1, 1, (B, act_dim)).astype(np.float32)
    rew = (-np.abs(rng.standard_normal(B))).astype(np.float32)
    obs_tp1 = synthetic.make_obs(rng, env_id, B, nfd)
    done = np.zeros(B, np.float32)
    return [obs, act, rew, obs_tp1, done]


def np_list(ts):
    return [np.asarray(t.numpy() if hasattr(t, 'numpy') else t) for t in ts]


def flat(ts):
    return np.concatenate([np.asarray(t, dtype=np.float64).ravel() for t in np_list(ts)])


def run_nadp(case):
    from learners.nadp import NADPLearner
    from policy import PolicyWithQs
    env_id, B, H, n, M, nfd = case['env_id'], case['B'], case['H'], case['n'], case['M'], case['nfd']
    args = default_args('NADP', env_id, replay_batch_size=B, M=M, num_future_data=nfd,
                        value_num_hidden_units=H, policy_num_hidden_units=H,
                        num_rollout_list_for_policy_update=[n], num_rollout_list_for_q_estimation=[n],
                        buffer_type=case.get('buffer_type', 'normal'))
    learner = NADPLearner(PolicyWithQs, args)
    w = synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=False)
    learner.set_weights(w)
    batch = make_batch(case['bseed'], env_id, B, nfd)
    rng = np.random.default_rng(case['nseed'])
    noise_q = synthetic.make_noise(rng, n, B * M)
    noise_p = synthetic.make_noise(rng, n, B * M)
    out = {}
    learner.get_batch_data(batch, None, None)
    obs, act = learner.batch_data['batch_obs'], learner.batch_data['batch_actions']
    NOISE.load(noise_q)
    out['q_targets'] = learner.model_rollout_for_q_estimation(obs, act).numpy()
    NOISE.load(noise_q)
    q_loss, q_grad = learner.q_forward_and_backward(obs, act)
    out['q_loss'], out['q_grad'] = q_loss.numpy(), flat(q_grad)
    NOISE.load(noise_p)
    p_loss, p_grad, vmean = learner.policy_forward_and_backward(obs)
    out['policy_loss'], out['policy_grad'], out['value_mean'] = p_loss.numpy(), flat(p_grad), vmean.numpy()
    NOISE.load(noise_q, noise_p)
    grads = learner.compute_gradient(batch, None, None, 7)
    out['compute_gradient'] = flat(grads)
    st = learner.get_stats()
    for k in ('q_loss', 'policy_loss', 'value_mean', 'q_gradient_norm', 'policy_gradient_norm'):
        out['stat_' + k] = np.asarray(st[k])
    if args.buffer_type != 'normal':
        out['td_error'] = np.asarray(learner.get_info_for_buffer()['td_error'])
    return out


def run_mpg(case):
    from learners.mpg_learner import MPGLearner
    from policy import PolicyWithQs
    env_id, B, H, M, nfd = case['env_id'], case['B'], case['H'], case['M'], case['nfd']
    ver = case.get('version', 'MPG-v2')
    args = default_args(ver, env_id, replay_batch_size=B, M=M, num_future_data=nfd,
                        value_num_hidden_units=H, policy_num_hidden_units=H,
                        num_rollout_list_for_policy_update=case['rollout_list'],
                        deriv_interval_policy=case.get('deriv_interval_policy', False),
                        buffer_type=case.get('buffer_type', 'normal'),
                        sample_num_in_learner=None)  # MPG-v1: 1-step TD target branch (mpg_learner.py:147-152)
    learner = MPGLearner(PolicyWithQs, args)
    dq = ver == 'MPG-v2'
    w = synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=dq)
    learner.set_weights(w)
    batch = make_batch(case['bseed'], env_id, B, nfd)
    n = max(case['rollout_list'])
    rng = np.random.default_rng(case['nseed'])
    noise_p = synthetic.make_noise(rng, n, B * M)
    ite = case['iteration']
    out = {}
    learner.get_batch_data(batch, None, None)
    bd = learner.batch_data
    out['batch_targets'] = np.asarray(bd['batch_targets'])
    if args.buffer_type != 'normal':
        out['td_error'] = np.asarray(learner.get_info_for_buffer()['td_error'])
    res = learner.q_forward_and_backward(bd['batch_obs'], bd['batch_actions'], bd['batch_targets'])
    if dq:
        out['q_loss1'], out['q_loss2'] = res[0].numpy(), res[1].numpy()
        out['q_grad1'], out['q_grad2'] = flat(res[2]), flat(res[3])
    else:
        out['q_loss1'], out['q_grad1'] = res[0].numpy(), flat(res[1])
    learner.policy_for_rollout.set_weights(learner.policy_with_value.get_weights())  # mpg_learner.py:422
    NOISE.load(noise_p)
    var, minus_ret, vmean = learner.model_rollout_for_policy_update(bd['batch_obs'])
    out['returns_var'], out['minus_returns'], out['value_mean'] = var.numpy(), minus_ret.numpy(), vmean.numpy()
    NOISE.load(noise_p)
    pg, total_loss, vmean2, ws, ws_new, all_losses = learner.policy_forward_and_backward(
        bd['batch_obs'], learner.tf.convert_to_tensor(ite, dtype=learner.tf.float32), None, learner.ws_old)
    out['policy_grad'], out['total_loss'], out['ws'] = flat(pg), total_loss.numpy(), ws.numpy()
    NOISE.load(noise_p)
    grads = learner.compute_gradient(batch, None, None, ite)
    out['compute_gradient'] = flat(grads)
    st = learner.get_stats()
    keys = ['value_mean', 'policy_total_loss', 'policy_gradient_norm', 'q_loss1', 'q_gradient_norm1']
    keys += ['q_loss2', 'q_gradient_norm2'] if dq else []
    for k in keys:
        out['stat_' + k] = np.asarray(st[k])
    out['stat_w_list'] = np.asarray(st['w_list'])
    out['stat_all_losses'] = np.asarray(st['all_losses'])
    return out


def run_model(case):
    """Open-loop (given actions) and closed-loop (reference policy + preprocessor) trajectories."""
    from envs_and_models import NAME2MODELCLS
    from policy import PolicyWithQs
    from preprocessor import Preprocessor
    env_id, B, H, n, nfd = case['env_id'], case['B'], case['H'], case['n'], case['nfd']
    args = default_args('NADP', env_id, num_future_data=nfd, value_num_hidden_units=H, policy_num_hidden_units=H)
    rng = np.random.default_rng(case['bseed'])
    obs0 = synthetic.make_obs(rng, env_id, B, nfd)
    acts = rng.uniform(-1, 1, (n, B, args.act_dim)).astype(np.float32)
    noise = synthetic.make_noise(np.random.default_rng(case['nseed']), n, B)
    tf = sys.modules['tensorflow']
    out = {}
    model = NAME2MODELCLS[env_id](**vars(args))
    NOISE.load(noise)
    model.reset(tf.convert_to_tensor(obs0))
    o_list, r_list = [], []
    for t in range(n):
        o, r = model.rollout_out(tf.convert_to_tensor(acts[t]))
        o_list.append(o.numpy()); r_list.append(r.numpy())
    out['open_obs'], out['open_rew'] = np.stack(o_list), np.stack(r_list)
    # closed loop, exactly the loop of learners/nadp.py:141-152
    pol = PolicyWithQs(**vars(args))
    pol.set_weights(synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=False))
    ppc = Preprocessor(args.obs_dim, args.obs_ptype, args.rew_ptype, args.obs_scale, args.rew_scale,
                       args.rew_shift, gamma=args.gamma)
    NOISE.load(noise)
    obses = tf.convert_to_tensor(obs0)
    model.reset(obses)
    actions, _ = pol.compute_action(ppc.tf_process_obses(obses))
    o_list, r_list, a_list = [], [], [actions.numpy()]
    for t in range(n):
        obses, rewards = model.rollout_out(actions)
        actions, _ = pol.compute_action(ppc.tf_process_obses(obses))
        o_list.append(obses.numpy()); r_list.append(ppc.tf_process_rewards(rewards).numpy()); a_list.append(actions.numpy())
    out['closed_obs'], out['closed_rew'], out['closed_act'] = np.stack(o_list), np.stack(r_list), np.stack(a_list)
    return out


def run_apply_gradients(case):
    """PolicyWithQs.apply_gradients (policy.py:123-171) for several iterations on seeded gradients."""
    from policy import PolicyWithQs
    H = case['H']
    args = default_args(case['version'], PT, value_num_hidden_units=H, policy_num_hidden_units=H)
    dq = case['version'] == 'MPG-v2'
    pol = PolicyWithQs(**vars(args))
    pol.set_weights(synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=dq))
    rng = np.random.default_rng(case['gseed'])
    for it in range(case['iters']):
        grads = [rng.standard_normal(np.shape(a)).astype(np.float32) * 0.1
                 for net in pol.get_weights()[: (3 if dq else 2)] for a in net]
        pol.apply_gradients(it, grads)
    return {'weights': np.concatenate([np.asarray(a, np.float64).ravel() for net in pol.get_weights() for a in net])}


def run_real_env(case):
    """The reference's real PathTrackingEnv (reset(init_obs) / step incl. done) and MPGLearner (MPG-v1)
    compute_n_step_target with sample_num_in_learner = T: T real-env steps from the replay (obs, action) with the
    online policy, Q1_target / policy_target bootstrap (mpg_learner.py:87-124,146-169)."""
    from envs_and_models.path_tracking_env import PathTrackingEnv
    from learners.mpg_learner import MPGLearner
    from policy import PolicyWithQs
    B, H, T, nfd = case['B'], case['H'], case['T'], case['nfd']
    rng = np.random.default_rng(case['bseed'])
    obs0 = synthetic.make_obs(rng, PT, B, nfd)
    acts = rng.uniform(-1.2, 1.2, (case['n_env'], B, 2)).astype(np.float32)   # some beyond the clip range
    env = PathTrackingEnv(num_future_data=nfd, num_agent=B)
    env.reset(init_obs=obs0.copy())
    o_l, r_l, d_l = [], [], []
    for t in range(case['n_env']):
        o, r, d, _ = env.step(acts[t])
        o_l.append(np.asarray(o)); r_l.append(np.asarray(r)); d_l.append(np.asarray(d))
    out = {'env_obs': np.stack(o_l), 'env_rew': np.stack(r_l), 'env_done': np.stack(d_l).astype(np.int32)}
    args = default_args('MPG-v1', PT, replay_batch_size=B, num_future_data=nfd, value_num_hidden_units=H,
                        policy_num_hidden_units=H, sample_num_in_learner=T)
    learner = MPGLearner(PolicyWithQs, args)
    learner.set_weights(synthetic.make_policy_with_qs_weights(case['wseed'], args.obs_dim, args.act_dim, H, double_q=False))
    batch = make_batch(case['bseed'] + 1, PT, B, nfd)
    learner.get_batch_data(batch, None, None)
    out['batch_targets'] = np.asarray(learner.batch_data['batch_targets'])
    return out


def run_replay(case):
    """The reference's own PrioritizedReplayBuffer / segment trees (pure Python, imported unmodified) driven through
    add(weight=None), find_prefixsum_idx on explicit masses, sample_with_weights_and_idxes and update_priorities.
    args.alpha / args.size are supplied because the reference's parsers never define them."""
    from argparse import Namespace
    from buffer import PrioritizedReplayBuffer
    rng = np.random.default_rng(case['seed'])
    cap, n0, n1, ns = case['capacity'], case['n_add0'], case['n_add1'], case['n_sample']
    args = Namespace(max_buffer_size=cap, replay_starts=1, replay_batch_size=ns, alpha=0.6, size=cap, replay_alpha=0.6,
                     replay_beta=0.4, buffer_log_interval=10 ** 9)
    rb = PrioritizedReplayBuffer(args, 0)
    out = {}

    def add(n):
        for _ in range(n):
            o, a = rng.standard_normal(6).astype(np.float32), rng.standard_normal(2).astype(np.float32)
            rb.add(o, a, np.float32(rng.standard_normal()), o + 1, np.float32(0.0), None)

    def draw(tag):
        u = rng.random(ns).astype(np.float32)
        total = rb._it_sum.sum()
        idx = np.array([rb._it_sum.find_prefixsum_idx(float(x) * total) for x in u], np.int32)
        smp = rb.sample_with_weights_and_idxes(idx)
        out[f'u_{tag}'], out[f'idx_{tag}'], out[f'w_{tag}'] = u, idx, np.asarray(smp[5], np.float64)
        out[f'rew_{tag}'], out[f'obs_{tag}'] = np.asarray(smp[2], np.float32), np.asarray(smp[0], np.float32)
        out[f'sum_{tag}'], out[f'min_{tag}'] = np.float64(total), np.float64(rb._it_min.min())

    add(n0)
    draw('a')
    upd_idx = rng.integers(0, n0, case['n_update']).astype(np.int32)
    upd_pr = (np.abs(rng.standard_normal(case['n_update'])) + 1e-3).astype(np.float32)
    rb.update_priorities(upd_idx, [float(p) for p in upd_pr])
    out['upd_idx'], out['upd_pr'] = upd_idx, upd_pr
    draw('b')
    add(n1)          # wraps around the ring when n0 + n1 > capacity; new entries take the running max priority
    draw('c')
    out['max_priority'] = np.float64(rb._max_priority)
    return out


def run_weights_rule(case):
    """MPGLearner.rule_based_weights (mpg_learner.py:384-399) at several iterations."""
    from learners.mpg_learner import MPGLearner
    from policy import PolicyWithQs
    args = default_args('MPG-v2', 'PathTracking-v0', replay_batch_size=4, value_num_hidden_units=8,
                        policy_num_hidden_units=8, num_rollout_list_for_policy_update=case['rollout_list'])
    learner = MPGLearner(PolicyWithQs, args)
    tf = learner.tf
    ws = [learner.rule_based_weights(tf.convert_to_tensor(float(i), dtype=tf.float32),
                                     args.rule_based_bias_total_ite, args.eta).numpy() for i in case['iterations']]
    return {'ws': np.stack(ws)}


PT, IP, IDP = 'PathTracking-v0', 'InvertedPendulumConti-v0', 'InvertedDoublePendulum-v2'
CASES = {
    # learner cases at the real width (the CUDA path is checked against these directly)
    'nadp_pt_h256': dict(fn='nadp', env_id=PT, B=16, H=256, n=25, M=1, nfd=0, wseed=11, bseed=12, nseed=13),
    'mpg2_pt_h256': dict(fn='mpg', env_id=PT, B=16, H=256, M=1, nfd=0, rollout_list=[0, 25], iteration=4000,
                         wseed=21, bseed=22, nseed=23),
    # variants at a small width (pin the oracle; the CUDA path is then checked against the oracle)
    'nadp_pt_h64_m2_nfd2': dict(fn='nadp', env_id=PT, B=12, H=64, n=10, M=2, nfd=2, wseed=31, bseed=32, nseed=33,
                                buffer_type='priority'),
    'nadp_ip_h64': dict(fn='nadp', env_id=IP, B=12, H=64, n=25, M=1, nfd=0, wseed=41, bseed=42, nseed=43),
    'nadp_idp_h64': dict(fn='nadp', env_id=IDP, B=12, H=64, n=25, M=1, nfd=0, wseed=51, bseed=52, nseed=53),
    'mpg2_pt_h64_list3': dict(fn='mpg', env_id=PT, B=12, H=64, M=2, nfd=2, rollout_list=[0, 3, 25], iteration=5000,
                              wseed=61, bseed=62, nseed=63, buffer_type='priority'),
    'mpg2_pt_h64_deriv': dict(fn='mpg', env_id=PT, B=12, H=64, M=1, nfd=0, rollout_list=[0, 25], iteration=2000,
                              deriv_interval_policy=True, wseed=71, bseed=72, nseed=73),
    'mpg1_pt_h64': dict(fn='mpg', env_id=PT, B=12, H=64, M=1, nfd=0, rollout_list=[0, 25], iteration=4500,
                        version='MPG-v1', wseed=81, bseed=82, nseed=83),
    'mpg2_ip_h64': dict(fn='mpg', env_id=IP, B=12, H=64, M=1, nfd=0, rollout_list=[0, 25], iteration=1500,
                        wseed=91, bseed=92, nseed=93),
    # model trajectories
    'model_pt': dict(fn='model', env_id=PT, B=32, H=64, n=25, nfd=0, wseed=101, bseed=102, nseed=103),
    'model_pt_nfd2': dict(fn='model', env_id=PT, B=32, H=64, n=25, nfd=2, wseed=111, bseed=112, nseed=113),
    'model_ip': dict(fn='model', env_id=IP, B=32, H=64, n=25, nfd=0, wseed=121, bseed=122, nseed=123),
    'model_idp': dict(fn='model', env_id=IDP, B=32, H=64, n=25, nfd=0, wseed=131, bseed=132, nseed=133),
    'apply_grads_v2_h64': dict(fn='apply', version='MPG-v2', H=64, iters=5, wseed=141, gseed=142),
    'apply_grads_nadp_h64': dict(fn='apply', version='NADP', H=64, iters=4, wseed=151, gseed=152),
    'real_env_h64': dict(fn='real_env', B=24, H=64, T=25, nfd=2, n_env=12, wseed=171, bseed=172),
    'real_env_h256': dict(fn='real_env', B=16, H=256, T=25, nfd=0, n_env=4, wseed=181, bseed=182),
    'replay': dict(fn='replay', capacity=64, n_add0=40, n_add1=40, n_sample=128, n_update=50, seed=161),
    'rule_weights': dict(fn='rule', rollout_list=[0, 25], iterations=[0, 2000, 4000, 4500, 5000, 9000, 27000]),
    'rule_weights3': dict(fn='rule', rollout_list=[0, 3, 25], iterations=[0, 3000, 4500, 6000, 12000]),
}
FNS = dict(nadp=run_nadp, mpg=run_mpg, model=run_model, rule=run_weights_rule, apply=run_apply_gradients,
           replay=run_replay, real_env=run_real_env)


def extract_mpc_fixture():
    d = np.load('/root/reference/mpc/mpc_rl.npy', allow_pickle=True)
    np.savez_compressed(os.path.join(HERE, 'mpc_rl_transitions.npz'),
                        mpc_obs=np.stack([e['mpc_obs'][0] for e in d]).astype(np.float32),
                        mpc_action=np.stack([np.asarray(e['mpc_action'], np.float64) for e in d]),
                        rl_obs=np.stack([e['rl_obs'][0] for e in d]).astype(np.float32),
                        rl_action=np.stack([np.asarray(e['rl_action'], np.float64) for e in d]))


def main():
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        res = {}
        for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
            tf_shim.set_dtype(dt)
            out = FNS[case['fn']](case)
            for k, v in out.items():
                v = np.asarray(v)
                # big gradient vectors (H=256 cases): keep only the fp64 truth, rounded to fp32, and
                # drop the concatenated compute_gradient copy (its clip norms are kept as stats)
                if v.size > 16384:
                    if tag == 'f32' or k == 'compute_gradient':
                        continue
                    v = v.astype(np.float32)
                res[f'{k}__{tag}'] = v
        res['case_json'] = np.asarray(json.dumps(case))
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)
        print('wrote', name, {k: v.shape for k, v in res.items() if k.endswith('f32')})
    if not only:
        extract_mpc_fixture()
        print('wrote mpc_rl_transitions')


if __name__ == '__main__':
    main()

"""TEST INFRASTRUCTURE ONLY (the checker) -- never imported by the product path (mpg_b200/).

CPU oracle for the MPG model-based learner hot path: a from-scratch PyTorch-CPU restatement of the
reference's algorithm, dtype-parametric (float64 = "truth", float32 = reference precision), with
gradients by autograd.  Each function cites the reference file:line it follows
(paths relative to /root/reference).

PINNING: the reference has no tests and its arithmetic lives in un-pinned
third-party TensorFlow 2.x / TensorFlow-Probability, which cannot be installed here.  The oracle is
pinned instead against
  (1) golden vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON under a torch-backed stand-in
      for the TF API (oracle/tf_shim.py, tests/golden/make_golden.py -> tests/golden/*.npz):
      model trajectories, closed-loop rollouts, Q targets, TD errors, Q / policy gradients and the
      full compute_gradient() output of NADPLearner and MPGLearner (v1, v2, M>1, nfd>0,
      deriv_interval_policy, 3 envs);
  (2) the reference's only recorded numeric data, mpc/mpc_rl.npy (v_x', v_y', r', x' of f_xu);
  (3) the numpy restatement of rule_based_weights inside learners/mpg_learner.py:458-477.
tests/test_oracle_golden.py runs these checks on CPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# networks: model.py:20-43, policy.py:193-241
# ----------------------------------------------------------------------------------------------


def to_t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x, dtype=np.float64)).to(dtype)


def mlp(w, x, out_act='linear'):
    """MLPNet.call (model.py:39-43): Dense(elu) -> Dense(elu) -> Dense(out_act); kernels (in,out)."""
    h = F.elu(x @ w[0] + w[1])
    h = F.elu(h @ w[2] + w[3])
    y = h @ w[4] + w[5]
    return torch.tanh(y) if out_act == 'tanh' else y


def policy_action(w_pi, p, out_act, action_range):
    """PolicyWithQs.compute_action, deterministic branch (policy.py:193-199): mean half of the
    logits; action_range*tanh(mean) when action_range is set."""
    logits = mlp(w_pi, p, out_act)
    mean = logits[:, : logits.shape[1] // 2]
    return action_range * torch.tanh(mean) if action_range is not None else mean


def q_value(w_q, p, a):
    """PolicyWithQs.compute_Q1 (policy.py:219-223): MLP(concat(obs, act))[:, 0]."""
    return mlp(w_q, torch.cat([p, a], 1))[:, 0]


# ----------------------------------------------------------------------------------------------
# environment models
# ----------------------------------------------------------------------------------------------
class PathTrackingModel:
    """envs_and_models/path_tracking_env.py:245-297 (+ VehicleDynamics :58-138, :181-199)."""
    C_f, C_r, a, b, mass, I_z = -128915.5, -85943.6, 1.06, 1.85, 1412.0, 1536.7
    noise_mean, noise_std = 0.5, 0.01
    state_dim = 6

    def __init__(self, num_future_data=0, **kw):
        self.nfd = num_future_data

    def reset(self, obses):
        # _get_state (:273-277): v_x = delta_v_x + 20
        o = obses
        self.s = torch.stack([o[:, 0] + 20.0, o[:, 1], o[:, 2], o[:, 3], o[:, 4], o[:, 5]], 1)

    def get_obs(self, s):
        # _get_obs (:265-271): future columns replicate delta_y
        cols = [s[:, 0] - 20.0, s[:, 1], s[:, 2], s[:, 3], s[:, 4], s[:, 5]] + [s[:, 3]] * self.nfd
        return torch.stack(cols, 1)

    def compute_rewards(self, s, u):
        # VehicleDynamics.compute_rewards (:181-199) on the SCALED action
        return -(0.01 * (s[:, 0] - 20.0) ** 2 + 0.04 * s[:, 3] ** 2 + 0.1 * s[:, 4] ** 2
                 + 0.02 * s[:, 2] ** 2 + 5 * u[:, 0] ** 2 + 0.05 * u[:, 1] ** 2)

    def f_xu(self, s, u, tau, eps=None):
        # VehicleDynamics.f_xu, if_model branch (:113-122)
        v_x, v_y, r, dy, dphi, x = (s[:, i] for i in range(6))
        steer, a_x = u[:, 0], u[:, 1]
        C_f, C_r, a, b, m, I_z = self.C_f, self.C_r, self.a, self.b, self.mass, self.I_z
        nxt = [
            v_x + tau * (a_x + v_y * r),
            (m * v_y * v_x + tau * (a * C_f - b * C_r) * r - tau * C_f * steer * v_x - tau * m * v_x ** 2 * r)
            / (m * v_x - tau * (C_f + C_r)),
            (-I_z * r * v_x - tau * (a * C_f - b * C_r) * v_y + tau * a * C_f * steer * v_x)
            / (tau * (a ** 2 * C_f + b ** 2 * C_r) - I_z * v_x),
            dy + tau * (v_x * torch.sin(dphi) + v_y * torch.cos(dphi))
            + (0.0 if eps is None else (self.noise_mean + self.noise_std * eps)),
            dphi + tau * r,
            x + tau * (v_x * torch.cos(dphi) - v_y * torch.sin(dphi)),
        ]
        return torch.stack(nxt, 1)

    def rollout_out(self, actions, eps):
        # rollout_out (:279-297): scale action, reward on the PRE-step state, step, clip v_x, wrap dphi
        u = torch.stack([actions[:, 0] * 1.2 * math.pi / 9, actions[:, 1] * 3.0], 1)
        rew = self.compute_rewards(self.s, u)
        s = self.f_xu(self.s, u, 1.0 / 10.0, eps)
        v_x = torch.clamp(s[:, 0], 1.0, 35.0)
        dphi = s[:, 4]
        dphi = torch.where(dphi > math.pi, dphi - 2 * math.pi, dphi)
        dphi = torch.where(dphi <= -math.pi, dphi + 2 * math.pi, dphi)
        self.s = torch.stack([v_x, s[:, 1], s[:, 2], s[:, 3], dphi, s[:, 5]], 1)
        return self.get_obs(self.s), rew


class InvertedPendulumModel:
    """envs_and_models/inverted_pendulum_model.py:16-97. tf.linalg.inv of the 2x2 mass matrix is
    restated in closed form (adjugate / determinant), which is what the CUDA kernel implements."""
    noise_mean, noise_std = 0.1, 0.5
    state_dim = 4

    def __init__(self, **kw):
        self.tau = 0.04

    def reset(self, obses):
        self.s = obses

    def get_obs(self, s):
        return s

    @staticmethod
    def consts(dtype):
        # :18-25, :37-44 evaluated in the working dtype, in the reference's operation order
        t = lambda v: torch.tensor(v, dtype=dtype)
        m, m1, m2, l1, g = t(9.42), t(4.89), t(0.0), t(0.6), t(9.81)
        d1 = m + m1 + m2
        d2 = (0.5 * m1 + m2) * l1
        d4 = (1. / 3 * m1 + m2) * l1 ** 2
        f1 = (0.5 * m1 + m2) * l1 * g
        return d1, d2, d4, f1

    def f_xu(self, s, u, tau, eps=None):
        p, th, pd, thd = (s[:, i] for i in range(4))
        d1, d2, d4, f1 = self.consts(s.dtype)
        c, sn = torch.cos(th), torch.sin(th)
        f0 = d2 * sn * thd ** 2 + u[:, 0]
        f1v = f1 * sn
        det = d1 * d4 - (d2 * c) ** 2
        pdd = (d4 * f0 - d2 * c * f1v) / det
        thdd = (-d2 * c * f0 + d1 * f1v) / det
        nxt = [p + tau * pd + (0.0 if eps is None else (self.noise_mean + self.noise_std * eps)),
               th + tau * thd, pd + tau * pdd, thd + tau * thdd]
        return torch.stack(nxt, 1)

    def compute_rewards(self, s):
        return -(0.01 * s[:, 0] ** 2 + s[:, 1] ** 2) - (1e-3 * s[:, 2] ** 2 + 1e-3 * s[:, 3] ** 2)

    def rollout_out(self, actions, eps):
        # :88-94: u = 100 a; reward on the POST-step state
        self.s = self.f_xu(self.s, 100.0 * actions, self.tau, eps)
        return self.s, self.compute_rewards(self.s)


class InvertedDoublePendulumModel:
    """envs_and_models/inverted_double_pendulum_model.py:14-53, 89-144 (f_xu_old, 5 sub-steps)."""
    state_dim = 6
    noise_mean, noise_std = 0.0, 0.0

    def __init__(self, **kw):
        self.tau = 0.01

    def reset(self, obses):
        o = obses  # _get_state (:126-132)
        self.s = torch.stack([o[:, 0], torch.atan2(o[:, 1], o[:, 3]), torch.atan2(o[:, 2], o[:, 4]),
                              o[:, 5], o[:, 6], o[:, 7]], 1)

    def get_obs(self, s):
        z = torch.zeros_like(s[:, 0])  # _get_obs (:118-124)
        return torch.stack([s[:, 0], torch.sin(s[:, 1]), torch.sin(s[:, 2]), torch.cos(s[:, 1]),
                            torch.cos(s[:, 2]), s[:, 3], s[:, 4], s[:, 5], z, z, z], 1)

    def f_xu(self, s, u, tau, eps=None):
        # f_xu_old (:26-53); M^-1 f restated via the adjugate (closed form of tf.linalg.inv)
        dt = s.dtype
        t = lambda v: torch.tensor(v, dtype=dt)
        m, m1, m2, l1, l2, g = t(9.42477796), t(4.1033127), t(4.1033127), t(0.6), t(0.6), t(9.81)
        p, t1, t2, pd, t1d, t2d = (s[:, i] for i in range(6))
        a11 = (m + m1 + m2) * torch.ones_like(p)
        a12 = l1 * (m1 + m2) * torch.cos(t1)
        a13 = m2 * l2 * torch.cos(t2)
        a22 = (l1 ** 2) * (m1 + m2) * torch.ones_like(p)
        a23 = l1 * l2 * m2 * torch.cos(t1 - t2)
        a33 = (l2 ** 2) * m2 * torch.ones_like(p)
        f0 = l1 * (m1 + m2) * t1d ** 2 * torch.sin(t1) + m2 * l2 * t2d ** 2 * torch.sin(t2) + u[:, 0]
        f1 = -l1 * l2 * m2 * t2d ** 2 * torch.sin(t1 - t2) + g * (m1 + m2) * l1 * torch.sin(t1)
        f2 = l1 * l2 * m2 * t1d ** 2 * torch.sin(t1 - t2) + g * l2 * m2 * torch.sin(t2)
        # symmetric 3x3 inverse via cofactors
        c11 = a22 * a33 - a23 * a23
        c12 = a13 * a23 - a12 * a33
        c13 = a12 * a23 - a13 * a22
        c22 = a11 * a33 - a13 * a13
        c23 = a12 * a13 - a11 * a23
        c33 = a11 * a22 - a12 * a12
        det = a11 * c11 + a12 * c12 + a13 * c13
        q0 = (c11 * f0 + c12 * f1 + c13 * f2) / det
        q1 = (c12 * f0 + c22 * f1 + c23 * f2) / det
        q2 = (c13 * f0 + c23 * f1 + c33 * f2) / det
        return torch.stack([p + tau * pd, t1 + tau * t1d, t2 + tau * t2d,
                            pd + tau * q0, t1d + tau * q1, t2d + tau * q2], 1)

    def compute_rewards(self, s):
        # :89-100 (l_rod1 = l_rod2 = 0.6 as python floats)
        tip_x = s[:, 0] + 0.6 * torch.sin(s[:, 1]) + 0.6 * torch.sin(s[:, 2])
        tip_y = 0.6 * torch.cos(s[:, 1]) + 0.6 * torch.cos(s[:, 2])
        return -(0.01 * tip_x ** 2 + (tip_y - 2) ** 2) - (1e-3 * s[:, 4] ** 2 + 5e-3 * s[:, 5] ** 2)

    def rollout_out(self, actions, eps):
        u = 500.0 * actions  # :134-141
        for _ in range(5):
            self.s = self.f_xu(self.s, u, self.tau)
        return self.get_obs(self.s), self.compute_rewards(self.s)


NAME2MODELCLS = {'PathTracking-v0': PathTrackingModel,
                 'InvertedDoublePendulum-v2': InvertedDoublePendulumModel,
                 'InvertedPendulumConti-v0': InvertedPendulumModel}


# ----------------------------------------------------------------------------------------------
# rollout: learners/mpg_learner.py:180-286, learners/nadp.py:87-171
# ----------------------------------------------------------------------------------------------
def rollout(args, dtype, w_first, w_rest, w_q, obs, noise, n, start_actions=None, keep_traj=False):
    """The shared rollout loop.
      w_first : policy weights for a_0 (ignored when start_actions is given)
      w_rest  : policy weights for a_1..a_n (policy_for_rollout in default MPG, same as w_first in
                NADP / deriv_interval_policy)
      w_q     : Q net used for the bootstrap on all n+1 steps (Q1 or Q1_target)
      noise   : (n, M*B) standard-normal eps or None
    Returns all_model_returns (n+1, B) = mean over the M tiles of R_t + gamma^t Q(p_t, a_t)
    (mpg_learner.py:264-272, nadp.py:154-164) and optionally the trajectory.
    """
    M = args.M
    sigma = to_t(args.obs_scale, dtype)
    model = NAME2MODELCLS[args.env_id](**vars(args))
    obs_t = obs.repeat(M, 1)
    p = obs_t * sigma  # preprocessor.py:134-145 (scale)
    if start_actions is None:
        a = policy_action(w_first, p, args.policy_out_activation, args.action_range)
    else:
        a = start_actions.repeat(M, 1)
    p_list, a_list = [p], [a]
    rsum = torch.zeros(obs_t.shape[0], dtype=dtype)
    rsum_list, traj_obs, traj_rew = [rsum], [], []
    model.reset(obs_t)
    for ri in range(n):
        obs_t, rew = model.rollout_out(a, None if noise is None else noise[ri])
        p = obs_t * sigma
        prew = (rew + args.rew_shift) * args.rew_scale  # preprocessor.py:147-159
        rsum = rsum + torch.pow(torch.tensor(args.gamma, dtype=dtype), ri) * prew
        rsum_list.append(rsum)
        a = policy_action(w_rest, p, args.policy_out_activation, args.action_range)
        p_list.append(p)
        a_list.append(a)
        if keep_traj:
            traj_obs.append(obs_t)
            traj_rew.append(prew)
    gammas = torch.cat([torch.pow(torch.tensor(args.gamma, dtype=dtype), t) * torch.ones_like(rsum)
                        for t in range(n + 1)])
    if w_q is not None:
        all_q = q_value(w_q, torch.cat(p_list, 0), torch.cat(a_list, 0))
    else:  # AMPC (learners/ampc.py:73-87): no bootstrap
        all_q = torch.zeros_like(gammas)
    final = (torch.cat(rsum_list, 0) + gammas * all_q).reshape(n + 1, M, -1)
    returns = final.mean(1)
    if keep_traj:
        return returns, dict(obs=traj_obs, rew=traj_rew, act=a_list)
    return returns


def clip_by_global_norm(grads, clip):
    """tf.clip_by_global_norm: g * clip / max(norm, clip); returns (clipped, norm)."""
    norm = torch.sqrt(sum((g * g).sum() for g in grads))
    scale = clip * torch.minimum(1.0 / norm, torch.tensor(1.0 / clip, dtype=norm.dtype))
    return [g * scale for g in grads], norm


def rule_based_weights(ite, total_ite, eta, rollout_list, dtype=torch.float32):
    """MPGLearner.rule_based_weights (mpg_learner.py:384-399), evaluated in `dtype` like the TF graph."""
    t = lambda v: torch.tensor(v, dtype=dtype)
    lam = torch.clamp(t(1. - eta) + t(2. * eta / total_ite) * t(float(ite)), 0, 1.5)
    if lam < 1.:
        biases = torch.stack([torch.pow(lam, i) for i in rollout_list])
    else:
        mx = max(rollout_list)
        biases = torch.stack([torch.pow(2 - lam, mx - i) for i in rollout_list])
    return torch.softmax(1. / (biases + 1e-8), -1)


def _leaf(ws, dtype):
    return [to_t(w, dtype).clone().requires_grad_(True) for w in ws]


class Nets:
    """Weights in PolicyWithQs.get_weights() order (policy.py:72-89,112-121)."""

    def __init__(self, weights, double_q, dtype):
        ws = [_leaf(w, dtype) for w in weights]
        if double_q:
            self.Q1, self.Q2, self.policy, self.Q1_t, self.Q2_t, self.policy_t = ws
        else:
            self.Q1, self.policy, self.Q1_t, self.policy_t = ws
            self.Q2 = self.Q2_t = None


def td_error(args, nets, batch, dtype):
    """compute_td_error (mpg_learner.py:136-144, nadp.py:67-76)."""
    sigma = to_t(args.obs_scale, dtype)
    obs, act, rew, obs1 = (to_t(b, dtype) for b in batch[:4])
    with torch.no_grad():
        p, p1 = obs * sigma, obs1 * sigma
        prew = (rew + args.rew_shift) * args.rew_scale
        a1 = policy_action(nets.policy_t, p1, args.policy_out_activation, args.action_range)
        return prew + args.gamma * q_value(nets.Q1_t, p1, a1) - q_value(nets.Q1, p, act)


def nadp_compute_gradient(args, weights, batch, noise_q, noise_p, dtype=torch.float64):
    """NADPLearner.compute_gradient (nadp.py:209-241) and its parts. Returns a dict of numpy values."""
    nets = Nets(weights, False, dtype)
    sigma = to_t(args.obs_scale, dtype)
    obs, act = to_t(batch[0], dtype), to_t(batch[1], dtype)
    nq = max(args.num_rollout_list_for_q_estimation)
    npol = max(args.num_rollout_list_for_policy_update)
    out = {}
    # q_forward_and_backward (nadp.py:173-184): forward-only rollout target with Q1_target
    with torch.no_grad():
        ret = rollout(args, dtype, None, nets.policy, nets.Q1_t, obs, None if noise_q is None else to_t(noise_q, dtype),
                      nq, start_actions=act)
        targets = torch.cat([ret[k] for k in args.num_rollout_list_for_q_estimation], 0)
    q_pred = q_value(nets.Q1, obs * sigma, act)
    q_loss = 0.5 * torch.mean((q_pred - targets) ** 2)
    q_grad = torch.autograd.grad(q_loss, nets.Q1)
    # policy_forward_and_backward (nadp.py:186-194)
    ret = rollout(args, dtype, nets.policy, nets.policy, nets.Q1, obs,
                  None if noise_p is None else to_t(noise_p, dtype), npol)
    reduced = ret.mean(1)
    policy_loss = -reduced[args.num_rollout_list_for_policy_update[0]]
    p_grad = torch.autograd.grad(policy_loss, nets.policy)
    qc, qn = clip_by_global_norm(q_grad, args.gradient_clip_norm)
    pc, pn = clip_by_global_norm(p_grad, args.gradient_clip_norm)
    flat = lambda gs: np.concatenate([g.detach().numpy().ravel() for g in gs])
    out.update(q_targets=targets.numpy(), q_loss=q_loss.item(), q_grad=flat(q_grad),
               policy_loss=policy_loss.item(), policy_grad=flat(p_grad), value_mean=reduced[0].item(),
               compute_gradient=flat(list(qc) + list(pc)), q_gradient_norm=qn.item(),
               policy_gradient_norm=pn.item())
    if args.buffer_type != 'normal':
        out['td_error'] = td_error(args, nets, batch, dtype).numpy()
    return out


def mpg_compute_gradient(args, weights, batch, noise_p, iteration, dtype=torch.float64):
    """MPGLearner.compute_gradient (mpg_learner.py:401-455) and its parts."""
    dq = args.learner_version == 'MPG-v2'
    nets = Nets(weights, dq, dtype)
    sigma = to_t(args.obs_scale, dtype)
    obs, act, rew, obs1 = (to_t(b, dtype) for b in batch[:4])
    out = {}
    with torch.no_grad():
        p1 = obs1 * sigma
        prew = (rew + args.rew_shift) * args.rew_scale
        a1 = policy_action(nets.policy_t, p1, args.policy_out_activation, args.action_range)
        if dq:  # compute_clipped_double_q_target (:126-134)
            target = prew + args.gamma * torch.minimum(q_value(nets.Q1_t, p1, a1), q_value(nets.Q2_t, p1, a1))
        else:   # compute_n_step_target, sample_num_in_learner None branch (:147-152)
            target = prew + args.gamma * q_value(nets.Q1_t, p1, a1)
    out['batch_targets'] = target.numpy()
    flat = lambda gs: np.concatenate([g.detach().numpy().ravel() for g in gs])
    # q_forward_and_backward (:326-354)
    p = obs * sigma
    q_grads_clipped = []
    for i, wq in enumerate([nets.Q1, nets.Q2] if dq else [nets.Q1], 1):
        loss = 0.5 * torch.mean((q_value(wq, p, act) - target) ** 2)
        g = torch.autograd.grad(loss, wq)
        gc, gn = clip_by_global_norm(g, args.gradient_clip_norm)
        out[f'q_loss{i}'], out[f'q_grad{i}'], out[f'q_gradient_norm{i}'] = loss.item(), flat(g), gn.item()
        q_grads_clipped += list(gc)
    # policy_forward_and_backward (:356-365) with policy_for_rollout = detached copy (:422)
    lst = args.num_rollout_list_for_policy_update
    w_rest = nets.policy if args.deriv_interval_policy else [w.detach() for w in nets.policy]
    ret = rollout(args, dtype, nets.policy, w_rest, nets.Q1, obs,
                  None if noise_p is None else to_t(noise_p, dtype), max(lst))
    reduced = ret.mean(1)
    var = ret.var(1, unbiased=False)
    minus_ret = torch.stack([-reduced[k] for k in lst])
    ws = rule_based_weights(iteration, args.rule_based_bias_total_ite, args.eta, lst, dtype)
    total_loss = torch.sum(ws.detach() * minus_ret)
    p_grad = torch.autograd.grad(total_loss, nets.policy)
    pc, pn = clip_by_global_norm(p_grad, args.gradient_clip_norm)
    out.update(returns_var=torch.stack([var[k] for k in lst]).detach().numpy(),
               minus_returns=minus_ret.detach().numpy(), value_mean=ret[0].mean().item(),
               policy_grad=flat(p_grad), total_loss=total_loss.item(), ws=ws.numpy(),
               policy_gradient_norm=pn.item(), compute_gradient=flat(q_grads_clipped + list(pc)))
    if args.buffer_type != 'normal':
        out['td_error'] = td_error(args, nets, batch, dtype).numpy()
    return out


def closed_loop(args, w_pi, obs0, noise, n, dtype=torch.float64):
    """Closed-loop trajectory (obs_t, processed reward_t, a_t) -- the loop of nadp.py:141-152."""
    w = [to_t(x, dtype) for x in w_pi]
    with torch.no_grad():
        _, traj = rollout(args, dtype, w, w, None, to_t(obs0, dtype),
                          None if noise is None else to_t(noise, dtype), n, keep_traj=True)
    return (torch.stack(traj['obs']).numpy(), torch.stack(traj['rew']).numpy(), torch.stack(traj['act']).numpy())


def open_loop(args, obs0, acts, noise, dtype=torch.float64):
    """Open-loop model trajectory for given actions (n, B, d_a): raw obs and raw rewards."""
    model = NAME2MODELCLS[args.env_id](**vars(args))
    model.reset(to_t(obs0, dtype))
    o_l, r_l = [], []
    for t in range(acts.shape[0]):
        o, r = model.rollout_out(to_t(acts[t], dtype), None if noise is None else to_t(noise[t], dtype))
        o_l.append(o)
        r_l.append(r)
    return torch.stack(o_l).numpy(), torch.stack(r_l).numpy()


# ----------------------------------------------------------------------------------------------
# optimiser step: PolicyWithQs.apply_gradients, policy.py:123-171
# ----------------------------------------------------------------------------------------------
def polynomial_decay(schedule, step):
    """keras PolynomialDecay(initial, decay_steps, end), power 1, no cycle."""
    init, decay_steps, end = schedule
    frac = min(float(step), float(decay_steps)) / float(decay_steps)
    return (init - end) * (1.0 - frac) + end


class AdamState:
    """keras OptimizerV2 Adam for one net: b1 0.9, b2 0.999, eps 1e-7 (the algorithm lives in un-vendored,
    un-pinned TensorFlow 2.x; restated from its published definition)."""

    def __init__(self, schedule):
        self.schedule, self.iterations, self.m, self.v = schedule, 0, None, None

    def apply(self, weights, grads):
        t = self.iterations + 1
        lr = polynomial_decay(self.schedule, self.iterations)
        lr_t = lr * math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        if self.m is None:
            self.m = [np.zeros_like(w) for w in weights]
            self.v = [np.zeros_like(w) for w in weights]
        out = []
        for i, (w, g) in enumerate(zip(weights, grads)):
            self.m[i] = 0.9 * self.m[i] + (1.0 - 0.9) * g
            self.v[i] = 0.999 * self.v[i] + (1.0 - 0.999) * g * g
            out.append(w - lr_t * self.m[i] / (np.sqrt(self.v[i]) + 1e-7))
        self.iterations = t
        return out


def apply_gradients(weights, states, iteration, grads, double_q, delay_update, tau, dtype=np.float64):
    """weights: get_weights() order (models + targets); states: dict name -> AdamState; grads: flat list in
    compute_gradient order. Returns the new weights list (policy.py:123-171, target=True)."""
    w = [[np.asarray(a, dtype) for a in net] for net in weights]
    g = [np.asarray(a, dtype) for a in grads]
    nq = 2 if double_q else 1
    for i in range(nq):
        w[i] = states[f'Q{i + 1}'].apply(w[i], g[6 * i: 6 * i + 6])
    if iteration % delay_update == 0:
        w[nq] = states['policy'].apply(w[nq], g[6 * nq: 6 * nq + 6])
        nm = nq + 1
        for i in range(nm):
            w[nm + i] = [tau * s_ + (1.0 - tau) * t_ for s_, t_ in zip(w[i], w[nm + i])]
    return w


# ----------------------------------------------------------------------------------------------
# real PathTracking environment: path_tracking_env.py:144-179,202-242,356-487
# ----------------------------------------------------------------------------------------------
class PathTrackingEnvOracle:
    """numpy restatement of PathTrackingEnv.reset(init_obs) / step / _get_obs / judge_done with
    VehicleDynamics.simulation and ReferencePath, in `dtype` (the reference runs it in float32)."""
    curves = [(7.5, 200., 0.), (2.5, 300., 0.), (-5., 400., 0.)]
    period = 1200.

    def __init__(self, num_future_data=0, dtype=np.float64):
        self.nfd, self.dt = num_future_data, dtype

    def path_y(self, x):
        y = np.zeros_like(x)
        for mag, T, sh in self.curves:
            y = y + self.dt(mag) * np.sin((x - self.dt(sh)) * self.dt(2) * self.dt(np.pi) / self.dt(T))
        return y

    def path_phi(self, x):
        d = np.zeros_like(x)
        for mag, T, sh in self.curves:
            d = d + self.dt(mag * 2 * np.pi / T) * np.cos((x - self.dt(sh)) * self.dt(2) * self.dt(np.pi) / self.dt(T))
        return np.arctan(d)

    def reset(self, obs):
        o = np.asarray(obs, self.dt)
        self.veh = np.stack([o[:, 0] + 20, o[:, 1], o[:, 2], o[:, 3], o[:, 4], o[:, 5]], 1).astype(self.dt)
        x = self.veh[:, 5]
        self.full = self.veh.copy()
        self.full[:, 4] = self.veh[:, 4] + self.path_phi(x)
        self.full[:, 3] = self.veh[:, 3] + self.path_y(x)

    def get_obs(self):
        v, f = self.veh, self.full
        cols = [v[:, 0] - 20, v[:, 1], v[:, 2], v[:, 3], v[:, 4], f[:, 5]]
        x_ = f[:, 5].copy()
        for _ in range(self.nfd):
            x_ = x_ + f[:, 0] * self.dt(1. / 200) * self.dt(20 * 2)
            cols.append(f[:, 3] - self.path_y(x_))
        return np.stack(cols, 1)

    def step(self, action):
        dt, pi = self.dt, self.dt(np.pi)
        a = np.asarray(action, dt)
        u = np.stack([a[:, 0] * dt(1.2) * pi / dt(9), a[:, 1] * dt(3)], 1)
        lim = np.array([1.2 * np.pi / 9, 3], np.float32).astype(dt)
        u = np.clip(u, -lim, lim)
        m = PathTrackingModel()
        st = torch.as_tensor(self.veh)
        reward = m.compute_rewards(st, torch.as_tensor(u)).numpy()
        C_f, C_r, aa, bb, mass, I_z, g = m.C_f, m.C_r, m.a, m.b, m.mass, m.I_z, 9.81
        veh, full = self.veh.copy(), self.full.copy()
        for _ in range(20):
            vx_in = veh[:, 0].copy()
            alpha_f = np.arctan((veh[:, 1] + aa * veh[:, 2]) / veh[:, 0]) - u[:, 0]
            alpha_r = np.arctan((veh[:, 1] - bb * veh[:, 2]) / veh[:, 0])
            nxt = m.f_xu(torch.as_tensor(veh), torch.as_tensor(u), 1 / 200).numpy()
            nxt[:, 0] = np.clip(nxt[:, 0], 1, 35)
            vxs, vys, rs = full[:, 0].copy(), full[:, 1].copy(), full[:, 2].copy()
            full[:, 4] = full[:, 4] + rs / dt(200)
            phis = full[:, 4]                      # numpy view semantics of the reference: UPDATED phi below
            full[:, 3] = full[:, 3] + (vxs * np.sin(phis) + vys * np.cos(phis)) / dt(200)
            full[:, 5] = full[:, 5] + (vxs * np.cos(phis) - vys * np.sin(phis)) / dt(200)
            full[:, 0:3] = nxt[:, 0:3]
            py, pphi = self.path_y(full[:, 5]), self.path_phi(full[:, 5])
            nxt[:, 4] = full[:, 4] - pphi
            nxt[:, 3] = full[:, 3] - py
            full[:, 4] = np.where(full[:, 4] > pi, full[:, 4] - 2 * pi, full[:, 4])
            full[:, 4] = np.where(full[:, 4] <= -pi, full[:, 4] + 2 * pi, full[:, 4])
            full[:, 5] = np.where(full[:, 5] > self.period, full[:, 5] - dt(self.period), full[:, 5])
            full[:, 5] = np.where(full[:, 5] <= 0, full[:, 5] + dt(self.period), full[:, 5])
            nxt[:, 5] = full[:, 5]
            nxt[:, 4] = np.where(nxt[:, 4] > pi, nxt[:, 4] - 2 * pi, nxt[:, 4])
            nxt[:, 4] = np.where(nxt[:, 4] <= -pi, nxt[:, 4] + 2 * pi, nxt[:, 4])
            veh = nxt.astype(dt)
        self.veh, self.full = veh, full
        F_zf, F_zr = bb * mass * g / (aa + bb), aa * mass * g / (aa + bb)
        ax = u[:, 1]
        F_xf = np.where(ax < 0, mass * ax / 2, 0.0)
        F_xr = np.where(ax < 0, mass * ax / 2, mass * ax)
        miu_f, miu_r = np.sqrt(F_zf ** 2 - F_xf ** 2) / F_zf, np.sqrt(F_zr ** 2 - F_xr ** 2) / F_zr
        af_b, ar_b, r_b = 3 * miu_f * F_zf / C_f, 3 * miu_r * F_zr / C_r, miu_r * g / np.abs(vx_in)
        r = veh[:, 2]
        done = (np.abs(veh[:, 3]) > 3) | (np.abs(veh[:, 4]) > np.pi / 4) | (veh[:, 0] < 2) | (alpha_f < -af_b) | \
               (alpha_f > af_b) | (alpha_r < -ar_b) | (alpha_r > ar_b) | (r < -r_b) | (r > r_b)
        return self.get_obs(), reward, done


def mpg_v1_n_step_target(args, weights, batch, T, dtype=torch.float64):
    """MPGLearner.compute_n_step_target with sample_num_in_learner = T (mpg_learner.py:87-124,146-169)."""
    nets = Nets(weights, False, dtype)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    sigma = to_t(args.obs_scale, dtype)
    env = PathTrackingEnvOracle(args.num_future_data, npdt)
    obs = np.asarray(batch[0], npdt)
    env.reset(obs)
    target = torch.zeros(obs.shape[0], dtype=dtype)
    with torch.no_grad():
        for t in range(T):
            a = policy_action(nets.policy, to_t(obs, dtype) * sigma, args.policy_out_activation, args.action_range).numpy()
            if t == 0:
                a = np.asarray(batch[1], npdt)
            obs, rew, _ = env.step(a)
            target = target + (args.gamma ** t) * (to_t(rew, dtype) + args.rew_shift) * args.rew_scale
        p = to_t(obs, dtype) * sigma
        a_t = policy_action(nets.policy_t, p, args.policy_out_activation, args.action_range)
        target = target + (args.gamma ** T) * q_value(nets.Q1_t, p, a_t)
    return target.numpy(), obs

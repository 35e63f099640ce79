"""Evaluator with the reference's interface (evaluator.py:25-235): fixed-step deterministic episodes in the real
environment, metrics written as JSON lines instead of tf.summary."""
import json
import os

import numpy as np
import torch

from .envs_and_models import PathTrackingEnv
from .preprocessor import Preprocessor
from .utils.misc import TimerStat


class Evaluator(object):
    def __init__(self, policy_cls, env_id, args, policy=None):
        self.args = args
        if self.args.env_id != 'PathTracking-v0':
            raise NotImplementedError('only the PathTracking ground-truth environment is built')
        self.env = PathTrackingEnv(num_agent=self.args.num_eval_agent, num_future_data=self.args.num_future_data, seed=12345)
        self.policy_with_value = policy if policy is not None else policy_cls(**vars(self.args))
        self.preprocessor = Preprocessor(self.args.obs_dim, self.args.obs_ptype, self.args.rew_ptype, self.args.obs_scale,
                                         self.args.rew_scale, self.args.rew_shift, gamma=self.args.gamma)
        self.log_dir = getattr(self.args, 'log_dir', None)
        self.stats, self.eval_timer, self.eval_times, self.iteration = {}, TimerStat(), 0, 0

    def get_stats(self):
        self.stats.update(dict(eval_time=self.eval_timer.mean))
        return self.stats

    def set_weights(self, weights):
        self.policy_with_value.set_weights(weights)

    def set_ppc_params(self, params):
        self.preprocessor.set_params(params)

    def run_n_episodes(self, n=None, seed=777, fused=True):
        """n batches (default 1) of num_eval_agent parallel episodes of args.fixed_steps steps with the deterministic
        policy (compute_mode), averaged (evaluator.py:115-145 runs n episodes one after the other).
        fused: a whole batch of episodes is one launch (mpg_env_sample without exploration noise and without restarts)."""
        n = max(1, int(n or 1))
        if n > 1:
            runs = [self.run_n_episodes(1, seed + k, fused) for k in range(n)]
            return {k: (runs[0][k] if k == 'episode_len' else float(np.mean([r[k] for r in runs]))) for k in runs[0]}
        self.env._rng = np.random.default_rng(seed)      # same start states at every evaluation
        obs = self.env.reset()
        steps, agents = self.args.fixed_steps, obs.shape[0]
        if fused:
            _, _, rew, obs_tp1, _ = self.policy_with_value.engine.env_sample(self.env.state, self.env.obs, None, None, 0.0,
                                                                             steps=steps)
            rew, o = rew.view(steps, agents), obs_tp1.view(steps, agents, -1)
        else:
            rews, obss = [], []
            for _ in range(steps):
                action = self.policy_with_value.compute_mode(self.preprocessor.torch_process_obses(obs))
                obs, reward, done, _ = self.env.step(action)
                rews.append(reward); obss.append(obs)
            rew, o = torch.stack(rews), torch.stack(obss)
        return dict(episode_return=float(rew.sum(0).mean()), episode_len=steps,
                    delta_y_mean=float(o[:, :, 3].abs().mean()), delta_phi_mean=float(o[:, :, 4].abs().mean()),
                    delta_v_mean=float(o[:, :, 0].abs().mean()))

    def run_evaluation(self, iteration):
        with self.eval_timer:
            self.iteration = iteration
            metrics = self.run_n_episodes(self.args.num_eval_episode)
        self.eval_times += 1
        if self.log_dir:
            os.makedirs(self.log_dir, exist_ok=True)
            with open(os.path.join(self.log_dir, 'evaluator.jsonl'), 'a') as f:
                f.write(json.dumps(dict(iteration=iteration, **metrics)) + '\n')
        return metrics

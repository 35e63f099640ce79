"""OffPolicyWorker with the reference's interface (worker.py:25-123): env stepping with the exploration policy and
ownership of the master weights / optimiser, here on the GPU (real PathTracking env kernel + policy forward kernel)."""
import logging

import torch

from .envs_and_models import PathTrackingEnv
from .preprocessor import Preprocessor

logger = logging.getLogger(__name__)


class OffPolicyWorker(object):
    def __init__(self, policy_cls, env_id, args, worker_id, policy=None):
        self.worker_id, self.args = worker_id, args
        self.num_agent = self.args.num_agent
        if self.args.env_id != 'PathTracking-v0':
            raise NotImplementedError('workers need the ground-truth environment; only PathTracking-v0 is built '
                                      '(the pendulum envs are mujoco simulations)')
        self.env = PathTrackingEnv(num_agent=self.num_agent, num_future_data=self.args.num_future_data,
                                   seed=getattr(args, 'seed', 0) + worker_id)
        self.policy_with_value = policy if policy is not None else policy_cls(**vars(self.args))
        self.batch_size = self.args.batch_size
        self.obs = self.env.reset()
        self.preprocessor = Preprocessor(self.args.obs_dim, self.args.obs_ptype, self.args.rew_ptype, self.args.obs_scale,
                                         self.args.rew_scale, self.args.rew_shift, gamma=self.args.gamma)
        self.explore_sigma = self.args.explore_sigma
        self.iteration = self.num_sample = self.sample_times = 0
        self.stats = {}
        self.generator = torch.Generator(device=self.env.engine.device)
        self.generator.manual_seed(1000 + worker_id)

    def get_stats(self):
        self.stats.update(dict(worker_id=self.worker_id, num_sample=self.num_sample))
        return self.stats

    def get_weights(self):
        return self.policy_with_value.get_weights()

    def set_weights(self, weights):
        return self.policy_with_value.set_weights(weights)

    def apply_gradients(self, iteration, grads):
        self.iteration = iteration
        self.policy_with_value.apply_gradients(iteration, grads)

    def save_weights(self, save_dir, iteration):
        self.policy_with_value.save_weights(save_dir, iteration)

    def load_weights(self, load_dir, iteration):
        self.policy_with_value.load_weights(load_dir, iteration)

    def save_ppc_params(self, save_dir):
        self.preprocessor.save_params(save_dir)

    def get_ppc_params(self):
        return self.preprocessor.get_params()

    def set_ppc_params(self, params):
        self.preprocessor.set_params(params)

    def sample_arrays(self, fused=True):
        """worker.py:91-119 on device tensors: -> (obs, act, rew, obs_tp1, done), each batch_size rows.
        The exploration noise and the reset observations of all steps are drawn up front; `fused` runs the whole
        loop as one kernel (mpg_env_sample), otherwise step by step through the same kernels a gym-style caller uses
        (identical numbers, tests/test_trainer.py)."""
        steps, n = int(self.batch_size / self.num_agent), self.num_agent
        act_dim, dev = self.args.act_dim, self.env.engine.device
        eps = torch.randn(steps, n, act_dim, device=dev, generator=self.generator) if self.explore_sigma is not None else None
        fresh = self.env._draw_on_device(steps).view(steps, n, -1)
        if fused:
            out = self.policy_with_value.engine.env_sample(self.env.state, self.env.obs, fresh, eps,
                                                           self.explore_sigma or 0.0)
            self.obs = self.env.obs
            self.env.done = torch.zeros(n, dtype=torch.int32, device=dev)   # every finished agent has been restarted
        else:
            cols = [[] for _ in range(5)]
            for t in range(steps):
                processed = self.preprocessor.torch_process_obses(self.obs)
                action, _ = self.policy_with_value.compute_action(processed)
                if eps is not None:
                    action = action + self.explore_sigma * eps[t]
                obs = self.obs
                obs_tp1, reward, done, _ = self.env.step(action)
                for c, v in zip(cols, (obs, action, reward, obs_tp1, done.float())):
                    c.append(v)
                self.obs = self.env.reset_done(fresh[t])
            out = [torch.cat(c, 0) for c in cols]
        if not torch.isfinite(out[1]).all():          # one host sync per sample() instead of one per env step
            raise ValueError('nan/inf action (judge_is_nan, utils/misc.py:27-36)')
        self.num_sample += out[0].shape[0]
        self.sample_times += 1
        return out

    def sample(self):
        """Reference format: list of (obs, action, reward, obs_tp1, done) tuples (numpy)."""
        o, a, r, o1, d = [t.cpu().numpy() for t in self.sample_arrays()]
        return [(o[i], a[i], r[i], o1[i], d[i]) for i in range(o.shape[0])]

    def sample_with_count(self):
        batch = self.sample()
        return batch, len(batch)

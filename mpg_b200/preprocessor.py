"""Preprocessor, 'scale' mode only (preprocessor.py:59-74,125-159): obs * obs_scale and
(rew + rew_shift) * rew_scale.  Inside the rollout these two multiplies are fused into the kernels
(the scales travel in mpg_config); this class serves the host-side callers (np_process_rewards in
compute_gradient, mpg_learner.py:411) and keeps set_params/get_params for set_ppc_params."""
import numpy as np


class Preprocessor(object):
    def __init__(self, obs_dim, obs_ptype=None, rew_ptype=None, obs_scale=None, rew_scale=None, rew_shift=None,
                 gamma=0.99, **kwargs):
        if obs_ptype not in ('scale', None) or rew_ptype not in ('scale', None):
            raise NotImplementedError("only the 'scale' preprocessing used by the model-based learners is built "
                                      "('normalize' keeps running statistics outside the hot path)")
        self.obs_ptype, self.rew_ptype = obs_ptype, rew_ptype
        self.obs_scale = np.array(obs_scale, dtype=np.float32) if obs_ptype == 'scale' else np.ones(obs_dim, np.float32)
        self.rew_scale = np.float32(rew_scale) if rew_ptype == 'scale' else np.float32(1.0)
        self.rew_shift = np.float32(rew_shift) if rew_ptype == 'scale' else np.float32(0.0)
        self.gamma = gamma

    def np_process_obses(self, obses):
        return obses * self.obs_scale

    def np_process_rewards(self, rewards):
        return (rewards + self.rew_shift) * self.rew_scale

    def torch_process_obses(self, obses):
        import torch
        return obses * torch.as_tensor(self.obs_scale, device=obses.device)

    def torch_process_rewards(self, rewards):
        return (rewards + float(self.rew_shift)) * float(self.rew_scale)

    def set_params(self, params):  # only the 'normalize' mode has parameters (preprocessor.py:161-165)
        pass

    def get_params(self):
        return {}

    def save_params(self, save_dir):  # preprocessor.py:176-182
        np.save(save_dir + '/ppc_params.npy', self.get_params())

    def load_params(self, load_dir):
        self.set_params(np.load(load_dir + '/ppc_params.npy', allow_pickle=True).item())

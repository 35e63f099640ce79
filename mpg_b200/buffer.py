"""Replay buffers with the reference's interface (buffer.py:21-189) on the GPU.

    buf = PrioritizedReplayBuffer(args, buffer_id); buf.add_batch(batch); samples = buf.replay()
    samples = [obs, act, rew, obs_tp1, done, weights, idxes]   (numpy, like the reference)
    buf.update_priorities(idxes, priorities)

The reference's PrioritizedReplayBuffer does not run as shipped: it reads args.alpha / args.size
which its parsers never define, add_batch inserts priority 0**alpha = 0 so nothing can be sampled, and
update_priorities asserts priority > 0 on signed TD errors.  Here: alpha/beta come from args.replay_alpha /
args.replay_beta, capacity from args.max_buffer_size, new transitions enter at the running max priority (the
behaviour of the reference's own `add(..., weight=None)` branch, buffer.py:132-133), and update_priorities takes
|td| + 1e-6.  Sum/min trees, sampling and gathers run in libmpg_b200 (csrc/replay.cuh).

`replay_device()` / `add_arrays()` are the zero-copy variants for a learner living on the same GPU."""
import ctypes
import logging

import numpy as np
import torch

from . import _lib

logger = logging.getLogger(__name__)


class _DeviceReplay(object):
    def __init__(self, capacity, obs_dim, act_dim, alpha, beta, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('mpg_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.mpg_replay_create(int(capacity), obs_dim, act_dim, float(alpha), float(beta), ctypes.byref(self.h))
        if rc != 0:
            raise RuntimeError('mpg_replay_create failed: ' + self.lib.mpg_replay_last_error(None).decode())
        self.generator = torch.Generator(device=self.device)

    def __del__(self):
        try:
            if self.h.value:
                self.lib.mpg_replay_destroy(self.h)
                self.h = ctypes.c_void_p()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f'libmpg_b200 replay error {rc}: ' + self.lib.mpg_replay_last_error(self.h).decode())

    def _dev(self, x, dtype=torch.float32):
        if isinstance(x, torch.Tensor):
            return x.to(self.device, dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32 if dtype == torch.float32 else np.int32)).to(self.device)

    @property
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def __len__(self):
        return self.lib.mpg_replay_size(self.h)

    def add(self, obs, act, rew, obs_tp1, done, priorities=None):
        obs, act, rew, obs_tp1, done = (self._dev(x) for x in (obs, act, rew, obs_tp1, done))
        pr = None if priorities is None else self._dev(priorities)
        self._check(self.lib.mpg_replay_add(self.h, obs.shape[0], self._p(obs), self._p(act), self._p(rew), self._p(obs_tp1),
                                            self._p(done), self._p(pr), self.stream))

    def sample(self, n, u=None):
        u = torch.rand(n, device=self.device, generator=self.generator) if u is None else self._dev(u)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        idx = torch.empty(n, dtype=torch.int32, device=self.device)
        w, obs, act, rew, obs1, done = f(n), f(n, self.obs_dim), f(n, self.act_dim), f(n), f(n, self.obs_dim), f(n)
        self._check(self.lib.mpg_replay_sample(self.h, n, self._p(u), self._p(idx), self._p(w), self._p(obs), self._p(act),
                                               self._p(rew), self._p(obs1), self._p(done), self.stream))
        return obs, act, rew, obs1, done, w, idx

    def update_priorities(self, idx, priorities):
        idx, pr = self._dev(idx, torch.int32), self._dev(priorities)
        self._check(self.lib.mpg_replay_update_priorities(self.h, idx.shape[0], self._p(idx), self._p(pr), self.stream))

    def tree_stats(self):
        s, m, mx = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.mpg_replay_tree_stats(self.h, ctypes.byref(s), ctypes.byref(m), ctypes.byref(mx), self.stream))
        return s.value, m.value, mx.value


class ReplayBuffer(object):
    """Uniform replay (buffer.py:21-91): every stored transition carries priority 1, so the proportional
    sampler is the uniform one; replay() returns [obs, act, rew, obs_tp1, done, idxes]."""
    prioritized = False

    def __init__(self, args, buffer_id):
        self.args, self.buffer_id = args, buffer_id
        self._maxsize = self.args.max_buffer_size
        self.replay_starts = self.args.replay_starts
        self.replay_batch_size = self.args.replay_batch_size
        self.stats, self.replay_times = {}, 0
        alpha = getattr(args, 'replay_alpha', 0.6) if self.prioritized else 1.0
        beta = getattr(args, 'replay_beta', 0.4) if self.prioritized else 0.0
        self.dev = _DeviceReplay(self._maxsize, args.obs_dim, args.act_dim, alpha, beta, getattr(args, 'device', None))
        logger.info('Buffer initialized')

    def get_stats(self):
        self.stats.update(dict(storage=len(self)))
        return self.stats

    def __len__(self):
        return len(self.dev)

    def add_arrays(self, obs, act, rew, obs_tp1, done, priorities=None):
        one = None if self.prioritized else torch.ones(len(rew), device=self.dev.device)
        self.dev.add(obs, act, rew, obs_tp1, done, priorities if self.prioritized else one)

    def add_batch(self, batch):
        """batch: list of (obs_t, action, reward, obs_tp1, done) transitions (worker.py sample format)."""
        cols = list(zip(*batch))
        self.add_arrays(np.stack(cols[0]), np.stack(cols[1]), np.asarray(cols[2], np.float32), np.stack(cols[3]),
                        np.asarray(cols[4], np.float32))

    def replay_device(self, batch_size=None):
        return self.dev.sample(batch_size or self.replay_batch_size)

    def replay(self):
        if len(self) < self.replay_starts:
            return None
        if self.buffer_id == 1 and self.replay_times % self.args.buffer_log_interval == 0:
            logger.info('Buffer info: {}'.format(self.get_stats()))
        self.replay_times += 1
        obs, act, rew, obs1, done, w, idx = self.replay_device()
        out = [t.cpu().numpy() for t in (obs, act, rew, obs1, done)]
        if self.prioritized:
            out.append(w.cpu().numpy())
        return out + [idx.cpu().numpy()]


class PrioritizedReplayBuffer(ReplayBuffer):
    """buffer.py:94-189 with the fixes listed in the module docstring."""
    prioritized = True

    def update_priorities(self, idxes, priorities):
        pr = priorities if isinstance(priorities, torch.Tensor) else np.asarray(priorities, np.float32)
        pr = abs(pr) + 1e-6   # TD errors are signed (learners' get_info_for_buffer); the tree needs priority > 0
        self.dev.update_priorities(idxes, pr)

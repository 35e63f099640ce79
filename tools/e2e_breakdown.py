"""Where one NADPLearner.compute_gradient (B = 65536, host numpy in / out) spends its time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpg_b200 import synthetic
from mpg_b200.config import default_args
from mpg_b200.learners import NADPLearner
from mpg_b200.policy import PolicyWithQs
import bench
B = 65536
args = default_args('NADP', 'PathTracking-v0', replay_batch_size=B)
L = NADPLearner(PolicyWithQs, args)
L.set_weights(synthetic.make_policy_with_qs_weights(0, 6, 2, 256, double_q=False))
L.engine.set_backend(1)   # (the default where the tensor-core path covers the configuration)
batch = bench.make_inputs("PathTracking-v0", B)
for _ in range(3): L.compute_gradient(batch, None, None, 0)
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('compute_gradient        %.3f ms' % t(lambda: L.compute_gradient(batch, None, None, 0)))
print('  upload (H2D, pinned)   %.3f ms' % t(lambda: L._upload_batch(batch)))
o, a = L._dev['batch_obs'], L._dev['batch_actions']
print('  q_forward_and_backward %.3f ms' % t(lambda: L.q_forward_and_backward(o, a)))
print('    rollout for q target %.3f ms' % t(lambda: L.model_rollout_for_q_estimation(o, a)))
print('  policy_fwd_and_bwd     %.3f ms' % t(lambda: L.policy_forward_and_backward(o)))
g = torch.zeros(136965 + 10, device='cuda')
print('  clip x2 + D2H          %.3f ms' % t(lambda: (L.engine.clip_global_norm(g[:68353], 3.0), L.engine.clip_global_norm(g[68353:136965], 3.0), L._to_host(g))))

# the same for the reference's flagship learner: MPG-v2 defaults (clipped double-Q targets, first-action policy gradient)
from mpg_b200.learners import MPGLearner
args2 = default_args('MPG-v2', 'PathTracking-v0', replay_batch_size=B)
L2 = MPGLearner(PolicyWithQs, args2)
L2.set_weights(synthetic.make_policy_with_qs_weights(0, 6, 2, 256, double_q=True))
for _ in range(3): L2.compute_gradient(batch, None, None, 100)
ms = t(lambda: L2.compute_gradient(batch, None, None, 100))
print('MPGLearner (MPG-v2) compute_gradient %.3f ms = %.1f M state-steps/s end to end, %.0f updates/s' % (ms, B * 25 / ms / 1e3, 1e3 / ms))
d = L2._dev
print('  policy_fwd_and_bwd     %.3f ms' % t(lambda: L2.policy_forward_and_backward(d['batch_obs'], 100)))
print('  q_fwd_and_bwd (2 nets) %.3f ms' % t(lambda: L2.q_forward_and_backward(d['batch_obs'], d['batch_actions'], d['batch_targets'])))

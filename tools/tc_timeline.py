"""clock64 timeline of ONE backward step (t = horizon-1) and one forward step (t = 2) of CTA 0 of tc_rollout_kernel
(needs a library built with the profile hooks: make VARIANT=-DMPG_DEBUG_PROBES)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from mpg_b200 import synthetic
from mpg_b200.config import default_args
from mpg_b200.policy import PolicyWithQs
B = 65536
args = default_args('NADP', 'PathTracking-v0', replay_batch_size=B)
pol = PolicyWithQs(**vars(args)); pol.set_weights(synthetic.make_policy_with_qs_weights(1, 6, 2, 256, double_q=False))
e = pol.engine; e.set_backend(1)
obs = e.dev(synthetic.make_obs(np.random.default_rng(2), 'PathTracking-v0', B))
buf = torch.zeros(64, dtype=torch.int64, device='cuda')
for full in (1, 0):
    e.lib.mpg_set_profile_buffer(e.h, ctypes.c_void_p(buf.data_ptr()))
    for _ in range(3): e.policy_grad(obs, [0, 25], [0.0, 1.0], full_bptt=bool(full), use_philox=True)
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    ep = t[:12] - t[0]; mm = t[32:37] - t[0]
    names = {0: 'start', 1: 'p image, load issued', 2: 'd3 ready', 3: 'h2 image landed', 6: 'D3 half / Ed2 start', 7: 'Ed2 done',
             8: 'g_h1 ready', 9: 'Ed1 done', 10: 'g_p ready', 11: 'step end'}
    print('full_bptt', full)
    for i, n in names.items(): print(f'  epi {n:22s} {ep[i]:8d}')
    print('  mma: dx-start', mm[3], 'dx-issued', mm[4])
    fw = t[16:23] - t[16]
    print('  forward step t=2: start 0, p image published', fw[1], 'E1 done', fw[2], 'z2 ready', fw[3], 'E2 done', fw[4], 'zpre', fw[5], 'next step', fw[6])

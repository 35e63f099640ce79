"""clock64 timeline of ONE backward step (t = horizon-1) and one forward step (t = 2) of CTA 0 of tc_rollout_kernel:
epilogue thread 0, row thread 0 and the mma thread."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from mpg_b200 import synthetic
from mpg_b200.config import default_args
from mpg_b200.policy import PolicyWithQs
B = 65536
args = default_args('NADP', 'PathTracking-v0', replay_batch_size=B)
pol = PolicyWithQs(debug_lib=True, **vars(args)); pol.set_weights(synthetic.make_policy_with_qs_weights(1, 6, 2, 256, double_q=False))
e = pol.engine; e.set_backend(1)
obs = e.dev(synthetic.make_obs(np.random.default_rng(2), 'PathTracking-v0', B))
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
for full in (1, 0):
    e.lib.mpg_set_profile_buffer(e.h, ctypes.c_void_p(buf.data_ptr()))
    for _ in range(3): e.policy_grad(obs, [0, 25], [0.0, 1.0], full_bptt=bool(full), use_philox=True)
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    t0 = t[0]
    ep = t[:24] - t0; mm = t[32:56] - t0; rw = t[64:88] - t0
    print('full_bptt', full)
    for i, n in {0: 'start', 12: 'acc_wait done', 1: 'load issued', 2: 'waiting for h2', 3: 'h2 image landed', 6: 'Ed2 start (D3 half, d3s)',
                 7: 'Ed2 done', 8: 'g_h1 ready', 9: 'Ed1 done', 10: 'g_p ready', 11: 'step end'}.items():
        print(f'  epi {n:26s} {ep[i]:8d}')
    for i, n in {0: 'start', 4: 'checkpoint loads issued', 5: 'env adjoint start', 12: 'acc_wait done', 13: 'p image written', 14: 'env adjoint done', 2: 'd3 published', 10: 'g_p ready', 11: 'lambda done'}.items():
        print(f'  row {n:26s} {rw[i]:8d}')
    print('  mma: d3-start', mm[5], 'd3-issued', mm[6], 'dx-start', mm[3], 'dx-issued', mm[4])
    fw = t[16:23] - t[16]; fr = t[80:87] - t[16]
    print('  forward step t=2 (epi): start 0, E1 start', fw[1], 'E1 done', fw[2], 'z2 ready', fw[3], 'E2 done', fw[4], 'next step', fw[6])
    print('  forward step t=2 (mma): enter', t[32 + 16] - t[16], 'all UMMAs issued', t[32 + 17] - t[16])
    g = t[96:124] - t[16]
    print('  forward step t=2 (mma): p image seen', g[21], 'W1 image seen', g[22], 'chunks 0,1 issued', g[23], 'chunk 2 issued', g[24], 'pad consumed', g[25])
    print('  forward step t=2 (mma): layer-2 GEMM starts', g[20], '; per k-block: A block seen / stages seen:', ' | '.join('%d / %s' % (g[kb * 5], ' '.join(str(x) for x in g[kb * 5 + 1: kb * 5 + 5])) for kb in range(4)))
    print('  forward step t=2 (row): start', fr[0], 'zpre', fr[5], 'next step', fr[6])

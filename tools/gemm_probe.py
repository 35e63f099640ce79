"""Per-GEMM time of the streamed 128x256x256 split-bf16 contraction (48 UMMAs + 256 KB of weight image through the
ring) in isolation: 1 CTA (shared-memory / tensor-pipe bound) versus one CTA per SM (adds L2 contention)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from mpg_b200.engine import Engine
    e = Engine(env_id='PathTracking-v0', obs_dim=6, act_dim=2, obs_scale=None, rew_scale=1.0, rew_shift=0.0, gamma=1.0, max_rows=128, max_horizon=1, debug_lib=True)
    X = e.dev(np.zeros((128, 256), np.float32)); W = e.dev(np.zeros((256, 256), np.float32))
    e.tc_selftest(0, X, W)            # packs an image into the scratch buffer
    reps = 4000
    if len(sys.argv) > 2:      # CTA-pair probe (cta_group::2): result check, then timing
        g = torch.Generator().manual_seed(0)
        X2 = torch.randn(256, 256, generator=g); W2 = torch.randn(256, 256, generator=g) / 16
        Z = e.tc_selftest(5, e.dev(X2), e.dev(W2), 1)
        torch.cuda.synchronize()
        ref = X2.double() @ W2.double().T
        print('pair probe rel-L2 error %.3e' % float((Z.cpu().double() - ref).norm() / ref.norm()))
        for _ in range(2): e.tc_selftest(5, e.dev(X2), e.dev(W2), repeats=reps)
        torch.cuda.synchronize()
        xa, wa = e.dev(X2), e.dev(W2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e.tc_selftest(5, xa, wa, repeats=reps); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print('grid %s, CTA pair (two 128-row tiles per GEMM): %.2f us per pair-GEMM = %.0f cycles @1.965 GHz' % (
            sys.argv[1], ms * 1e3 / reps, ms * 1e-3 / reps * 1.965e9))
        sys.exit(0)
    for kind, name in ((3, 'streamed weights'), (4, 'resident operands')):
        for _ in range(2): e.tc_selftest(kind, X, W, repeats=reps)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e.tc_selftest(kind, X, W, repeats=reps); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print('grid %s, %s: %.2f us per GEMM = %.0f cycles @1.965 GHz' % (sys.argv[1], name, ms * 1e3 / reps, ms * 1e-3 / reps * 1.965e9))
else:
    for g in (1, 148):
        subprocess.run([sys.executable, __file__, str(g)], env=dict(os.environ, MPG_SELFTEST_GRID=str(g)))
    for g in (2, 148):
        subprocess.run([sys.executable, __file__, str(g), 'pair'], env=dict(os.environ, MPG_SELFTEST_GRID=str(g)), timeout=120)

"""cProfile of MPGLearner.compute_gradient at the reference's batch size (256) on device-resident replay samples."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpg_b200.trainer import Trainer
from mpg_b200.config import default_args
alg = sys.argv[1] if len(sys.argv) > 1 else 'MPG-v2'
args = default_args(alg, 'PathTracking-v0', replay_batch_size=256, batch_size=512, num_agent=8, explore_sigma=0.1,
                    max_buffer_size=100000, replay_starts=2048, buffer_log_interval=10 ** 9, num_eval_agent=64,
                    num_eval_episode=1, fixed_steps=60, eval_interval=10 ** 9, log_interval=10 ** 9, max_iter=400, log_dir=None)
tr = Trainer(args)
tr.train(50)
L, rb = tr.learner, tr.buffer
samples = rb.replay_device()
def one():
    return L.compute_gradient(samples[:5], rb, samples[-1], 100)
for _ in range(20): one()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): one()
torch.cuda.synchronize(); print('compute_gradient %.3f ms' % ((time.perf_counter() - t0) / 200 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): one()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20): one()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=70))

"""Target of the ncu captures: NADPLearner.compute_gradient at the bench workload (B = 65536, n = 25), `reps` updates.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 2
    ncu --set full --clock-control none --import-source on -k regex:tc_rollout_kernel -s 3 -c 3 -o gpurun_out/prof python tools/ncu_target.py 2"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpg_b200 import synthetic
from mpg_b200.config import default_args
from mpg_b200.learners import NADPLearner
from mpg_b200.policy import PolicyWithQs
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
args = default_args('NADP', 'PathTracking-v0', replay_batch_size=rows)
learner = NADPLearner(PolicyWithQs, args)
learner.set_weights(synthetic.make_policy_with_qs_weights(0, args.obs_dim, args.act_dim, 256, double_q=False))
learner.engine.set_backend(1)
batch = bench.make_inputs('PathTracking-v0', rows)
for it in range(reps):
    learner.compute_gradient(batch, None, None, it)
torch.cuda.synchronize()
print('done', learner.get_stats()['pg_time'], learner.get_stats()['q_timer'])

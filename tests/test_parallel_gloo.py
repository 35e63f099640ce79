"""CPU, world_size 2, gloo: the N>1 host path of the learner.  Each rank computes, with the oracle, the
gradient SUM of its contiguous shard scaled by 1/B_global (what mpg_policy_grad / mpg_q_grad return with
global_rows set) on the slice of the globally-keyed Philox noise it owns; one all-reduce of the flat
[q grad | policy grad | scalar sums] buffer must reproduce the full-batch oracle gradient on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    from mpg_b200 import parallel, synthetic
    from mpg_b200.config import default_args
    from oracle import mpg_oracle as O
    from tests.util import make_batch

    assert parallel.dist_info() == (world, rank)
    args = default_args('NADP', 'PathTracking-v0', replay_batch_size=B, value_num_hidden_units=32,
                        policy_num_hidden_units=32, num_rollout_list_for_policy_update=[n],
                        num_rollout_list_for_q_estimation=[n], gradient_clip_norm=1e9)
    w = synthetic.make_policy_with_qs_weights(5, args.obs_dim, args.act_dim, 32, double_q=False)
    batch = make_batch(6, args.env_id, B, 0)
    local = B // world
    g_rows, off = parallel.shard_rows(local, world, rank)
    assert (g_rows, off) == (B, local * rank)
    # noise keyed by (seed, global row, step): the shard's stream is a slice of the global one
    nq_full, np_full = synthetic.philox_normal(7, B, n), synthetic.philox_normal(8, B, n)
    nq = synthetic.philox_normal(7, local, n, global_rows=B, row_offset=off)
    npol = synthetic.philox_normal(8, local, n, global_rows=B, row_offset=off)
    assert np.array_equal(nq, nq_full[:, off:off + local]) and np.array_equal(npol, np_full[:, off:off + local])
    shard = [b[off:off + local] for b in batch]
    import copy
    largs = copy.copy(args)
    largs.replay_batch_size = local
    r = O.nadp_compute_gradient(largs, w, shard, nq, npol, torch.float64)
    scale = local / B   # oracle means over the shard -> sums scaled by 1/B_global
    flat = torch.tensor(np.concatenate([r['q_grad'] * scale, r['policy_grad'] * scale,
                                        [r['q_loss'] * local, -r['policy_loss'] * local]]))
    parallel.allreduce_flat(flat, world)
    full = O.nadp_compute_gradient(args, w, batch, nq_full, np_full, torch.float64)
    ref = np.concatenate([full['q_grad'], full['policy_grad']])
    got = flat.numpy()
    err = np.linalg.norm(got[:-2] - ref) / np.linalg.norm(ref)
    grads = parallel.split_flat(got[:-2].astype(np.float32), args.obs_dim, args.act_dim, ['q', 'pi'], hidden=32)
    ok = (err < 1e-10 and abs(got[-2] / B - full['q_loss']) < 1e-10 * abs(full['q_loss'])
          and abs(-got[-1] / B - full['policy_loss']) < 1e-10 * abs(full['policy_loss'])
          and len(grads) == 12 and grads[0].shape == (8, 32) and grads[-1].shape == (4,))
    open(os.path.join(out_dir, f'rank{rank}.txt'), 'w').write(f'{int(ok)} {err:.3e}')
    dist.destroy_process_group()


def test_sharded_gradient_allreduce_world2(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, 24, 6, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok, err = open(tmp_path / f'rank{r}.txt').read().split()
        assert ok == '1', (r, err)

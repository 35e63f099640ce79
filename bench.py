#!/usr/bin/env python3
"""bench.py -- model state-steps/s (fwd+bwd, n = 25) of the MPG model-based learner hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--backend auto|ffma|tc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): PathTrackingModel, NADP (pure n-step ADP gradient, full BPTT),
n = 25, B = 65536 rows per GPU, H = 256, synthetic seeded states/weights, in-kernel Philox noise.
  value : B*n*N / (device time of ONE policy forward+backward rollout, mpg_policy_grad, inputs resident
          in HBM; for N > 1 the NCCL all-reduce of the flat policy gradient is inside the step)
  e2e   : the same unit through the reference-facing call NADPLearner.compute_gradient(batch, rb, idx, it)
          with HOST numpy buffers: H2D of the replay batch, Q-target rollout, Q gradient, policy
          rollout fwd+bwd, clip, D2H of the 12 gradient arrays + stats are all inside the timed region.
  --impl reference : the restated reference learner (oracle, PyTorch CPU fp32 -- TensorFlow is not
          installable in this image) timed on the host cores for the same compute_gradient.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter)
# is sent to stderr; the JSON line goes to a private duplicate of the original stdout.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_JSON_FD, (json.dumps(line) + '\n').encode())


from mpg_b200 import synthetic  # noqa: E402
from mpg_b200.config import default_args  # noqa: E402

ENV_ID, N_STEPS, ROWS_PER_GPU, HID = 'PathTracking-v0', 25, 65536, 256
F_PI = 2 * (6 * HID + HID * HID + HID * 4)        # 136,192 FLOP per policy forward row
F_Q = 2 * (8 * HID + HID * HID + HID)             # 135,680
# algorithmic FLOP per trajectory, full BPTT: (n+1)F_pi fwd + n*2F_pi bwd + (2F_pi - 2*d_o*H) + 3F_Q
FLOP_PER_TRAJ = (N_STEPS + 1) * F_PI + N_STEPS * 2 * F_PI + (2 * F_PI - 2 * 6 * HID) + 3 * F_Q
FLOP_PER_STATE_STEP = FLOP_PER_TRAJ / N_STEPS     # 441,078


def make_inputs(rows, seed=1234):
    rng = np.random.default_rng(seed)
    obs = synthetic.make_obs(rng, ENV_ID, rows)
    act = np.clip(rng.normal(0.0, 0.5, (rows, 2)), -1, 1).astype(np.float32)
    rew = (-np.abs(rng.standard_normal(rows))).astype(np.float32)
    obs_tp1 = synthetic.make_obs(rng, ENV_ID, rows)
    done = np.zeros(rows, np.float32)
    return [obs, act, rew, obs_tp1, done]


def cpu_reference_update(args, weights, batch, threads, repeats):
    """Time the restated reference learner (oracle, fp32) for one compute_gradient. Returns seconds (median)."""
    from oracle import mpg_oracle as O
    torch.set_num_threads(threads)
    rows = batch[0].shape[0]
    rng = np.random.default_rng(7)
    nq, npol = synthetic.make_noise(rng, N_STEPS, rows), synthetic.make_noise(rng, N_STEPS, rows)
    times = []
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        O.nadp_compute_gradient(args, weights, batch, nq, npol, torch.float32)
        if i:
            times.append(time.perf_counter() - t0)
    return float(np.median(times))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '100', '-i', str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi needs up to a second to start on an 8-GPU box: do not begin before it is sampling."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.05)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       power_w_max=float(max(power)), samples=len(sm))
        return out


def workload_name(rows):
    return ('PathTrackingModel NADP (pure n-step ADP gradient, full BPTT), n=25, H=256, '
            f'B={rows} rows per GPU (BASELINE.json configs[1])')


def run_reference(opts, rank):
    """Reference arm: the restated TF2 learner on the host cores, bounded sample of the workload."""
    if rank != 0:
        return
    sample_rows = 4096
    args = default_args('NADP', ENV_ID, replay_batch_size=sample_rows)
    weights = synthetic.make_policy_with_qs_weights(0, args.obs_dim, args.act_dim, HID, double_q=False)
    batch = make_inputs(sample_rows)
    threads = os.cpu_count() or 1
    from oracle import mpg_oracle as O
    torch.set_num_threads(threads)
    rng = np.random.default_rng(7)
    nq, npol = synthetic.make_noise(rng, N_STEPS, sample_rows), synthetic.make_noise(rng, N_STEPS, sample_rows)
    for _ in range(opts.warmup):
        O.nadp_compute_gradient(args, weights, batch, nq, npol, torch.float32)
    t0 = time.perf_counter()
    for _ in range(opts.steps):
        O.nadp_compute_gradient(args, weights, batch, nq, npol, torch.float32)
    dt = (time.perf_counter() - t0) / opts.steps
    value = sample_rows * N_STEPS / dt
    sample = (f'{sample_rows} of {ROWS_PER_GPU} rows per step, full NADP compute_gradient (Q-target rollout + Q grad + '
              f'policy rollout fwd+bwd + clip), PyTorch-CPU fp32 restatement of the TF2 learner, {threads} threads')
    emit({
        'impl': 'reference', 'metric': 'model state-steps/s (fwd+bwd, n=25)', 'value': value,
        'unit': 'state-steps/s', 'n_gpus': opts.gpus, 'steps': opts.steps, 'warmup': opts.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'updates_per_s': 1.0 / dt,
        'config': {'workload': workload_name(ROWS_PER_GPU), 'global_batch': ROWS_PER_GPU * opts.gpus, 'horizon': N_STEPS,
                   'sample': f'{sample_rows} of {ROWS_PER_GPU} rows per step on the host cores'},
        'cpu_baseline': {'value': value, 'unit': 'state-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'state-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--backend', default='auto', choices=['auto', 'ffma', 'tc'])
    ap.add_argument('--rows', type=int, default=ROWS_PER_GPU)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    opts = ap.parse_args()
    opts.warmup = max(opts.warmup, 3) if opts.impl == 'ours' else opts.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if opts.impl == 'reference':
        run_reference(opts, rank)
        return
    if world != opts.gpus:
        if opts.gpus != 1:
            raise SystemExit(f'--gpus {opts.gpus} needs torchrun with {opts.gpus} processes (WORLD_SIZE={world})')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    from mpg_b200 import _lib
    from mpg_b200.learners import NADPLearner
    from mpg_b200.policy import PolicyWithQs

    rows = opts.rows
    args = default_args('NADP', ENV_ID, replay_batch_size=rows)
    weights = synthetic.make_policy_with_qs_weights(0, args.obs_dim, args.act_dim, HID, double_q=False)
    learner = NADPLearner(PolicyWithQs, args)
    learner.set_weights(weights)
    e = learner.engine
    backend = 'ffma'
    if opts.backend in ('auto', 'tc') and e.tc_available():
        e.set_backend(1)
        backend = 'tc'
    elif opts.backend == 'tc':
        raise SystemExit('tensor-core backend unavailable for this configuration')
    else:
        e.set_backend(0)
    batch = make_inputs(rows, seed=1234 + rank)   # every rank owns different rows of the global batch
    obs_dev = e.dev(batch[0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2
    P = e.param_count(_lib.NET_POLICY)
    global_rows, row_offset = rows * world, rows * rank

    def device_step():
        g, _ = e.policy_grad(obs_dev, [N_STEPS], [1.0], full_bptt=True, q_net=_lib.NET_Q1, use_philox=True,
                             noise_seed=7, global_rows=global_rows, row_offset=row_offset, want_returns=False)
        if world > 1:
            dist.all_reduce(g)
        return g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident metric ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None   # runs through warm-up, timed region and e2e loop
    if sampler:
        sampler.wait_first_sample()
    for _ in range(opts.warmup):
        device_step()
    e.set_timing(True)
    barrier()
    l0 = e.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(opts.steps)]
    kernel_ms = []
    for a, b in ev:
        flush.zero_()                       # evict L2 between timed iterations
        a.record()
        device_step()
        b.record()
        kernel_ms.append(None)
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = e.launch_count - l0
    # dominant-kernel duration (events recorded inside the library right around the rollout kernel)
    for _ in range(3):
        flush.zero_()
        device_step()
        torch.cuda.synchronize()
        kernel_ms.append(e.kernel_ms())
    kernel_ms = [k for k in kernel_ms if k is not None and k > 0]
    e.set_timing(False)
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / opts.steps
    value = rows * world * N_STEPS / (ms_per_step * 1e-3)

    # ---------------- end to end through the learner API with host buffers ----------------
    for _ in range(2):
        learner.compute_gradient(batch, None, None, 0)
    e2e_steps = max(3, min(opts.steps, 10))
    block_s = []
    for _ in range(2):            # two blocks of K iterations, the faster one is reported: a one-off host stall
        barrier()                 # (page faults, a noisy neighbour on the box) must not decide the wall-clock number
        t0 = time.perf_counter()
        for it in range(e2e_steps):
            grads = learner.compute_gradient(batch, None, None, it)   # returns host numpy arrays (D2H inside)
        torch.cuda.synchronize()
        block_s.append((time.perf_counter() - t0) / e2e_steps)
    e2e_s = torch.tensor([min(block_s)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    d2h = int(sum(g.nbytes for g in grads)) + 4 * 8
    e2e_value = rows * world * N_STEPS / e2e_s
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    k_ms = float(np.mean(kernel_ms)) if kernel_ms else ms_per_step
    achieved_tf = rows * N_STEPS * FLOP_PER_STATE_STEP / (k_ms * 1e-3) / 1e12
    peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
    ffma_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel, one ncu --set full capture
    # tc: the rollout runs as two launches of the same kernel (full waves + tail wave); their bytes are summed
    prof = os.path.join(ROOT, 'profiles', 'r1c_tc_rollout_kernel_ncu_full.csv' if backend == 'tc' else
                        'r1_ffma_rollout_kernel_ncu_full.csv')
    try:
        mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tot = 0.0
        for line in open(prof):
            c = line.strip().split(',')
            if c[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                tot += float(c[2]) * mult[c[1]]
        traffic = tot * (rows / float(ROWS_PER_GPU)) if tot > 0 else None
    except Exception:
        traffic = None
    roofline = {
        'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf,
        'traffic': traffic,
        'traffic_note': 'bytes per launch from profiles/ (ncu --set full of the same kernel at B=65536); algorithmic '
                        'bytes are 56 B/state-step = 92 MB; the tc path adds the dW operand store (2.1 KB/state-step: h1 and '
                        'delta2 images for dW2); tc kernel_ms spans both launches of the rollout (full waves + tail wave)',
        'kernel': 'rollout_kernel<PathTracking,BWD> (fused forward rollout + BPTT, %s backend)' % backend,
        'kernel_ms': k_ms, 'algorithmic_flop_per_state_step': FLOP_PER_STATE_STEP,
        'peak_source': ('MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1.4 PFLOP/s (of fallback)'),
        'fp32_ffma_peak_tflops_at_run_clock': ffma_peak, 'frac_of_fp32_ffma_peak': achieved_tf / ffma_peak,
        'hbm_bytes_per_state_step_algorithmic': 56,
    }
    cpu_baseline = None
    if world == 1 and not opts.no_cpu_baseline:
        cpu_args = default_args('NADP', ENV_ID, replay_batch_size=4096)
        cpu_batch = make_inputs(4096)
        threads = os.cpu_count() or 1
        t_all = cpu_reference_update(cpu_args, weights, cpu_batch, threads, 3)
        b256 = [b[:256] for b in cpu_batch]
        t_one = cpu_reference_update(default_args('NADP', ENV_ID, replay_batch_size=256), weights, b256, 1, 3)
        cpu_baseline = {
            'value': 4096 * N_STEPS / t_all, 'unit': 'state-steps/s', 'cores': threads, 'kind': 'port',
            'sample': ('4096-row sample of the 65536-row workload, full NADP compute_gradient, PyTorch-CPU fp32 '
                       'restatement of the TF2 learner (TensorFlow not installable here), median of 3'),
            'single_thread_b256_value': 256 * N_STEPS / t_one,
            'single_thread_b256_note': 'reference default: 1 intra/inter-op thread per learner, batch 256',
        }
    line = {
        'metric': 'model state-steps/s (fwd+bwd, n=25)', 'value': value, 'unit': 'state-steps/s', 'n_gpus': world,
        'steps': opts.steps, 'warmup': opts.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(rows),
                   'global_batch': rows * world, 'horizon': N_STEPS, 'backend': backend,
                   'noise': 'in-kernel Philox4x32-10 keyed (seed, global row, step)',
                   'cache': 'L2 flushed between timed iterations (256 MiB memset)',
                   'step': 'one policy forward+backward rollout (mpg_policy_grad)' + (
                       ' + NCCL all-reduce of the flat policy gradient' if world > 1 else ''),
                   'e2e_step': 'NADPLearner.compute_gradient with host numpy buffers (adds Q-target rollout, Q gradient, clip); '
                               'wall clock, faster of two blocks of %d updates' % e2e_steps},
        'updates_per_s': 1.0 / e2e_s,
        'e2e': {'value': e2e_value, 'unit': 'state-steps/s', 'h2d_bytes_per_step': int(learner.h2d_bytes),
                'd2h_bytes_per_step': d2h, 'ms_per_update': e2e_s * 1e3},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'cpu_baseline': cpu_baseline,
        'target_state_steps_per_s_per_gpu': 1e8,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

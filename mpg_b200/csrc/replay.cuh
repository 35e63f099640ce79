// GPU prioritized replay: sum / min segment trees (utils/segment_tree.py:13-151),
// proportional sampling with importance weights and priority updates (buffer.py:94-189).
// Trees are fp64 arrays of 2*capacity nodes (node 1 = root, leaves at [capacity, 2*capacity)), rebuilt level by
// level after leaf updates -- capacity 2^19 is 8 MB per tree and L2 resident; sampling is one thread per
// draw walking log2(capacity) nodes exactly like SumSegmentTree.find_prefixsum_idx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mpg {

__global__ void replay_write_kernel(int n, int capacity, int maxsize, int next_idx, int obs_dim, int act_dim,
                                    const float* __restrict__ obs, const float* __restrict__ act,
                                    const float* __restrict__ rew, const float* __restrict__ obs_tp1,
                                    const float* __restrict__ done, const float* __restrict__ prio,
                                    const double* __restrict__ max_prio, double alpha,
                                    float* s_obs, float* s_act, float* s_rew, float* s_obs1, float* s_done,
                                    double* sum_tree, double* min_tree) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = (next_idx + i) % maxsize;   // ring of maxsize transitions; `capacity` only places the tree leaves
  for (int k = 0; k < obs_dim; ++k) {
    s_obs[(size_t)slot * obs_dim + k] = obs[(size_t)i * obs_dim + k];
    s_obs1[(size_t)slot * obs_dim + k] = obs_tp1[(size_t)i * obs_dim + k];
  }
  for (int k = 0; k < act_dim; ++k) s_act[(size_t)slot * act_dim + k] = act[(size_t)i * act_dim + k];
  s_rew[slot] = rew[i];
  s_done[slot] = done ? done[i] : 0.f;
  const double w = prio ? (double)prio[i] : *max_prio;      // weight None -> max priority (buffer.py:132-133)
  const double p = pow(w, alpha);
  sum_tree[capacity + slot] = p;
  min_tree[capacity + slot] = p;
}

// last occurrence of a duplicated index wins, like the reference's sequential loop
__global__ void replay_owner_kernel(int n, const int32_t* __restrict__ idx, int* owner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(&owner[idx[i]], i);
}
__global__ void replay_set_prio_kernel(int n, int capacity, const int32_t* __restrict__ idx,
                                       const float* __restrict__ prio, double alpha, int* owner, double* sum_tree,
                                       double* min_tree) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = idx[i];
  if (owner[id] == i) {
    const double p = pow((double)prio[i], alpha);
    sum_tree[capacity + id] = p;
    min_tree[capacity + id] = p;
  }
}
__global__ void replay_owner_reset_kernel(int n, const int32_t* __restrict__ idx, int* owner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) owner[idx[i]] = -1;
}
__global__ void replay_max_prio_kernel(int n, const float* __restrict__ prio, double* max_prio) {
  __shared__ float red[256];
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, prio[i]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0 && (double)red[0] > *max_prio) *max_prio = (double)red[0];
}
// one tree level: nodes [first, 2*first)
__global__ void replay_level_kernel(int first, double* sum_tree, double* min_tree) {
  const int node = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= 2 * first) return;
  sum_tree[node] = sum_tree[2 * node] + sum_tree[2 * node + 1];
  min_tree[node] = fmin(min_tree[2 * node], min_tree[2 * node + 1]);
}
// the top levels (first <= 512) in one block
__global__ void replay_top_kernel(int first, double* sum_tree, double* min_tree) {
  for (int f = first; f >= 1; f >>= 1) {
    for (int node = f + threadIdx.x; node < 2 * f; node += blockDim.x) {
      sum_tree[node] = sum_tree[2 * node] + sum_tree[2 * node + 1];
      min_tree[node] = fmin(min_tree[2 * node], min_tree[2 * node + 1]);
    }
    __syncthreads();
  }
}

__global__ void replay_sample_kernel(int n, int capacity, int size, int obs_dim, int act_dim, double beta,
                                     const float* __restrict__ u, const double* __restrict__ sum_tree,
                                     const double* __restrict__ min_tree, const float* __restrict__ s_obs,
                                     const float* __restrict__ s_act, const float* __restrict__ s_rew,
                                     const float* __restrict__ s_obs1, const float* __restrict__ s_done,
                                     int32_t* __restrict__ idx_out, float* __restrict__ w_out, float* __restrict__ obs_out,
                                     float* __restrict__ act_out, float* __restrict__ rew_out,
                                     float* __restrict__ obs1_out, float* __restrict__ done_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double total = sum_tree[1];
  double prefix = (double)u[i] * total;               // mass = random() * sum (buffer.py:141)
  int node = 1;
  while (node < capacity) {                            // find_prefixsum_idx (segment_tree.py:118-139)
    const double left = sum_tree[2 * node];
    if (left > prefix) node = 2 * node;
    else { prefix -= left; node = 2 * node + 1; }
  }
  int id = node - capacity;
  if (id >= size) id = size - 1;                       // rounding at the very end of the mass range
  idx_out[i] = id;
  // importance weights (buffer.py:150-160)
  const double p_min = min_tree[1] / total;
  const double max_w = pow(p_min * size, -beta);
  const double p_s = sum_tree[capacity + id] / total;
  if (w_out) w_out[i] = (float)(pow(p_s * size, -beta) / max_w);
  for (int k = 0; k < obs_dim; ++k) {
    obs_out[(size_t)i * obs_dim + k] = s_obs[(size_t)id * obs_dim + k];
    obs1_out[(size_t)i * obs_dim + k] = s_obs1[(size_t)id * obs_dim + k];
  }
  for (int k = 0; k < act_dim; ++k) act_out[(size_t)i * act_dim + k] = s_act[(size_t)id * act_dim + k];
  rew_out[i] = s_rew[id];
  done_out[i] = s_done[id];
}

}  // namespace mpg

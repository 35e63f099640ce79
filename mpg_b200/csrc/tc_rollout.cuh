// Tensor-core (tcgen05) path: state owned by the handle, weight-image packing, launchers.
#pragma once
#include "tc_gemm.cuh"

namespace mpg {

struct TcNetImages {
  uint8_t* big_fwd = nullptr;   // Wt[n][k] = W2[k][n]   (z2 = h1 . W2)
  uint8_t* big_dx = nullptr;    // Wt[k][n] = W2[k][n]   (g_h1 = delta2 . W2^T)
  uint8_t* l1 = nullptr;        // [W1; b1] as 256 x 16
  uint8_t* in = nullptr;        // W1 as 16 x 256 (input gradient)
};

struct TcState {
  bool ready = false;
  TcNetImages nets[MPG_NUM_NETS];
  uint8_t* scratch_img = nullptr;   // self-test image
};

inline bool tc_init(TcState& t, const mpg_config& cfg, int /*sms*/, size_t& ws_bytes) {
  if (cfg.obs_dim + cfg.act_dim + 1 > 16) return true;   // first-layer K must fit one UMMA k-step; stay on FFMA
  auto alloc = [&](uint8_t** p, size_t n) {
    if (cudaMalloc(p, n) != cudaSuccess) return false;
    ws_bytes += n;
    return true;
  };
  bool ok = alloc(&t.scratch_img, 8 * tc::STAGE_BYTES);
  for (int n = 0; n < MPG_NUM_NETS && ok; ++n)
    ok = alloc(&t.nets[n].big_fwd, 8 * tc::STAGE_BYTES) && alloc(&t.nets[n].big_dx, 8 * tc::STAGE_BYTES)
         && alloc(&t.nets[n].l1, 16384) && alloc(&t.nets[n].in, 16384);
  if (!ok) return false;
  if (cudaFuncSetAttribute(tc::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemMap::TOTAL + 1024)
      != cudaSuccess)
    return false;
  t.ready = false;   // flipped on once the rollout kernels are wired in
  return true;
}

inline void tc_destroy(TcState& t) {
  cudaFree(t.scratch_img);
  for (int n = 0; n < MPG_NUM_NETS; ++n) {
    cudaFree(t.nets[n].big_fwd); cudaFree(t.nets[n].big_dx); cudaFree(t.nets[n].l1); cudaFree(t.nets[n].in);
  }
}

// flat: Keras-order natural weights W1|b1|W2|b2|W3|b3 on the device
inline bool tc_pack_weights(TcState& t, int net, const float* flat, int in_dim, int out_dim, cudaStream_t st) {
  if (!t.nets[net].big_fwd) return true;   // tensor-core path not configured for this handle
  const GradLayout L(in_dim, out_dim);
  tc::pack_big_image<<<32, 256, 0, st>>>(flat + L.oW2, 1, H, t.nets[net].big_fwd);     // value(n,k) = W2[k][n]
  tc::pack_big_image<<<32, 256, 0, st>>>(flat + L.oW2, H, 1, t.nets[net].big_dx);      // value(k,n) = W2[k][n]
  tc::pack_l1_image<<<2, 256, 0, st>>>(flat + L.oW1, flat + L.ob1, in_dim, 15, t.nets[net].l1);
  tc::pack_in_image<<<2, 256, 0, st>>>(flat + L.oW1, in_dim, t.nets[net].in);
  return cudaGetLastError() == cudaSuccess;
}

// self test of one GEMM kind (see tc_gemm.cuh): W is fp32 [256 x 256] (kind 0), [16 x 256] (kinds 1, 2)
inline cudaError_t tc_selftest(TcState& t, int kind, const float* X, const float* W, float* Z, int repeats, cudaStream_t st) {
  if (!t.scratch_img) return cudaErrorNotSupported;
  if (kind == 0) tc::pack_big_image<<<32, 256, 0, st>>>(W, H, 1, t.scratch_img);
  else if (kind == 1) tc::pack_l1_image<<<2, 256, 0, st>>>(W, W, 16, -1, t.scratch_img);
  else tc::pack_in_image<<<2, 256, 0, st>>>(W, 16, t.scratch_img);
  tc::selftest_kernel<<<1, tc::CTA_THREADS, tc::SmemMap::TOTAL + 1024, st>>>(kind, X, t.scratch_img, Z, repeats);
  return cudaGetLastError();
}

}  // namespace mpg

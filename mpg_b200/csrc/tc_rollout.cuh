// Tensor-core (tcgen05) rollout path -- state owned by the handle. Filled in by tc_rollout.cu.
#pragma once
#include "common.cuh"

namespace mpg {

struct TcState {
  bool ready = false;
};

inline bool tc_init(TcState&, const mpg_config&, int /*sms*/, size_t& /*ws_bytes*/) { return true; }
inline void tc_destroy(TcState&) {}
inline bool tc_pack_weights(TcState&, int /*net*/, const float* /*flat*/, int /*in_dim*/, int /*out_dim*/, cudaStream_t) {
  return true;
}

}  // namespace mpg

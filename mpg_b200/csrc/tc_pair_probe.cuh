// CTA-pair (cta_group::2) probe of the streamed 128x256x256 split-bf16 contraction: two CTAs of a cluster each own
// 128 rows of A and HALF of every weight stage (128 of the 256 N rows), the leader issues M = 256 UMMAs that read
// both shared memories.  Per CTA and GEMM that halves the ring writes (128 KB instead of 256 KB) and the B-operand
// reads (192 KB instead of 384 KB) -- the shared-memory traffic that bounds the single-CTA kernel.
// Used by tools/gemm_probe.py (timing) and tests/test_gpu_tc.py (result check); not on the product path.
//
// All tcgen05 instructions of a kernel must use one cta_group, hence a kernel of its own.
#pragma once
#include "tc_gemm.cuh"

namespace mpg {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {   // one full warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair when all previously issued UMMAs are complete
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

#ifndef MPG_PAIR_NSLOT
#define MPG_PAIR_NSLOT 6
#endif
constexpr int PSTAGE_BYTES = 16384;      // the probe's own stage: one split of a 128-feature x 64-element weight block
constexpr int PNSLOT = MPG_PAIR_NSLOT;   // ring depth of the probe (the pair kernel has no MISC / p-image regions to fit)
struct PairBars {
  uint64_t full[PNSLOT], empty[PNSLOT];
  uint64_t peer_full[PNSLOT];   // leader only: the peer's half of the stage has landed (relayed by a peer thread)
  uint64_t a_ready;            // leader only: both CTAs have written their A image
  uint64_t d_full;
  uint32_t tmem_base;
};
// weight image of the probe: 16 stages (k-block, split, n-half) of 128 features x 64 elements, SW128 K-major -- each CTA of
// the pair streams ONE n-half of every (k-block, split)
__global__ void pack_pair_image(const float* __restrict__ src, int rs, int cs, uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, 8-element chunk): 256 x 32
  if (idx >= 256 * 32) return;
  const int row = idx >> 5, cc = idx & 31;
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = src[(size_t)row * rs + (size_t)(cc * 8 + e) * cs];
  uint4 h, l;
  split2x<false>(x[0], x[1], h.x, l.x); split2x<false>(x[2], x[3], h.y, l.y); split2x<false>(x[4], x[5], h.z, l.z); split2x<false>(x[6], x[7], h.w, l.w);
  const int hh = row >> 7, rr = row & 127, kb = cc >> 3, c = cc & 7;
  const size_t stage = (size_t)(kb * 4 + hh) * PSTAGE_BYTES;
  const uint32_t off = (rr >> 3) * 1024 + (rr & 7) * 128 + ((c ^ (rr & 7)) << 4);
  *reinterpret_cast<uint4*>(img + stage + off) = h;
  *reinterpret_cast<uint4*>(img + stage + 2 * PSTAGE_BYTES + off) = l;
}
constexpr int PAIR_SMEM = 2 * ACT_SPLIT + PNSLOT * PSTAGE_BYTES + 256 + 1024;
static_assert(PAIR_SMEM <= 232448, "pair probe shared memory");

// X: fp32 [256 x 256] (rows 0..127 -> CTA 0, 128..255 -> CTA 1), img: packed weight image (pack_pair_image),
// Z: fp32 [256 x 256] = X . Wt^T accumulated `repeats` times when write_z (rep > 1 re-accumulates: timing only)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_probe_kernel(const float* __restrict__ X, const uint8_t* __restrict__ img, float* __restrict__ Z, int repeats) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* act = smem;
  uint8_t* ring = smem + 2 * ACT_SPLIT;
  PairBars* b = reinterpret_cast<PairBars*>(ring + PNSLOT * PSTAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < PNSLOT; ++i) { mbar_init(&b->full[i], 1); mbar_init(&b->empty[i], 1); mbar_init(&b->peer_full[i], 1); }
    mbar_init(&b->a_ready, 2);
    mbar_init(&b->d_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc2(&b->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = b->tmem_base;
  const int pair = blockIdx.x >> 1;                      // independent pairs all run the same problem (timing)
  (void)pair;

  if (warp < 4) {
    // ---- epilogue warps: write this CTA's 128 rows of A as the split-bf16 image, later read D back ----
    const int row = warp * 32 + lane;
    const float* xr = X + (size_t)(rank * 128 + row) * 256;
    for (int cc = 0; cc < 32; ++cc) {
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = xr[cc * 8 + e];
      act_store8(act, act + ACT_SPLIT, row, cc, x);
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
      if (leader) mbar_arrive(&b->a_ready);
      else mbar_arrive_remote(&b->a_ready, 0);
    }
    mbar_wait_cluster(&b->d_full, 0);
    tc_fence_after();
    if (blockIdx.x < 2) {
      const uint32_t lane_base = tmem + TM_WORK + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < 256; c0 += 32) {
        float v[32];
        tmem_ld32(lane_base + c0, v);
        for (int j = 0; j < 32; ++j) Z[(size_t)(rank * 128 + row) * 256 + c0 + j] = v[j];
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ---- producer: this CTA's N half of every (k-block, split) stage; the peer relays "landed" to the leader ----
    if (lane == 0) {
      uint32_t st = 0;
      for (int rep = 0; rep < repeats; ++rep)
        for (int i = 0; i < 8; ++i, ++st) {              // i = kb * 2 + sp
          const uint32_t slot = st % PNSLOT, par = (st / PNSLOT) & 1;
          mbar_wait_cluster(&b->empty[slot], par ^ 1);
          mbar_expect_tx(&b->full[slot], PSTAGE_BYTES);
          bulk_g2s(ring + slot * PSTAGE_BYTES, img + (size_t)((i >> 1) * 4 + (i & 1) * 2 + rank) * PSTAGE_BYTES, PSTAGE_BYTES,
                   &b->full[slot]);
        }
    } else if (lane == 1 && !leader) {
      uint32_t st = 0;
      for (int rep = 0; rep < repeats; ++rep)
        for (int i = 0; i < 8; ++i, ++st) {
          const uint32_t slot = st % PNSLOT, par = (st / PNSLOT) & 1;
          mbar_wait(&b->full[slot], par);
          mbar_arrive_remote(&b->peer_full[slot], 0);
        }
    }
  } else {
    // ---- mma: the leader issues M = 256 UMMAs over both CTAs' A images and both halves of the stage ----
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, 256, 0, 0);
      const uint32_t act_addr = smem_u32(act), ring_addr = smem_u32(ring);
      mbar_wait_cluster(&b->a_ready, 0);
      tc_fence_after();
      uint32_t st = 0;
      for (int rep = 0; rep < repeats; ++rep)
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t dah = make_desc(act_addr + kb * ACT_BLOCK, 16, 1024, LAYOUT_SW128);
          const uint64_t dal = make_desc(act_addr + ACT_SPLIT + kb * ACT_BLOCK, 16, 1024, LAYOUT_SW128);
#pragma unroll
          for (int sp = 0; sp < 2; ++sp, ++st) {
            const uint32_t slot = st % PNSLOT, par = (st / PNSLOT) & 1;
            mbar_wait(&b->full[slot], par);
            mbar_wait_cluster(&b->peer_full[slot], par);
            tc_fence_after();
            const uint64_t db = make_desc(ring_addr + slot * PSTAGE_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ko = (uint64_t)(ks * 2);
              if (sp == 0) {
                umma2_bf16(tmem + TM_WORK, dah + ko, db + ko, idesc, (rep | kb | ks) ? 1u : 0u);
                umma2_bf16(tmem + TM_WORK, dal + ko, db + ko, idesc, 1u);
              } else {
                umma2_bf16(tmem + TM_WORK, dah + ko, db + ko, idesc, 1u);
              }
            }
            umma2_commit_both(&b->empty[slot]);
          }
        }
      umma2_commit_both(&b->d_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 5) tmem_dealloc2(tmem, TMEM_COLS);
}

}  // namespace tc
}  // namespace mpg

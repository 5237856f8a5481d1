// lbm_fuse2.cuh — temporal blocking: TWO time steps per pass over HBM.
//
// The one-step kernel already moves exactly the algorithmic 72 B per cell update at
// the HBM copy ceiling, so the only way to go faster is to move fewer bytes: read the
// state once, advance it two steps and write it once — 36 B per update instead of 72
// (plus a ~1 % halo).  A thread block marches up a column strip of TX = 128*W cells.
// In every iteration it
//   phase 1: advances row y+1 of step t by one step (collision, step t+1's
//            accelerate_flow if it is row ny-2) and keeps the nine step-(t+1) values in
//            a ring of rows in shared memory instead of storing them to HBM.  A spare
//            warp computes the two halo columns left and right of the strip redundantly;
//   phase 2: pulls row y of step t+1 from the ring (rows y-1, y, y+1), collides it and
//            stores step t+2 to the destination lattice (+ the neighbours' ghost rows).
// Arithmetic is the one-step kernel's, so the lattice is bit-identical to running two
// one-step launches, and the per-segment fp32 speed sums (hence av_vels) are identical too.
//
// This header holds what the two-step kernels share (arguments, tiling, mbarrier / bulk-copy
// wrappers): fuse2q_kernel (lbm_fuse2q.cuh, the default: two-deep stage) and its A/B predecessor
// fuse2p_kernel (lbm_fuse2p.cuh).  Round 1's earlier generations (register prefetch 106 GLUPS, first
// TMA-staged version 128 GLUPS) were retired.
//
// Replaces two iterations of the reference's host loop d2q9-bgk.c:221-238
// (2 x accelerate_flow + 2 x timestep, kernels.cl:9-231).
#pragma once

#include "lbm_kernels.cuh"

namespace lbm {

struct Fuse2Args {
  StepArgs s;              // lattice, mask, constants, ghost pointers and ring flags (s.segs, s.accel_row, s.partials unused)
  int y0;                  // global row of local row 0
  int ny;                  // global rows
  int last;                // 1: the second step is the run's last step (no accelerate folded into its stores)
  int strips;              // column strips of TX cells
  int segs_y;              // row segments per strip
  int seg_rows;            // rows per segment
  int seg_long, n_long;    // fuse2p_kernel: the first n_long segments of a strip have seg_long rows (0: all seg_rows)
  int l2_ahead;            // TMA kernel: rows ahead of the stage load that are prefetched into L2 (0 = off; measured: off is best)
  double2* partials1;      // Σ|u| partials of the first step  [per_step entries, first strips*segs_y used]
  double2* partials2;      // Σ|u| partials of the second step
  long long per_step;      // entries per step in the partial buffer (the tail is zeroed here)
};

// rows [ys, ye) of row segment sy: n_long long segments first, then segments of seg_rows rows
__host__ __device__ inline void f2_segment_rows(int sy, int seg_rows, int seg_long, int n_long, int rows, int& ys, int& ye) {
  if (sy < n_long) {
    ys = sy * seg_long;
    ye = ys + seg_long;
  } else {
    ys = n_long * seg_long + (sy - n_long) * seg_rows;
    ye = ys + seg_rows;
  }
  if (ye > rows) ye = rows;
}

// global row index (periodic) of local row r, which may be a ghost row
__device__ __forceinline__ int global_row(int r, int y0, int ny) {
  int g = y0 + r;
  if (g < 0) g += ny;
  if (g >= ny) g -= ny;
  return g;
}

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) wrappers shared by the two-step kernels
// ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

}  // namespace lbm

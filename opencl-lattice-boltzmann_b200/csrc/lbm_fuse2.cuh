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
// wrappers) and fuse2_tma_kernel, the first TMA-staged version, kept as the A/B predecessor of
// the default fuse2p_kernel (lbm_fuse2p.cuh).  (Round 1's register-prefetch variant, 106 GLUPS,
// was retired.)
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
// fuse2_tma_kernel — the same two-step march with the step-t rows brought in by
// the TMA engine instead of by the warps.
//
// One thread issues nine cp.async.bulk copies per row (one per plane: the TX+8
// floats [x0-4, x0+TX+4) of the source row that plane is pulled from) into a
// stage buffer, completion signalled on an mbarrier; they are in flight while the
// block runs phase 2 of the previous row, so no warp waits on HBM.  Phase 1 then
// reads the stage exactly like phase 2 reads the ring (aligned LDS.128 + shuffles;
// the x-1 / x+4 neighbours sit in the stage as well).  Only the periodic x wrap at
// the two outermost strips and the strip's two halo columns still use scalar global
// loads, prefetched one row ahead.
//
// The ring is sized by lifetime: of a step-(t+1) row, planes 4,7,8 are consumed in
// the iteration that produces them (2 slots), planes 0,1,3 one iteration later
// (3 slots), planes 2,5,6 two iterations later (4 slots) — 27 plane-rows instead of
// 36, so ring + stage take the shared memory the plain ring took.
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

template <int W>
constexpr int fuse2_tma_smem_bytes() { return (27 + NSPEEDS) * (128 * W + 8) * (int)sizeof(float) + 16; }

template <int W, bool PACKED, int MINB>
__global__ void __launch_bounds__(32 * (W + 1), MINB) fuse2_tma_kernel(const __grid_constant__ Fuse2Args fa) {
  constexpr int V = 4;
  constexpr int TX = 128 * W;
  constexpr int RS = TX + 8;        // row stride per plane: columns x0-4 .. x0+TX+3 (cell j at index 4+j)
  extern __shared__ __align__(128) float smem[];
  float* stage = smem;                               // [9][RS]   step-t rows for the next phase 1
  float* ring_n = stage + NSPEEDS * RS;              // [2][3][RS] planes 4,7,8 of step t+1
  float* ring_m = ring_n + 2 * 3 * RS;               // [3][3][RS] planes 0,1,3
  float* ring_s = ring_m + 3 * 3 * RS;               // [4][3][RS] planes 2,5,6
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_s + 4 * 3 * RS);
  __shared__ double part_hi[2][W], part_lo[2][W];

  const StepArgs& a = fa.s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool halo_warp = (warp == W);
  const int nx = a.nx, rows = a.rows;
  const long long ps = a.plane_stride;

  int strip, sy;
  {
    const int b = blockIdx.x, ns = fa.strips;
    if (fa.segs_y < 2 || b < ns) { sy = b / ns; strip = b - sy * ns; }
    else if (b < 2 * ns) { sy = fa.segs_y - 1; strip = b - ns; }
    else { sy = 1 + (b - 2 * ns) / ns; strip = (b - 2 * ns) % ns; }
  }
  const int ys = sy * fa.seg_rows, ye = min(rows, ys + fa.seg_rows);
  const int x0 = strip * TX;
  const int ncol = min(TX, nx - x0);
  const int j0 = (warp * 32 + lane) * V;
  const int c = 4 + j0;                         // the thread's first column in stage / ring rows
  const int xb = x0 + j0;
  const bool active = !halo_warp && j0 < ncol;
  const bool need_l = active && lane == 0;
  const bool need_r = active && (lane == 31 || j0 + V >= ncol);
  const bool wrap_l = need_l && xb == 0;        // x-1 wraps to nx-1: not in the stage
  const bool wrap_r = need_r && xb + V >= nx;   // x+4 wraps to 0
  // segments that read ghost rows or whose rows 0,1 / rows-2,rows-1 are stored into a neighbour's ghost rows
  const bool touches_bottom = (ys < 2), touches_top = (ye >= rows - 1);

  if (threadIdx.x == 0) mbar_init(full, 1);
  if (a.edge_count != nullptr && (touches_bottom || touches_top)) {
    if (threadIdx.x == 0) {
      if (touches_top) wait_epoch(a.flag_from_up, a.epoch - 1, a.error_word, a.wait_timeout_ns);
      if (touches_bottom) wait_epoch(a.flag_from_down, a.epoch - 1, a.error_word, a.wait_timeout_ns);
    }
  }
  __syncthreads();

  const int accel_g = fa.ny - 2;
  double hi1 = 0.0, lo1 = 0.0, hi2 = 0.0, lo2 = 0.0;

  // which source row plane k is pulled from (kernels.cl:104-112): 0,1,3 own row; 2,5,6 south; 4,7,8 north
  auto issue_row = [&](int r) {   // one thread: the nine plane-rows phase 1 of row r needs -> stage
    mbar_expect_tx(full, NSPEEDS * RS * (uint32_t)sizeof(float));
    const float* base = a.src + (long long)r * a.pitch + (x0 - 4);
#pragma unroll
    for (int k = 0; k < NSPEEDS; k++) {
      const int dy = (k == 2 || k == 5 || k == 6) ? -1 : (k == 4 || k == 7 || k == 8) ? 1 : 0;
      tma_load_1d(stage + k * RS, base + k * ps + (long long)dy * a.pitch, RS * (uint32_t)sizeof(float), full);
    }
    // optional: pull the rows a few iterations ahead from HBM into L2 (measured slower at 16384^2: 128 -> 110-120 GLUPS)
    const int rp = r + fa.l2_ahead;
    if (fa.l2_ahead > 0 && rp <= ye) {
      const float* pb = a.src + (long long)rp * a.pitch + (x0 - 4);
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) {
        const int dy = (k == 2 || k == 5 || k == 6) ? -1 : (k == 4 || k == 7 || k == 8) ? 1 : 0;
        tma_prefetch_l2(pb + k * ps + (long long)dy * a.pitch, RS * (uint32_t)sizeof(float));
      }
    }
  };

  // scalar global loads that the stage cannot serve, prefetched one row ahead:
  // the x wrap of the outermost strips (body lanes) and the two halo columns (halo warp)
  float we1 = 0.f, we5 = 0.f, we8 = 0.f, we3 = 0.f, we6 = 0.f, we7 = 0.f;
  float ht[NSPEEDS];
  bool hfluid = true;
#pragma unroll
  for (int k = 0; k < NSPEEDS; k++) ht[k] = 0.0f;
  const int xh = (lane == 0) ? ((x0 == 0) ? nx - 1 : x0 - 1) : ((x0 + ncol >= nx) ? 0 : x0 + ncol);
  const int xhw = (xh == 0) ? nx - 1 : xh - 1;
  const int xhe = (xh + 1 >= nx) ? 0 : xh + 1;
  auto load_scalars = [&](int r) {
    const float* s_mid = a.src + (long long)r * a.pitch;
    const float* s_south = s_mid - a.pitch;
    const float* s_north = s_mid + a.pitch;
    if (wrap_l) {
      we1 = load_one<0>(s_mid + 1 * ps + nx - 1);
      we5 = load_one<0>(s_south + 5 * ps + nx - 1);
      we8 = load_one<0>(s_north + 8 * ps + nx - 1);
    }
    if (wrap_r) {
      we3 = load_one<0>(s_mid + 3 * ps);
      we6 = load_one<0>(s_south + 6 * ps);
      we7 = load_one<0>(s_north + 7 * ps);
    }
    if (halo_warp && lane < 2) {
      ht[0] = load_one<0>(s_mid + 0 * ps + xh);
      ht[1] = load_one<0>(s_mid + 1 * ps + xhw);
      ht[2] = load_one<0>(s_south + 2 * ps + xh);
      ht[3] = load_one<0>(s_mid + 3 * ps + xhe);
      ht[4] = load_one<0>(s_north + 4 * ps + xh);
      ht[5] = load_one<0>(s_south + 5 * ps + xhw);
      ht[6] = load_one<0>(s_south + 6 * ps + xhe);
      ht[7] = load_one<0>(s_north + 7 * ps + xhe);
      ht[8] = load_one<0>(s_north + 8 * ps + xhw);
      hfluid = ((__ldg(a.mask + (long long)r * a.mask_pitch + (xh >> 5)) >> (xh & 31)) & 1u) == 0u;
    }
  };

  // ring addressing: row r's planes 4,7,8 / 0,1,3 / 2,5,6 (index within the group: 0,1,2)
  auto rn = [&](int r) { return ring_n + ((r + 4) & 1) * (3 * RS); };
  auto rm = [&](int r) { return ring_m + ((r + 6) % 3) * (3 * RS); };
  auto rs = [&](int r) { return ring_s + (r & 3) * (3 * RS); };

  auto lds4 = [&](const float* q, float (&v)[V]) {
    const float4 f = *reinterpret_cast<const float4*>(q);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  };

  // ---- phase 1: step t -> t+1 of row r, inputs from the stage ----
  auto phase1 = [&](int r) {
    const bool accel = (global_row(r, fa.y0, fa.ny) == accel_g);
    if (!halo_warp) {
      // shared-memory reads are unconditional: every index used lies inside the row buffers, and what
      // inactive lanes (columns beyond the strip) compute from it is never stored or summed
      float q[NSPEEDS][V];
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) lds4(stage + k * RS + c, q[k]);
      uint32_t bits = 0;
      if (active) bits = __ldg(a.mask + (long long)r * a.mask_pitch + (xb >> 5)) >> (xb & 31);
      float f1 = stage[1 * RS + c - 1], f5 = stage[5 * RS + c - 1], f8 = stage[8 * RS + c - 1];
      float f3 = stage[3 * RS + c + V], f6 = stage[6 * RS + c + V], f7 = stage[7 * RS + c + V];
      if (wrap_l) { f1 = we1; f5 = we5; f8 = we8; }
      if (wrap_r) { f3 = we3; f6 = we6; f7 = we7; }
      float l1 = __shfl_up_sync(FULL, q[1][V - 1], 1);
      float l5 = __shfl_up_sync(FULL, q[5][V - 1], 1);
      float l8 = __shfl_up_sync(FULL, q[8][V - 1], 1);
      float r3 = __shfl_down_sync(FULL, q[3][0], 1);
      float r6 = __shfl_down_sync(FULL, q[6][0], 1);
      float r7 = __shfl_down_sync(FULL, q[7][0], 1);
      if (lane == 0) { l1 = f1; l5 = f5; l8 = f8; }
      if (need_r) { r3 = f3; r6 = f6; r7 = f7; }

      float out[NSPEEDS][V];
      float tot = compute_cells<V, PACKED>(q, l1, l5, l8, r3, r6, r7, bits, a.omega, accel, a.w1, a.w2, out);
      if (active) {
        float* n = rn(r) + c;
        float* m = rm(r) + c;
        float* so = rs(r) + c;
        auto st4 = [&](float* d, const float (&v)[V]) { *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]); };
        st4(m + 0 * RS, out[0]); st4(m + 1 * RS, out[1]); st4(m + 2 * RS, out[3]);
        st4(so + 0 * RS, out[2]); st4(so + 1 * RS, out[5]); st4(so + 2 * RS, out[6]);
        st4(n + 0 * RS, out[4]); st4(n + 1 * RS, out[7]); st4(n + 2 * RS, out[8]);
      } else {
        tot = 0.0f;
      }
      if (r >= ys && r < ye) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) tot = __fadd_rn(tot, __shfl_xor_sync(FULL, tot, s));
        dd_add(hi1, lo1, (double)tot, 0.0);   // every lane keeps the same sum: no branch
      }
    } else if (lane < 2) {
      float o[NSPEEDS];
      collide_cell(ht, hfluid, a.omega, o);
      if (accel) accelerate_cell(o, hfluid, a.w1, a.w2);
      const int idx = (lane == 0) ? 3 : 4 + ncol;
      float* n = rn(r) + idx;
      float* m = rm(r) + idx;
      float* so = rs(r) + idx;
      m[0 * RS] = o[0]; m[1 * RS] = o[1]; m[2 * RS] = o[3];
      so[0 * RS] = o[2]; so[1 * RS] = o[5]; so[2 * RS] = o[6];
      n[0 * RS] = o[4]; n[1 * RS] = o[7]; n[2 * RS] = o[8];
    }
  };

  // ---- phase 2: step t+1 -> t+2 of row y, inputs from the ring ----
  auto phase2 = [&](int y) {
    if (halo_warp) return;
    const float* m = rm(y) + c;        // planes 0,1,3 of row y
    const float* so = rs(y - 1) + c;   // planes 2,5,6 of row y-1
    const float* n = rn(y + 1) + c;    // planes 4,7,8 of row y+1
    float q[NSPEEDS][V];
    lds4(m + 0 * RS, q[0]); lds4(m + 1 * RS, q[1]); lds4(m + 2 * RS, q[3]);
    lds4(so + 0 * RS, q[2]); lds4(so + 1 * RS, q[5]); lds4(so + 2 * RS, q[6]);
    lds4(n + 0 * RS, q[4]); lds4(n + 1 * RS, q[7]); lds4(n + 2 * RS, q[8]);
    uint32_t bits = 0;
    if (active) bits = __ldg(a.mask + (long long)y * a.mask_pitch + (xb >> 5)) >> (xb & 31);
    const float f1 = m[1 * RS - 1], f5 = so[1 * RS - 1], f8 = n[2 * RS - 1];
    const float f3 = m[2 * RS + V], f6 = so[2 * RS + V], f7 = n[1 * RS + V];
    float l1 = __shfl_up_sync(FULL, q[1][V - 1], 1);
    float l5 = __shfl_up_sync(FULL, q[5][V - 1], 1);
    float l8 = __shfl_up_sync(FULL, q[8][V - 1], 1);
    float r3 = __shfl_down_sync(FULL, q[3][0], 1);
    float r6 = __shfl_down_sync(FULL, q[6][0], 1);
    float r7 = __shfl_down_sync(FULL, q[7][0], 1);
    if (lane == 0) { l1 = f1; l5 = f5; l8 = f8; }
    if (need_r) { r3 = f3; r6 = f6; r7 = f7; }

    const bool accel = !fa.last && (global_row(y, fa.y0, fa.ny) == accel_g);
    float out[NSPEEDS][V];
    float tot = compute_cells<V, PACKED>(q, l1, l5, l8, r3, r6, r7, bits, a.omega, accel, a.w1, a.w2, out);
    if (active) {
      float* d = a.dst + (long long)y * a.pitch + xb;
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(d + k * ps, out[k]);
      if (y >= rows - 2) {
        float* g = a.up_ghost + (long long)(y - (rows - 1)) * a.pitch + xb;
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(g + k * a.up_plane_stride, out[k]);
      }
      if (y < 2) {
        float* g = a.down_ghost + (long long)y * a.pitch + xb;
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(g + k * a.down_plane_stride, out[k]);
      }
    } else {
      tot = 0.0f;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) tot = __fadd_rn(tot, __shfl_xor_sync(FULL, tot, s));
    dd_add(hi2, lo2, (double)tot, 0.0);
  };

  // prologue: rows ys-1 and ys of step t+1, then row ys+1's inputs in flight
  uint32_t parity = 0;
  load_scalars(ys - 1);
  if (threadIdx.x == 0) issue_row(ys - 1);
  for (int r = ys - 1; r <= ys; r++) {
    mbar_wait(full, parity);
    parity ^= 1;
    phase1(r);
    load_scalars(r + 1);
    __syncthreads();                       // every warp is done with the stage
    if (threadIdx.x == 0) issue_row(r + 1);
  }
  for (int y = ys; y < ye; y++) {
    mbar_wait(full, parity);               // the stage holds row y+1's inputs
    parity ^= 1;
    phase1(y + 1);
    if (y + 1 < ye) load_scalars(y + 2);
    __syncthreads();                       // stage free again; rows y-1, y, y+1 of step t+1 visible in the ring
    if (threadIdx.x == 0 && y + 1 < ye) issue_row(y + 2);   // in flight during phase 2
    phase2(y);
  }

  if (!halo_warp && lane == 0) {
    part_hi[0][warp] = hi1; part_lo[0][warp] = lo1;
    part_hi[1][warp] = hi2; part_lo[1][warp] = lo2;
  }
  if (a.edge_count != nullptr && (touches_bottom || touches_top)) __threadfence_system();
  __syncthreads();
  if (threadIdx.x < 2) {
    double h = 0.0, l = 0.0;
#pragma unroll
    for (int i = 0; i < W; i++) dd_add(h, l, part_hi[threadIdx.x][i], part_lo[threadIdx.x][i]);
    (threadIdx.x == 0 ? fa.partials1 : fa.partials2)[blockIdx.x] = make_double2(h, l);
  }
  for (long long i = (long long)gridDim.x + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < fa.per_step;
       i += (long long)gridDim.x * blockDim.x) {
    fa.partials1[i] = make_double2(0.0, 0.0);
    fa.partials2[i] = make_double2(0.0, 0.0);
  }
  if (a.edge_count != nullptr && threadIdx.x == 0) {
    if (touches_bottom && atomicAdd(a.edge_count + 0, 1ULL) + 1ULL == a.edge_target) {
      __threadfence_system();
      st_release_sys(a.peer_down_flag, a.epoch);
    }
    if (touches_top && atomicAdd(a.edge_count + 1, 1ULL) + 1ULL == a.edge_target_top) {
      __threadfence_system();
      st_release_sys(a.peer_up_flag, a.epoch);
    }
  }
}

}  // namespace lbm

// lbm_fuse2.cuh — temporal blocking: TWO time steps per pass over HBM.
//
// The one-step kernel already moves exactly the algorithmic 72 B per cell update at
// the HBM copy ceiling, so the only way to go faster is to move fewer bytes: this
// kernel reads the state once, advances it two steps and writes it once — 36 B per
// update instead of 72 (plus a ~1 % halo).  A thread block marches up a column strip
// of TX = 128*W cells.  In every iteration it
//   phase 1: pulls row y+1 of step t from global memory (exactly like the one-step
//            kernel), collides it, applies step t+1's accelerate_flow if it is row
//            ny-2, and keeps the nine step-(t+1) values in a 4-row ring in shared
//            memory instead of storing them to HBM.  A spare warp computes the two
//            halo columns left and right of the strip redundantly;
//   phase 2: pulls row y of step t+1 from the ring (rows y-1, y, y+1), collides it and
//            stores step t+2 to the destination lattice (+ the neighbours' ghost rows).
// One __syncthreads per row.  Arithmetic is the same compute_cells() as the one-step
// kernel, so the lattice is bit-identical to running two one-step launches, and the
// per-segment fp32 speed sums (hence av_vels) are identical too.
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
  double2* partials1;      // Σ|u| partials of the first step  [per_step entries, first strips*segs_y used]
  double2* partials2;      // Σ|u| partials of the second step
  long long per_step;      // entries per step in the partial buffer (the tail is zeroed here)
};

constexpr int F2_RING = 4;

template <int W>
constexpr int fuse2_smem_bytes() { return F2_RING * NSPEEDS * (128 * W + 8) * (int)sizeof(float); }

// global row index (periodic) of local row r, which may be a ghost row
__device__ __forceinline__ int global_row(int r, int y0, int ny) {
  int g = y0 + r;
  if (g < 0) g += ny;
  if (g >= ny) g -= ny;
  return g;
}

template <int W, bool PACKED, int MINB>
__global__ void __launch_bounds__(32 * (W + 1), MINB) fuse2_kernel(const __grid_constant__ Fuse2Args fa) {
  constexpr int V = 4;
  constexpr int TX = 128 * W;
  constexpr int RS = TX + 8;        // ring row stride per plane: [pad 3][left halo][TX cells][right halo][pad 3]
  extern __shared__ __align__(16) float ring[];   // [F2_RING][NSPEEDS][RS]
  __shared__ double part_hi[2][W], part_lo[2][W];

  const StepArgs& a = fa.s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool halo_warp = (warp == W);
  const int nx = a.nx, rows = a.rows;
  const long long ps = a.plane_stride;

  // block -> (strip, row segment); the segments touching the slab's edges first
  int strip, sy;
  {
    const int b = blockIdx.x, ns = fa.strips;
    if (fa.segs_y < 2 || b < ns) { sy = b / ns; strip = b - sy * ns; }
    else if (b < 2 * ns) { sy = fa.segs_y - 1; strip = b - ns; }
    else { sy = 1 + (b - 2 * ns) / ns; strip = (b - 2 * ns) % ns; }
  }
  const int ys = sy * fa.seg_rows, ye = min(rows, ys + fa.seg_rows);
  const int x0 = strip * TX;
  const int ncol = min(TX, nx - x0);            // columns of this strip (a multiple of 4)
  const int j0 = (warp * 32 + lane) * V;        // the thread's first column within the strip (body warps)
  const bool active = !halo_warp && j0 < ncol;
  const bool touches_bottom = (ys == 0), touches_top = (ye == rows);

  if (a.edge_count != nullptr && (touches_bottom || touches_top)) {  // ring: neighbours' previous epoch complete
    if (threadIdx.x == 0) {
      if (touches_top) wait_epoch(a.flag_from_up, a.epoch - 1);
      if (touches_bottom) wait_epoch(a.flag_from_down, a.epoch - 1);
    }
    __syncthreads();
  }

  const int accel_g = fa.ny - 2;                // kernels.cl:18
  double hi1 = 0.0, lo1 = 0.0, hi2 = 0.0, lo2 = 0.0;   // lane 0 of each body warp: Σ of its segment sums

  // ---- phase 1: step t -> t+1 for local row r (ys-1 .. ye; -1 and `rows` are ghost rows), into the ring ----
  auto phase1 = [&](int r) {
    float* rrow = ring + (r & (F2_RING - 1)) * (NSPEEDS * RS);
    const bool accel = (global_row(r, fa.y0, fa.ny) == accel_g);   // the second step always follows
    const float* s_mid = a.src + (long long)r * a.pitch;
    const float* s_south = s_mid - a.pitch;
    const float* s_north = s_mid + a.pitch;
    if (!halo_warp) {
      float p[NSPEEDS][V];
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++)
#pragma unroll
        for (int j = 0; j < V; j++) p[k][j] = 0.0f;
      float e1 = 0.f, e5 = 0.f, e8 = 0.f, e3 = 0.f, e6 = 0.f, e7 = 0.f;
      uint32_t bits = 0;
      const int x = x0 + j0;
      const bool need_l = active && lane == 0;
      const bool need_r = active && (lane == 31 || j0 + V >= ncol);
      if (active) {
        load_vec<V, 0>(s_mid + 0 * ps + x, p[0]);
        load_vec<V, 0>(s_mid + 1 * ps + x, p[1]);
        load_vec<V, 0>(s_south + 2 * ps + x, p[2]);
        load_vec<V, 0>(s_mid + 3 * ps + x, p[3]);
        load_vec<V, 0>(s_north + 4 * ps + x, p[4]);
        load_vec<V, 0>(s_south + 5 * ps + x, p[5]);
        load_vec<V, 0>(s_south + 6 * ps + x, p[6]);
        load_vec<V, 0>(s_north + 7 * ps + x, p[7]);
        load_vec<V, 0>(s_north + 8 * ps + x, p[8]);
        bits = __ldg(a.mask + (long long)r * a.mask_pitch + (x >> 5)) >> (x & 31);
      }
      if (need_l) {
        const int xl = (x == 0) ? nx - 1 : x - 1;
        e1 = load_one<0>(s_mid + 1 * ps + xl);
        e5 = load_one<0>(s_south + 5 * ps + xl);
        e8 = load_one<0>(s_north + 8 * ps + xl);
      }
      if (need_r) {
        const int xr = (x + V >= nx) ? 0 : x + V;
        e3 = load_one<0>(s_mid + 3 * ps + xr);
        e6 = load_one<0>(s_south + 6 * ps + xr);
        e7 = load_one<0>(s_north + 7 * ps + xr);
      }
      float l1 = __shfl_up_sync(FULL, p[1][V - 1], 1);
      float l5 = __shfl_up_sync(FULL, p[5][V - 1], 1);
      float l8 = __shfl_up_sync(FULL, p[8][V - 1], 1);
      float r3 = __shfl_down_sync(FULL, p[3][0], 1);
      float r6 = __shfl_down_sync(FULL, p[6][0], 1);
      float r7 = __shfl_down_sync(FULL, p[7][0], 1);
      if (lane == 0) { l1 = e1; l5 = e5; l8 = e8; }
      if (need_r) { r3 = e3; r6 = e6; r7 = e7; }

      float out[NSPEEDS][V];
      float tot = compute_cells<V, PACKED>(p, l1, l5, l8, r3, r6, r7, bits, a.omega, accel, a.w1, a.w2, out);
      if (active) {
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++)
          *reinterpret_cast<float4*>(rrow + k * RS + 4 + j0) = make_float4(out[k][0], out[k][1], out[k][2], out[k][3]);
      } else {
        tot = 0.0f;
      }
      if (r >= ys && r < ye) {   // owned rows count towards the first step's average
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) tot = __fadd_rn(tot, __shfl_xor_sync(FULL, tot, s));
        if (lane == 0) dd_add(hi1, lo1, (double)tot, 0.0);
      }
    } else if (lane < 2) {
      // the column left (lane 0) / right (lane 1) of the strip, with the x wrap (kernels.cl:100-102)
      const int xh = (lane == 0) ? ((x0 == 0) ? nx - 1 : x0 - 1) : ((x0 + ncol >= nx) ? 0 : x0 + ncol);
      const int xw = (xh == 0) ? nx - 1 : xh - 1;
      const int xe = (xh + 1 >= nx) ? 0 : xh + 1;
      float t[NSPEEDS], o[NSPEEDS];
      t[0] = load_one<0>(s_mid + 0 * ps + xh);
      t[1] = load_one<0>(s_mid + 1 * ps + xw);
      t[2] = load_one<0>(s_south + 2 * ps + xh);
      t[3] = load_one<0>(s_mid + 3 * ps + xe);
      t[4] = load_one<0>(s_north + 4 * ps + xh);
      t[5] = load_one<0>(s_south + 5 * ps + xw);
      t[6] = load_one<0>(s_south + 6 * ps + xe);
      t[7] = load_one<0>(s_north + 7 * ps + xe);
      t[8] = load_one<0>(s_north + 8 * ps + xw);
      const bool fluid = ((__ldg(a.mask + (long long)r * a.mask_pitch + (xh >> 5)) >> (xh & 31)) & 1u) == 0u;
      collide_cell(t, fluid, a.omega, o);
      if (accel) accelerate_cell(o, fluid, a.w1, a.w2);
      const int idx = (lane == 0) ? 3 : 4 + ncol;
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) rrow[k * RS + idx] = o[k];
    }
  };

  // ---- phase 2: step t+1 -> t+2 for local row y (ys .. ye-1), pulled from the ring, stored to dst ----
  auto phase2 = [&](int y) {
    if (halo_warp) return;
    const float* r_mid = ring + (y & (F2_RING - 1)) * (NSPEEDS * RS);
    const float* r_south = ring + ((y - 1) & (F2_RING - 1)) * (NSPEEDS * RS);
    const float* r_north = ring + ((y + 1) & (F2_RING - 1)) * (NSPEEDS * RS);
    const int c = 4 + j0;
    const int x = x0 + j0;
    float p[NSPEEDS][V];
#pragma unroll
    for (int k = 0; k < NSPEEDS; k++)
#pragma unroll
      for (int j = 0; j < V; j++) p[k][j] = 0.0f;
    float e1 = 0.f, e5 = 0.f, e8 = 0.f, e3 = 0.f, e6 = 0.f, e7 = 0.f;
    uint32_t bits = 0;
    const bool need_l = active && lane == 0;
    const bool need_r = active && (lane == 31 || j0 + V >= ncol);
    auto lds4 = [&](const float* q, float (&v)[V]) {
      const float4 f = *reinterpret_cast<const float4*>(q);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    };
    if (active) {
      lds4(r_mid + 0 * RS + c, p[0]);
      lds4(r_mid + 1 * RS + c, p[1]);
      lds4(r_south + 2 * RS + c, p[2]);
      lds4(r_mid + 3 * RS + c, p[3]);
      lds4(r_north + 4 * RS + c, p[4]);
      lds4(r_south + 5 * RS + c, p[5]);
      lds4(r_south + 6 * RS + c, p[6]);
      lds4(r_north + 7 * RS + c, p[7]);
      lds4(r_north + 8 * RS + c, p[8]);
      bits = __ldg(a.mask + (long long)y * a.mask_pitch + (x >> 5)) >> (x & 31);
    }
    if (need_l) { e1 = r_mid[1 * RS + c - 1]; e5 = r_south[5 * RS + c - 1]; e8 = r_north[8 * RS + c - 1]; }
    if (need_r) { e3 = r_mid[3 * RS + c + V]; e6 = r_south[6 * RS + c + V]; e7 = r_north[7 * RS + c + V]; }
    float l1 = __shfl_up_sync(FULL, p[1][V - 1], 1);
    float l5 = __shfl_up_sync(FULL, p[5][V - 1], 1);
    float l8 = __shfl_up_sync(FULL, p[8][V - 1], 1);
    float r3 = __shfl_down_sync(FULL, p[3][0], 1);
    float r6 = __shfl_down_sync(FULL, p[6][0], 1);
    float r7 = __shfl_down_sync(FULL, p[7][0], 1);
    if (lane == 0) { l1 = e1; l5 = e5; l8 = e8; }
    if (need_r) { r3 = e3; r6 = e6; r7 = e7; }

    const bool accel = !fa.last && (global_row(y, fa.y0, fa.ny) == accel_g);
    float out[NSPEEDS][V];
    float tot = compute_cells<V, PACKED>(p, l1, l5, l8, r3, r6, r7, bits, a.omega, accel, a.w1, a.w2, out);
    if (active) {
      float* d = a.dst + (long long)y * a.pitch + x;
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(d + k * ps, out[k]);
      if (y >= rows - 2) {   // the up neighbour's ghost rows -1, -2
        float* g = a.up_ghost + (long long)(y - (rows - 1)) * a.pitch + x;
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(g + k * a.up_plane_stride, out[k]);
      }
      if (y < 2) {           // the down neighbour's ghost rows rows, rows+1
        float* g = a.down_ghost + (long long)y * a.pitch + x;
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(g + k * a.down_plane_stride, out[k]);
      }
    } else {
      tot = 0.0f;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) tot = __fadd_rn(tot, __shfl_xor_sync(FULL, tot, s));
    if (lane == 0) dd_add(hi2, lo2, (double)tot, 0.0);
  };

  phase1(ys - 1);
  phase1(ys);
  for (int y = ys; y < ye; y++) {
    phase1(y + 1);
    __syncthreads();   // rows y-1, y, y+1 of step t+1 are in the ring; the slot of row y+2 is free again
    phase2(y);
  }

  // Σ|u| partials of both steps: body warps' double-doubles added error-free
  if (!halo_warp && lane == 0) {
    part_hi[0][warp] = hi1; part_lo[0][warp] = lo1;
    part_hi[1][warp] = hi2; part_lo[1][warp] = lo2;
  }
  if (a.edge_count != nullptr && (touches_bottom || touches_top)) __threadfence_system();   // edge stores before the count
  __syncthreads();
  if (threadIdx.x < 2) {
    double h = 0.0, l = 0.0;
#pragma unroll
    for (int i = 0; i < W; i++) dd_add(h, l, part_hi[threadIdx.x][i], part_lo[threadIdx.x][i]);
    (threadIdx.x == 0 ? fa.partials1 : fa.partials2)[blockIdx.x] = make_double2(h, l);
  }
  // the partial buffer is sized for the one-step kernel's block count: clear the unused tail
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
    if (touches_top && atomicAdd(a.edge_count + 1, 1ULL) + 1ULL == a.edge_target) {
      __threadfence_system();
      st_release_sys(a.peer_up_flag, a.epoch);
    }
  }
}

}  // namespace lbm

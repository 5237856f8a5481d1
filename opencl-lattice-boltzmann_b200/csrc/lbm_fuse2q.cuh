// lbm_fuse2q.cuh — fuse2p_kernel with a TWO-DEEP stage: the default two-step kernel (option fuse2_tma = 3).
//
// fuse2p_kernel keeps one stage buffer per block: the bulk copies of row r+1 can only be requested once
// every body warp has read row r's stage, so a row's inputs have exactly one row time to arrive, and ncu
// shows the warps waiting for them (mbarrier wait = long scoreboard, ~20 % of the stall cycles).  A
// second stage buffer needs 18.7 KB that a block does not have at three blocks per SM — unless the ring
// gives them up: with a SECOND block barrier per row (after phase 2) no warp can run ahead into the next
// phase 1 while another still reads the ring, so every ring group needs one slot less (18 plane-rows
// instead of 27).  Same shared memory, same results; rows are requested two ahead and have two row times
// to arrive; the price is one more bar.sync per row.  Measured on one B200 at 16384^2: 175.9 GLUPS against
// fuse2p_kernel's 168.2 on the same box.
//
// Everything else (tiling, arithmetic, mask ring, wrap columns, halo warp, ghost-row stores, epoch flags,
// partial sums) is fuse2p_kernel's; see lbm_fuse2p.cuh.
#pragma once

#include "lbm_fuse2p.cuh"

namespace lbm {

template <int W>
constexpr int fuse2q_smem_bytes() {
  return (18 + 2 * NSPEEDS) * (128 * W + 8) * (int)sizeof(float) + 5 * 8 + NSPEEDS * 8 + 4 * 4 * W * (int)sizeof(uint32_t) +
         2 * 6 * 4 * (int)sizeof(float);
}

// MODE bit 0: one reciprocal / square-root range check per thread instead of per pair (compute_quad<JOINT>).
// MODE bit 1: dry run for bandwidth experiments — same memory traffic, no arithmetic (results are garbage).
template <int W, bool PACKED, bool FULLW, int MODE>
__global__ void __launch_bounds__(32 * (W + 1), 3) fuse2q_kernel(const __grid_constant__ Fuse2Args fa) {
  constexpr int V = 4;
  constexpr int TX = 128 * W;
  constexpr int RS = TX + 8;        // row stride per plane: columns x0-4 .. x0+TX+3 (cell j at index 4+j)
  constexpr bool JOINT = (MODE & 1) != 0, DRY = (MODE & 2) != 0;
  extern __shared__ __align__(128) float smem[];
  float* stage0 = smem;                              // [2][9][RS] step-t rows of the next TWO phase 1s (slot = row parity)
  float* ring_n = stage0 + 2 * NSPEEDS * RS;         // [1][3][RS] planes 4,7,8 of step t+1
  float* ring_m = ring_n + 1 * 3 * RS;               // [2][3][RS] planes 0,1,3
  float* ring_s = ring_m + 2 * 3 * RS;               // [3][3][RS] planes 2,5,6
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_s + 3 * 3 * RS);   // full[2], empty[2]: one pair per stage slot
  uint64_t* empty = full + 2;
  const float** ptab = reinterpret_cast<const float**>(full + 5);   // per plane: source of column x0-4 of local row 0
  uint32_t* mring = reinterpret_cast<uint32_t*>(ptab + NSPEEDS);    // [4][4*W] obstacle words of the strip, rows r & 3
  constexpr int MW = 4 * W;         // obstacle words per strip row
  float* wst0 = reinterpret_cast<float*>(mring + 4 * MW);           // [2][6][4] periodic x wrap of the outermost strips:
                                                                    // planes 1,5,8 at columns nx-4..nx-1, planes 3,6,7 at 0..3
  constexpr bool MTMA = FULLW;      // mask rows travel with the stage copies (needs 16-byte aligned strip starts)
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
  int ys, ye;
  f2_segment_rows(sy, fa.seg_rows, fa.seg_long, fa.n_long, rows, ys, ye);
  const int x0 = strip * TX;
  const int ncol = FULLW ? TX : min(TX, nx - x0);
  const int j0 = (warp * 32 + lane) * V;
  const int c = 4 + j0;                         // the thread's first column in stage / ring rows
  const int xb = x0 + j0;
  const bool active = !halo_warp && (FULLW || j0 < ncol);
  const bool touches_bottom = (ys < 2), touches_top = (ye >= rows - 1);
  if (!halo_warp && lane < 2) {   // the warp's Σ|u| accumulators of the two steps
    part_hi[lane][warp] = 0.0;
    part_lo[lane][warp] = 0.0;
  }
  if (threadIdx.x == 0) {
    mbar_init(full + 0, 1);
    mbar_init(full + 1, 1);
    mbar_init(empty + 0, W);
    mbar_init(empty + 1, W);
  }
  if (threadIdx.x < NSPEEDS) {   // which source row plane k is pulled from (kernels.cl:104-112)
    const int k = threadIdx.x;
    const int dy = (k == 2 || k == 5 || k == 6) ? -1 : (k == 4 || k == 7 || k == 8) ? 1 : 0;
    ptab[k] = a.src + k * ps + (long long)dy * a.pitch + (x0 - 4);
  }
  if (a.edge_count != nullptr && (touches_bottom || touches_top)) {
    if (threadIdx.x == 0) {
      if (touches_top) wait_epoch(a.flag_from_up, a.epoch - 1, a.error_word, a.wait_timeout_ns);
      if (touches_bottom) wait_epoch(a.flag_from_down, a.epoch - 1, a.error_word, a.wait_timeout_ns);
    }
  }
  __syncthreads();
  // ring: this block's ghost rows were written by the neighbour GPU (generic proxy, ordered by the
  // acquire above + the barrier); the bulk copies read them through the async proxy
  const bool ring_edge = a.edge_count != nullptr && (touches_bottom || touches_top);
  if (ring_edge) asm volatile("fence.proxy.async.global;" ::: "memory");

  const int accel_g = fa.ny - 2;

  const bool strip_first = (x0 == 0), strip_last = (x0 + ncol >= nx);
  // stage slot and mbarrier parity of row r: rows alternate between the two slots; a slot's k-th use has parity k & 1
  auto slot_of = [&](int r) { return (r - (ys - 1)) & 1; };
  auto parity_of = [&](int r) { return (uint32_t)(((r - (ys - 1)) >> 1) & 1); };
  auto issue_row = [&](int r) {   // one thread: the nine plane-rows phase 1 of row r needs -> its stage slot
    const int sl = slot_of(r);
    uint64_t* fb = full + sl;
    float* stage = stage0 + sl * NSPEEDS * RS;
    float* wst = wst0 + sl * 24;
    mbar_expect_tx(fb, NSPEEDS * RS * (uint32_t)sizeof(float) + (MTMA ? MW * (uint32_t)sizeof(uint32_t) : 0u) +
                             (strip_first ? 48u : 0u) + (strip_last ? 48u : 0u));
    const long long off = (long long)r * a.pitch;
    if (strip_first) {   // x-1 of column 0 is column nx-1 (kernels.cl:102): planes 1,5,8, columns nx-4..nx-1
      tma_load_1d(wst + 0, ptab[1] + off + nx, 16u, fb);
      tma_load_1d(wst + 4, ptab[5] + off + nx, 16u, fb);
      tma_load_1d(wst + 8, ptab[8] + off + nx, 16u, fb);
    }
    if (strip_last) {    // x+1 of column nx-1 is column 0 (kernels.cl:100-101): planes 3,6,7, columns 0..3
      tma_load_1d(wst + 12, ptab[3] + off + (4 - x0), 16u, fb);
      tma_load_1d(wst + 16, ptab[6] + off + (4 - x0), 16u, fb);
      tma_load_1d(wst + 20, ptab[7] + off + (4 - x0), 16u, fb);
    }
    // optional: pull the rows of iteration r + l2_ahead from HBM into L2 now, so that their bulk copies
    // hit L2 later — more bytes in flight at the DRAM than the one stage buffer per block allows
    const int rp = r + fa.l2_ahead;
    if (fa.l2_ahead > 0 && rp <= ye) {
      const long long poff = (long long)rp * a.pitch;
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) tma_prefetch_l2(ptab[k] + poff, RS * (uint32_t)sizeof(float));
    }
#pragma unroll
    for (int k = 0; k < NSPEEDS; k++) tma_load_1d(stage + k * RS, ptab[k] + off, RS * (uint32_t)sizeof(float), fb);
    if constexpr (MTMA)   // the strip's obstacle words of row r: read by phase 1 of row r and, an iteration later, by phase 2
      tma_load_1d(mring + (r & 3) * MW, a.mask + (long long)r * a.mask_pitch + (x0 >> 5), MW * (uint32_t)sizeof(uint32_t), fb);
  };
  // body warps, once they hold their share of the stage in registers: the halo warp (which has the
  // time) waits for all of them and requests the next row — nobody on the critical path waits or issues
  auto stage_consumed = [&](int r) {
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot_of(r));
  };
  // the block barrier, reached from two different loops (body warps / halo warp)
  auto block_sync = [&]() { asm volatile("bar.sync 0;" ::: "memory"); };

  // ring addressing: row r's planes 4,7,8 / 0,1,3 / 2,5,6 (index within the group: 0,1,2)
  // (one slot fewer per group than fuse2p_kernel: the second block barrier per row closes the window in which a
  // fast warp's next phase 1 could overwrite what a slow warp's phase 2 still reads)
  auto rn = [&](int r) { (void)r; return ring_n; };
  auto rm = [&](int r) { return ring_m + ((r + 2) & 1) * (3 * RS); };
  auto rs = [&](int r) { return ring_s + ((r + 3) % 3) * (3 * RS); };

  if (warp == 0 && elect_one()) {   // both stage slots: the first two rows
    issue_row(ys - 1);
    issue_row(ys);
  }

  if (halo_warp) {
    // =====================================================================================
    // halo warp: lanes 0 / 1 advance the column left / right of the strip by one step
    // (phase 1 only), from the stage, or from global memory where the column is the
    // periodic wrap (kernels.cl:100-102).  Same waits, counts and barriers as the body warps.
    // =====================================================================================
    const bool hl = lane < 2;
    const int xh = (lane == 0) ? ((x0 == 0) ? nx - 1 : x0 - 1) : ((x0 + ncol >= nx) ? 0 : x0 + ncol);
    const int xhw = (xh == 0) ? nx - 1 : xh - 1;
    const int xhe = (xh + 1 >= nx) ? 0 : xh + 1;
    const bool h_global = hl && ((lane == 0) ? (x0 == 0) : (x0 + ncol >= nx));   // wrapped column: not in the stage
    const int hidx = (lane == 0) ? 3 : 4 + ncol;                               // its index in stage / ring rows
    float ht[NSPEEDS];
    uint32_t hmask = 0;
#pragma unroll
    for (int k = 0; k < NSPEEDS; k++) ht[k] = 0.0f;
    auto halo_prefetch = [&](int r) {   // global loads, one row ahead: obstacle bit (+ the wrapped column)
      if (hl) hmask = __ldg(a.mask + (long long)r * a.mask_pitch + (xh >> 5));
      if (h_global) {
        const float* s_mid = a.src + (long long)r * a.pitch;
        const float* s_south = s_mid - a.pitch;
        const float* s_north = s_mid + a.pitch;
        // (ghost rows of a ring are rewritten during the kernel: coherent L2 loads there, never ld.global.nc)
        const float* q[NSPEEDS] = {s_mid + 0 * ps + xh,   s_mid + 1 * ps + xhw,   s_south + 2 * ps + xh,
                                   s_mid + 3 * ps + xhe,  s_north + 4 * ps + xh,  s_south + 5 * ps + xhw,
                                   s_south + 6 * ps + xhe, s_north + 7 * ps + xhe, s_north + 8 * ps + xhw};
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++) ht[k] = ring_edge ? __ldcg(q[k]) : __ldg(q[k]);
      }
    };
    auto halo_row = [&](int r) {
      const int sl = slot_of(r);
      const float* stage = stage0 + sl * NSPEEDS * RS;
      mbar_wait(full + sl, parity_of(r));
      if (hl && !h_global) {
        ht[0] = stage[0 * RS + hidx];
        ht[1] = stage[1 * RS + hidx - 1];
        ht[2] = stage[2 * RS + hidx];
        ht[3] = stage[3 * RS + hidx + 1];
        ht[4] = stage[4 * RS + hidx];
        ht[5] = stage[5 * RS + hidx - 1];
        ht[6] = stage[6 * RS + hidx + 1];
        ht[7] = stage[7 * RS + hidx + 1];
        ht[8] = stage[8 * RS + hidx - 1];
      }
      const bool accel = (global_row(r, fa.y0, fa.ny) == accel_g);
      const bool hfluid = ((hmask >> (xh & 31)) & 1u) == 0u;
      float o[NSPEEDS];
      if (hl) {
        collide_cell(ht, hfluid, a.omega, o);   // (consumes the stage values: they are in registers from here on)
        if (accel) accelerate_cell(o, hfluid, a.w1, a.w2);
      }
      __syncwarp();
      if (r + 2 <= ye) {   // once the body warps have read their share of this slot: request the row after the next
        mbar_wait(empty + sl, parity_of(r));
        if (elect_one()) issue_row(r + 2);
      }
      if (hl) {
        float* n = rn(r) + hidx;
        float* m = rm(r) + hidx;
        float* so = rs(r) + hidx;
        m[0 * RS] = o[0]; m[1 * RS] = o[1]; m[2 * RS] = o[3];
        so[0 * RS] = o[2]; so[1 * RS] = o[5]; so[2 * RS] = o[6];
        n[0 * RS] = o[4]; n[1 * RS] = o[7]; n[2 * RS] = o[8];
      }
    };
    halo_prefetch(ys - 1);
    for (int r = ys - 1; r <= ys; r++) {
      halo_row(r);
      halo_prefetch(r + 1);
    }
    for (int y = ys; y < ye; y++) {
      halo_row(y + 1);
      block_sync();                             // rows y-1, y, y+1 of step t+1 are in the ring
      if (y + 1 < ye) halo_prefetch(y + 2);     // lands while the body warps run phase 2
      block_sync();                             // phase 2 of row y has read the ring
    }
  } else {
    // =====================================================================================
    // body warps
    // =====================================================================================
    const bool need_r = active && (lane == 31 || (!FULLW && j0 + V >= ncol));
    const bool wrap_l = active && lane == 0 && xb == 0;   // x-1 wraps to nx-1: not in the stage
    const bool wrap_r = need_r && xb + V >= nx;           // x+4 wraps to 0
    // obstacle bits of the thread's four cells in row r (bit 0 = the first cell): from the mask ring the
    // stage copies fill, or (ragged widths) straight from global memory
    auto row_bits = [&](int r) -> uint32_t {
      if constexpr (MTMA) return mring[(r & 3) * MW + (j0 >> 5)] >> (j0 & 31);
      else return active ? (__ldg(a.mask + (long long)r * a.mask_pitch + (xb >> 5)) >> (xb & 31)) : 0u;
    };

    auto lds4 = [&](const float* q, float (&v)[V]) {
      const float4 f = *reinterpret_cast<const float4*>(q);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    };
    auto st4 = [&](float* d, const float (&v)[V]) { *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]); };
    auto cells = [&](const float (&q)[NSPEEDS][V], float l1, float l5, float l8, float r3, float r6, float r7, uint32_t bits,
                     bool accel, float (&out)[NSPEEDS][V]) -> float {
      if constexpr (DRY) {
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++)
#pragma unroll
          for (int j = 0; j < V; j++) out[k][j] = q[k][j];
        return l1 + r3;
      }
      if constexpr (PACKED) return compute_quad<JOINT>(q, l1, l5, l8, r3, r6, r7, bits, a.omega, accel, a.w1, a.w2, out);
      else return compute_cells<V, false>(q, l1, l5, l8, r3, r6, r7, bits, a.omega, accel, a.w1, a.w2, out);
    };
    // Σ|u| of a row: fixed butterfly, the warp's fp32 sum added error-free to the warp's double-double of
    // step `st` — kept in shared memory (lane 0 updates it) so that no accumulator is carried in registers
    auto flush = [&](float& pend, int st) {
      float t = pend;
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) t = __fadd_rn(t, __shfl_xor_sync(FULL, t, s));
      if (lane == 0) {
        double hi = part_hi[st][warp], lo = part_lo[st][warp];
        dd_add(hi, lo, (double)t, 0.0);
        part_hi[st][warp] = hi;
        part_lo[st][warp] = lo;
      }
      pend = 0.0f;
    };

    // ---- phase 1: step t -> t+1 of row r: stage -> registers, count out of the stage, collide, ring stores ----
    auto phase1 = [&](int r) -> float {
      const int sl = slot_of(r);
      const float* stage = stage0 + sl * NSPEEDS * RS;
      const float* wst = wst0 + sl * 24;
      mbar_wait(full + sl, parity_of(r));    // the slot holds row r's inputs
      // reads are unconditional: every index lies inside the row buffers, and what inactive
      // lanes (columns beyond a ragged strip) compute from it is never stored or summed
      float q[NSPEEDS][V];
#pragma unroll
      for (int k = 0; k < NSPEEDS; k++) lds4(stage + k * RS + c, q[k]);
      float f1 = 0.f, f5 = 0.f, f8 = 0.f, f3 = 0.f, f6 = 0.f, f7 = 0.f;
      if (lane == 0) { f1 = stage[1 * RS + c - 1]; f5 = stage[5 * RS + c - 1]; f8 = stage[8 * RS + c - 1]; }
      if (need_r) { f3 = stage[3 * RS + c + V]; f6 = stage[6 * RS + c + V]; f7 = stage[7 * RS + c + V]; }
      if (wrap_l) { f1 = wst[3]; f5 = wst[7]; f8 = wst[11]; }      // the stage holds the wrong row there
      if (wrap_r) { f3 = wst[12]; f6 = wst[16]; f7 = wst[20]; }
      const uint32_t bits = row_bits(r);
      stage_consumed(r);                     // the halo warp requests row r+2's copies into this slot from here on

      const bool accel = (global_row(r, fa.y0, fa.ny) == accel_g);   // the second step always follows
      float l1 = __shfl_up_sync(FULL, q[1][V - 1], 1);
      float l5 = __shfl_up_sync(FULL, q[5][V - 1], 1);
      float l8 = __shfl_up_sync(FULL, q[8][V - 1], 1);
      float r3 = __shfl_down_sync(FULL, q[3][0], 1);
      float r6 = __shfl_down_sync(FULL, q[6][0], 1);
      float r7 = __shfl_down_sync(FULL, q[7][0], 1);
      if (lane == 0) { l1 = f1; l5 = f5; l8 = f8; }
      if (need_r) { r3 = f3; r6 = f6; r7 = f7; }
      float out[NSPEEDS][V];
      float tot = cells(q, l1, l5, l8, r3, r6, r7, bits, accel, out);
      if (active) {
        float* n = rn(r) + c;
        float* m = rm(r) + c;
        float* so = rs(r) + c;
        st4(m + 0 * RS, out[0]); st4(m + 1 * RS, out[1]); st4(m + 2 * RS, out[3]);
        st4(so + 0 * RS, out[2]); st4(so + 1 * RS, out[5]); st4(so + 2 * RS, out[6]);
        st4(n + 0 * RS, out[4]); st4(n + 1 * RS, out[7]); st4(n + 2 * RS, out[8]);
      } else {
        tot = 0.0f;
      }
      return tot;
    };

    // ---- phase 2: step t+1 -> t+2 of row y, inputs from the ring; returns the thread's Σ|u| ----
    auto phase2 = [&](int y) -> float {
      const float* m = rm(y) + c;        // planes 0,1,3 of row y
      const float* so = rs(y - 1) + c;   // planes 2,5,6 of row y-1
      const float* n = rn(y + 1) + c;    // planes 4,7,8 of row y+1
      float g[NSPEEDS][V];
      lds4(m + 0 * RS, g[0]); lds4(m + 1 * RS, g[1]); lds4(m + 2 * RS, g[3]);
      lds4(so + 0 * RS, g[2]); lds4(so + 1 * RS, g[5]); lds4(so + 2 * RS, g[6]);
      lds4(n + 0 * RS, g[4]); lds4(n + 1 * RS, g[7]); lds4(n + 2 * RS, g[8]);
      float e1 = 0.f, e5 = 0.f, e8 = 0.f, e3 = 0.f, e6 = 0.f, e7 = 0.f;
      if (lane == 0) { e1 = m[1 * RS - 1]; e5 = so[1 * RS - 1]; e8 = n[2 * RS - 1]; }
      if (need_r) { e3 = m[2 * RS + V]; e6 = so[2 * RS + V]; e7 = n[1 * RS + V]; }
      const uint32_t bits = row_bits(y);
      float l1 = __shfl_up_sync(FULL, g[1][V - 1], 1);
      float l5 = __shfl_up_sync(FULL, g[5][V - 1], 1);
      float l8 = __shfl_up_sync(FULL, g[8][V - 1], 1);
      float r3 = __shfl_down_sync(FULL, g[3][0], 1);
      float r6 = __shfl_down_sync(FULL, g[6][0], 1);
      float r7 = __shfl_down_sync(FULL, g[7][0], 1);
      if (lane == 0) { l1 = e1; l5 = e5; l8 = e8; }
      if (need_r) { r3 = e3; r6 = e6; r7 = e7; }

      const bool accel = !fa.last && (global_row(y, fa.y0, fa.ny) == accel_g);
      float out[NSPEEDS][V];
      float tot = cells(g, l1, l5, l8, r3, r6, r7, bits, accel, out);
      if (active) {
        float* d = a.dst + (long long)y * a.pitch + xb;
#pragma unroll
        for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(d + k * ps, out[k]);
        if (y >= rows - 2) {   // the up neighbour's ghost rows -1, -2
          float* gh = a.up_ghost + (long long)(y - (rows - 1)) * a.pitch + xb;
#pragma unroll
          for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(gh + k * a.up_plane_stride, out[k]);
        }
        if (y < 2) {           // the down neighbour's ghost rows rows, rows+1
          float* gh = a.down_ghost + (long long)y * a.pitch + xb;
#pragma unroll
          for (int k = 0; k < NSPEEDS; k++) store_vec<V, 0>(gh + k * a.down_plane_stride, out[k]);
        }
      } else {
        tot = 0.0f;
      }
      return tot;
    };

    // ---- prologue: rows ys-1 and ys of step t+1 into the ring; row ys+1's inputs requested ----
    float pend1 = 0.0f, pend2 = 0.0f;    // row sums whose butterfly is still to be done
    for (int r = ys - 1; r <= ys; r++) {
      const float t = phase1(r);
      if (r == ys) {   // row ys-1 belongs to the segment below (or is a ghost row)
        pend1 = t;
        flush(pend1, 0);
      }
    }

    // ---- main loop: phase 1 of row y+1, phase 2 of row y ----
    for (int y = ys; y < ye; y++) {
      const bool more = (y + 1 < ye);
      flush(pend2, 1);                // Σ|u| of the previous iteration's phase 2
      const float t = phase1(y + 1);
      pend1 = more ? t : 0.0f;               // row ye belongs to the segment above (or is a ghost row)
      block_sync();                          // rows y-1, y, y+1 of step t+1 are in the ring
      flush(pend1, 0);
      pend2 = phase2(y);
      block_sync();                          // every warp has read the ring: the next phase 1 may overwrite its oldest rows
    }
    flush(pend2, 1);
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

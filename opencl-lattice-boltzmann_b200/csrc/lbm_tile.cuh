// lbm_tile.cuh — the small-deck path: K time steps per hand-off, lattice tiles in shared memory.
//
// The reference's three small decks (128x128, 128x256, 256x256: 16 K - 64 K cells, 0.6 - 2.3 MB per
// lattice) are not bound by bandwidth or arithmetic on a B200 — one time step of all of 128x128 is
// ~50 ns of issue slots spread over 148 SMs — but by LATENCY: the host loop d2q9-bgk.c:221-238 costs
// two kernel launches per step (17 us per step on the K20m of d2q9-bgk.out), and round 1's persistent
// kernel, which moved that loop onto the device, still pays one flag hand-off between neighbouring
// blocks plus an L2 round trip per step (2.6 us).  This kernel pays the hand-off once per K steps:
//
//   * the lattice is cut into tiles_x x tiles_y tiles, one thread block per tile, all blocks
//     co-resident (cooperative launch), each block keeps its tile for the whole launch;
//   * a round = K time steps.  The block loads its tile plus a K-cell halo on every side
//     (periodic, kernels.cl:91-102) from the current global lattice into shared memory, then
//     advances it K times entirely on chip — step i is valid on the cells at least i cells inside
//     the haloed tile, so after K steps exactly the tile itself is left (trapezoid / overlapped
//     tiling: the halo is recomputed redundantly by the neighbours instead of being exchanged every
//     step), ping-ponging between two shared-memory copies with one __syncthreads per step;
//   * the last step of a round stores the tile to the OTHER global lattice, then the block publishes
//     its round counter (release) and, before loading the next round's halo, waits for the counters of
//     its eight neighbouring tiles (acquire).  "Neighbours finished round r-1" covers both the
//     read-after-write on the halo cells and the write-after-read on the buffer written in round r;
//   * one thread per cell of the haloed tile, scalar arithmetic = collide_cell (lbm_kernels.cuh), the
//     same correctly rounded operations in the same order as every other kernel and the CPU oracle:
//     the lattice is bit-identical however it is tiled and however many steps a round takes;
//   * accelerate_flow (kernels.cl:9-53) is folded into the stores of row ny-2 of every step but the
//     run's last, halo copies included; step 0's is the host's pre-pass as for every kernel;
//   * av_vels (kernels.cl:198-229, :234-290): each owned cell's speed of each step is parked in shared
//     memory and summed in double-double by otherwise idle warps while the block waits for its
//     neighbours — one (hi, lo) per tile per step, off the critical path, order-independent.
//
// Replaces K iterations of the reference's host loop (K x accelerate_flow + K x timestep + the
// per-step part of reduce) per hand-off.
#pragma once

#include "lbm_kernels.cuh"

namespace lbm {

struct TileArgs {
  float* buf[2];             // row 0 of plane 0 of the two lattice buffers
  long long plane_stride;    // floats between planes
  int pitch, nx, ny;
  const uint32_t* mask;      // bit (x & 31) of word [y*mask_pitch + (x >> 5)]: 1 = blocked
  int mask_pitch;
  float omega, w1, w2;
  int tiles_x, tiles_y;      // block b = tile (b % tiles_x, b / tiles_x)
  int K;                     // time steps per round = halo depth; <= the smallest tile's width and height
  int lw, lh;                // haloed size of the LARGEST tile: blockDim.x >= lw*lh; row stride of the smem copies = lw
  int nsteps;                // time steps in this launch
  int first_buf;             // buffer that holds the state when the launch starts
  int accel_row;             // global row ny-2 whose stores get the next step's accelerate_flow, or -1
  int skip_last_accel;       // 1: the launch's last step is the run's last step
  unsigned int* progress;    // [tiles*32] rounds completed by each tile in this launch (one 128-byte line each; zeroed)
  double2* partials;         // [nsteps][tiles] per-tile sum of cell speeds of each step, as (hi, lo)
  long long* timing;         // optional (development): [TILE_TIMING_ROUNDS][TILE_TIMING_SLOTS] SM clocks of tile 0, thread 0
};

constexpr int TILE_TIMING_ROUNDS = 64, TILE_TIMING_SLOTS = 16;

// x range [x0, x0+w) of part i of n of an axis of `len` cells (sizes differ by at most one)
__host__ __device__ inline void tile_range(int len, int n, int i, int& x0, int& w) {
  const int base = len / n, rem = len % n;
  w = base + (i < rem ? 1 : 0);
  x0 = i * base + (i < rem ? i : rem);
}

// A block has at most 1024 threads = haloed cells (times the cells per thread), so every shared-memory plane gets the
// same fixed stride:
// plane offsets are then immediates of the LDS / STS instructions instead of per-step address arithmetic
// (the steps are bound by issue slots, ~1 instruction per cycle and scheduler: every instruction counts).
constexpr int TILE_PLANE = 1024;

__host__ __device__ inline int tile_smem_bytes(int cells_per_thread, int K, int max_owned) {
  // two copies of nine planes + K steps of owned-cell speeds
  return (2 * NSPEEDS * TILE_PLANE * cells_per_thread + K * max_owned) * (int)sizeof(float);
}

__device__ __forceinline__ int wrap(int v, int n) {   // v in [-n, 2n)
  return v < 0 ? v + n : (v >= n ? v - n : v);
}

// CPT = cells per thread (1 or 2): a block of at most 1024 threads steps a haloed tile of up to CPT * 1024 cells.
// CPT = 1 is the kernel of the reference's three small decks; CPT = 2 extends it to lattices of up to ~260 K cells
// (512 x 512), which would otherwise fall to the persistent kernel's one-hand-off-per-step (tools/sizes_bench.py).
template <int CPT>
__global__ void __launch_bounds__(1024, 1) tile_kernel(const __grid_constant__ TileArgs ta) {
  extern __shared__ __align__(16) float tsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int ntiles = ta.tiles_x * ta.tiles_y;
  const int tile = blockIdx.x;
  const int txi = tile % ta.tiles_x, tyi = tile / ta.tiles_x;
  const int K = ta.K, nx = ta.nx, ny = ta.ny;
  int x0, y0, w, h;
  tile_range(nx, ta.tiles_x, txi, x0, w);
  tile_range(ny, ta.tiles_y, tyi, y0, h);
  const int LW = ta.lw;                       // smem row stride (the largest tile's haloed width)
  const int lwt = w + 2 * K, lht = h + 2 * K; // this tile's haloed size
  constexpr int plane = TILE_PLANE * CPT;
  float* A = tsm;
  float* B = tsm + NSPEEDS * plane;
  float* speeds = tsm + 2 * NSPEEDS * plane;  // [K][max_owned]; this tile uses [K][w*h]
  const int nown = w * h;

  // ---- per-thread constants: the cells of the haloed tile this thread owns for the whole launch ----
  int c[CPT], margin[CPT], own_idx[CPT];       // index in a smem plane; cells from the halo's rim (-1: no cell)
  bool fluid[CPT], accel_cell[CPT], owned[CPT];
  long long goff[CPT], ghost_off[CPT];
#pragma unroll
  for (int j = 0; j < CPT; j++) {
    c[j] = tid + j * (int)blockDim.x;
    const int lx = c[j] % LW, ly = c[j] / LW;
    const bool in_tile = lx < lwt && ly < lht;
    margin[j] = in_tile ? min(min(lx, lwt - 1 - lx), min(ly, lht - 1 - ly)) : -1;
    const int gx = wrap((x0 - K + lx) % nx, nx), gy = wrap((y0 - K + ly) % ny, ny);
    owned[j] = margin[j] >= K;
    own_idx[j] = (ly - K) * w + (lx - K);
    fluid[j] = true;
    accel_cell[j] = false;
    goff[j] = 0;
    if (in_tile) {
      fluid[j] = ((ta.mask[(long long)gy * ta.mask_pitch + (gx >> 5)] >> (gx & 31)) & 1u) == 0u;
      accel_cell[j] = (gy == ta.accel_row);
      goff[j] = (long long)gy * ta.pitch + gx;
    }
    // ghost rows of the arena mirror the lattice's edge rows (other kernels read them): kept up to date by the
    // launch's last round.  Row y < 2 is also ghost row ny + y; row y >= ny-2 is also ghost row y - ny.
    ghost_off[j] = (gy < 2) ? (long long)ny * ta.pitch : (gy >= ny - 2) ? -(long long)ny * ta.pitch : 0;
  }

  // the eight neighbouring tiles (periodic); thread j < 8 polls neighbour j
  int nb_tile = tile;
  if (tid < 8) {
    const int j = tid < 4 ? tid : tid + 1;    // skip the centre of the 3x3
    const int ddx = j % 3 - 1, ddy = j / 3 - 1;
    nb_tile = wrap(tyi + ddy, ta.tiles_y) * ta.tiles_x + wrap(txi + ddx, ta.tiles_x);
  }

  const int rounds = (ta.nsteps + K - 1) / K;
  for (int r = 0; r < rounds; r++) {
    const int s0 = r * K;                              // first step of the round (index within the launch)
    const int k = min(K, ta.nsteps - s0);              // steps in this round
    const float* src = ta.buf[(ta.first_buf + r) & 1];
    float* dst = ta.buf[(ta.first_buf + r + 1) & 1];
    // development: phase clocks of tile 0 (slot 0 round start, 1 neighbours' flags seen, 2 halo in shared memory,
    // 3.. after each step's barrier, 15 own last step done)
    long long* tm = (ta.timing != nullptr && tile == 0 && tid == 0 && r >= 8 && r < 8 + TILE_TIMING_ROUNDS)
                        ? ta.timing + (r - 8) * TILE_TIMING_SLOTS : nullptr;
    if (tm) tm[0] = clock64();

    // ---- wait for the neighbours' previous round; meanwhile sum the previous round's speeds ----
    if (r > 0) {
      if (tid < 8) {
        while (ld_relaxed_gpu(ta.progress + 32 * nb_tile) < (unsigned)r) { }
        __threadfence();
      } else if (warp >= 1) {
        const int kp = K;                              // every round but the last has K steps
        for (int s = warp - 1; s < kp; s += nwarps - 1) {
          double hi = 0.0, lo = 0.0;
          for (int i = lane; i < nown; i += 32) dd_add(hi, lo, (double)speeds[s * nown + i], 0.0);
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) {
            const double oh = __shfl_xor_sync(FULL, hi, d), ol = __shfl_xor_sync(FULL, lo, d);
            dd_add(hi, lo, oh, ol);
          }
          if (lane == 0) ta.partials[(long long)(s0 - K + s) * ntiles + tile] = make_double2(hi, lo);
        }
      }
      __syncthreads();
    }
    if (tm) tm[1] = clock64();

    // ---- the haloed tile of the current state -> shared memory (coherent L2 loads: other SMs wrote it) ----
#pragma unroll
    for (int j = 0; j < CPT; j++)
      if (margin[j] >= 0) {
        const float* g = src + goff[j];
#pragma unroll
        for (int q = 0; q < NSPEEDS; q++) A[q * plane + c[j]] = __ldcg(g + q * ta.plane_stride);
      }
    __syncthreads();
    if (tm) tm[2] = clock64();

    // ---- k time steps on chip: k-1 steps from one shared-memory copy to the other, the last one to the lattice ----
    // (two separate code paths: the global addresses of the last step stay out of the shared-memory steps, whose
    // issue slots are what bounds the round)
    auto pull_collide = [&](const float* cur, int i, int cc, bool fl, bool ac, bool own, float (&o)[NSPEEDS]) -> float {
      float t[NSPEEDS];
      const float* pc = cur + cc;                    // own row; the row below / above at -LW / +LW
      const float* ps = pc - LW;
      const float* pn = pc + LW;
      t[0] = pc[0 * plane];                          // pull, kernels.cl:104-112
      t[1] = pc[1 * plane - 1];
      t[2] = ps[2 * plane];
      t[3] = pc[3 * plane + 1];
      t[4] = pn[4 * plane];
      t[5] = ps[5 * plane - 1];
      t[6] = ps[6 * plane + 1];
      t[7] = pn[7 * plane + 1];
      t[8] = pn[8 * plane - 1];
      const float sp = collide_cell(t, fl, ta.omega, o, own);   // only the tile's own cells are summed
      if (ac && !(ta.skip_last_accel && s0 + i == ta.nsteps)) accelerate_cell(o, fl, ta.w1, ta.w2);
      return sp;
    };
    float* cur = A;
    float* nxt = B;
    float* spd = speeds;                             // the slots of step i-1
    for (int i = 1; i < k; i++) {
#pragma unroll
      for (int j = 0; j < CPT; j++)
        if (margin[j] >= i) {
          float o[NSPEEDS];
          const float sp = pull_collide(cur, i, c[j], fluid[j], accel_cell[j], owned[j], o);
          if (owned[j]) spd[own_idx[j]] = sp;
#pragma unroll
          for (int q = 0; q < NSPEEDS; q++) nxt[c[j] + q * plane] = o[q];
        }
      spd += nown;
      __syncthreads();
      if (tm && 2 + i < 15) tm[2 + i] = clock64();
      float* sw = cur; cur = nxt; nxt = sw;
    }
#pragma unroll
    for (int j = 0; j < CPT; j++)
      if (owned[j]) {
        float o[NSPEEDS];
        spd[own_idx[j]] = pull_collide(cur, k, c[j], fluid[j], accel_cell[j], true, o);
        float* g = dst + goff[j];
#pragma unroll
        for (int q = 0; q < NSPEEDS; q++) g[q * ta.plane_stride] = o[q];
        if (r == rounds - 1 && ghost_off[j] != 0) {
#pragma unroll
          for (int q = 0; q < NSPEEDS; q++) g[q * ta.plane_stride + ghost_off[j]] = o[q];
        }
      }
    if (tm) tm[15] = clock64();
    __syncthreads();
    if (tm && 2 + k < 15) tm[2 + k] = clock64();

    // ---- publish: this tile's state after round r is in the other lattice ----
    if (tid == 0) {
      __threadfence();
      *reinterpret_cast<volatile unsigned int*>(ta.progress + 32 * tile) = (unsigned)(r + 1);
    }
  }

  // ---- speeds of the last round ----
  {
    const int s0 = (rounds - 1) * K, k = ta.nsteps - s0;
    for (int s = warp; s < k; s += nwarps) {
      double hi = 0.0, lo = 0.0;
      for (int i = lane; i < nown; i += 32) dd_add(hi, lo, (double)speeds[s * nown + i], 0.0);
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const double oh = __shfl_xor_sync(FULL, hi, d), ol = __shfl_xor_sync(FULL, lo, d);
        dd_add(hi, lo, oh, ol);
      }
      if (lane == 0) ta.partials[(long long)(s0 + s) * ntiles + tile] = make_double2(hi, lo);
    }
  }
}

}  // namespace lbm

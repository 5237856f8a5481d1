// lbm_cuda.cu — the C-ABI of include/lbm.h over the sm_100a kernels.
//
// Replaces the device half of the reference host program: t_ocl and its buffers
// (d2q9-bgk.c:97-119, :687-710), the uploads (:200-209), the time loop's three
// launchers (:221-238, :282-393), clFinish (:239), the downloads (:251-260) and
// the releases (:729-741).  No OpenCL, no JIT: kernels are compiled ahead of time
// for sm_100a and physics constants travel as kernel arguments at full precision
// (the reference bakes 6-decimal -D constants into a run-time build, :643-645).
#include "lbm.h"
#include "lbm_kernels.cuh"
#include "lbm_fuse2.cuh"
#include "lbm_fuse2p.cuh"
#include "lbm_fuse2q.cuh"
#include "lbm_tile.cuh"

#include <cuda_runtime.h>

#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return 1;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("CUDA error during '%s' on line %d: %s", #call, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

// Every exported call leaves the caller's current CUDA device as it found it.
struct DeviceGuard {
  int dev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      (void)cudaGetLastError();
      dev = -1;
    }
  }
  ~DeviceGuard() {
    if (dev >= 0 && cudaSetDevice(dev) != cudaSuccess) (void)cudaGetLastError();
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

constexpr unsigned long long BLOB_MAGIC = 0x4c424d4232303032ULL;  // "LBMB2002"

// Geometry of one slab's lattice arena, enough for a neighbour to address its
// ghost rows, flags and obstacle mask: [buffer 0 | buffer 1 | flags | mask].
// Every plane has GHOST ghost rows below row 0 and GHOST above row rows-1; the
// mask has one ghost row on each side.
constexpr int GHOST = 2;
struct ArenaLayout {
  long long plane_stride;  // floats = (rows + 2*GHOST) * pitch
  long long buf_floats;    // 9 * plane_stride
  long long flags_offset;  // bytes from arena base
  long long mask_offset;   // bytes from arena base to mask ghost row -1
  int rows;
  int pitch;
};

// What every rank of a ring must agree on before the first launch: the kernels wait on each
// other's epoch flags, so a rank that picked another kernel (or another grid) would leave its
// neighbours spinning.  Travels in the export blob; lbm_connect compares it with its own.
struct RingPlan {
  int nx, ny, nranks;
  int fuse2, f2_kernel, f2_rows, f2_long, V;
};

struct Blob {
  unsigned long long magic;
  cudaIpcMemHandle_t handle;
  ArenaLayout layout;
  int device;
  int rank;
  RingPlan plan;
};

struct Neighbour {
  float* arena = nullptr;  // peer-addressable base of the neighbour's arena
  ArenaLayout layout{};
  bool ipc = false;        // arena came from cudaIpcOpenMemHandle
};

struct Slab {
  int device = 0;
  int y0 = 0, rows = 0;    // global rows [y0, y0 + rows)
  cudaStream_t stream = nullptr;
  float* arena = nullptr;
  ArenaLayout layout{};
  uint32_t* mask = nullptr;      // row 0 of the obstacle bit mask (inside the arena; ghost rows -1 and rows)
  int* stage = nullptr;          // upload staging buffer for the obstacle ints
  float* fs_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // output-stage staging (u_x, u_y, |u|, pressure), kept
  long long fs_capacity = 0;     // floats per fs_stage buffer
  double2* partials = nullptr;   // chunk_steps x blocks_per_step block partials
  double2* scratch = nullptr;    // chunk_steps x splits range sums of av_finalize_kernel
  unsigned int* tickets = nullptr;
  long long* tile_timing = nullptr;        // tile_kernel, option tile_debug: phase clocks of tile 0
  unsigned int* progress = nullptr;        // per-block step counters of the persistent kernel
  long long progress_capacity = 0;
  long long pblocks = 0;                   // grid of the persistent kernel
  int rows_per_block = 1;
  long long partial_capacity = 0;          // double2 entries in `partials`
  long long blocks = 0;          // step-kernel blocks = partials per step
  int splits = 1;
  double* av_hi = nullptr;
  double* av_lo = nullptr;
  long long av_capacity = 0;
  Neighbour up, down;
  unsigned long long edge_expected = 0;      // bottom-edge completions counted so far on this slab (edge_target of the next launch)
  unsigned long long edge_expected_top = 0;  // same for the top edge
  int f2_strips = 1, f2_segs_y = 1;      // tiling of the two-step kernel
  int f2_n_long = 0;                     // fuse2p_kernel: leading segments of f2_long rows (the rest have f2_rows)
  long long pstride = 0;                 // partial entries per step = max(step-kernel blocks, two-step blocks)
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;

  float* row0(int buf) const { return arena + (long long)buf * layout.buf_floats + (long long)GHOST * layout.pitch; }
  // [0] flag_from_up, [1] flag_from_down, [2] bottom-edge count, [3] top-edge count, [4] error word
  unsigned long long* flags() const {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(arena) + layout.flags_offset);
  }
};

float* nb_ghost_below(const Neighbour& n, int buf) {  // its ghost row -1 (row -2 is one pitch lower)
  return n.arena + (long long)buf * n.layout.buf_floats + (long long)(GHOST - 1) * n.layout.pitch;
}
float* nb_ghost_above(const Neighbour& n, int buf) {  // its ghost row `rows` (row rows+1 is one pitch higher)
  return n.arena + (long long)buf * n.layout.buf_floats + (long long)(n.layout.rows + GHOST) * n.layout.pitch;
}
uint32_t* nb_mask_ghost_below(const Neighbour& n) {   // its mask ghost row -1
  return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(n.arena) + n.layout.mask_offset);
}
uint32_t* nb_mask_ghost_above(const Neighbour& n) {   // its mask ghost row `rows`
  return nb_mask_ghost_below(n) + (long long)(n.layout.rows + 1) * (n.layout.pitch / 32);
}
unsigned long long* nb_flags(const Neighbour& n) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(n.arena) + n.layout.flags_offset);
}

}  // namespace

struct lbm_ctx {
  lbm_params p{};
  std::vector<Slab> slabs;
  std::vector<std::pair<int, cudaStream_t>> streams;  // one per distinct device
  int rank = 0, nranks = 1;
  int y0 = 0, rows = 0;        // rows held by this context
  int pitch = 0, mask_pitch = 0;
  bool ring = false;           // more than one slab in the whole ring -> flag protocol
  bool connected = false;
  bool uploaded = false;
  bool failed = false;         // a ring wait timed out: the state is garbage, every later call fails
  long wait_timeout_ms = 20000;
  int cur = 0;                 // buffer holding the current state
  unsigned long long epoch = 0;
  long long steps_done = 0;    // since creation
  long long steps_since_upload = 0;
  long long launches = 0;
  // options
  int opt_v = 0, opt_tpb = 0, opt_streaming = -1, opt_persistent = -1, opt_chunk = 0, opt_sync = 0, opt_tps = 0, opt_packed = -1,
      opt_tile_debug = 0, opt_tile = -1, opt_tile_steps = 0, opt_tile_w = 0, opt_tile_h = 0,
      opt_fuse2 = -1, opt_f2_rows = 0, opt_f2_tma = 3, opt_f2_l2ahead = 0, opt_f2_mode = 1, opt_f2_long = -1, opt_f2_nlong = -1;
  // resolved
  int fuse2 = 0, f2_warps = 4, f2_rows = 256, f2_long = 0;
  int f2_kernel = 3;           // 3: fuse2q_kernel (two-deep stage, packed arithmetic), 2: fuse2p_kernel (its A/B predecessor)
  int V = 1, tpb = 256, tps = 1024, packed = 0, streaming = 0, chunk_steps = 1, segs = 1, persistent = 0;
  // the multi-step tile kernel (lbm_tile.cuh): tiling of the lattice, steps per hand-off, block size
  int tile_cpt = 1;            // cells per thread of the tile kernel (1: up to 1024 haloed cells per tile, 2: up to 2048)
  int tile = 0, tile_K = 4, tiles_x = 1, tiles_y = 1, tile_lw = 0, tile_lh = 0, tile_threads = 0, tile_smem = 0;
  long long per_step = 0;      // largest slab's partials per step (slab i has rows_i * segs)
  float w1 = 0.f, w2 = 0.f;
};

namespace {

int set_device(const Slab& s) {
  CK(cudaSetDevice(s.device));
  return 0;
}

cudaStream_t stream_for(lbm_ctx* ctx, int device, int* err) {
  *err = 0;
  for (auto& ds : ctx->streams)
    if (ds.first == device) return ds.second;
  cudaStream_t st = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
    *err = fail("cannot create a stream on device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  ctx->streams.emplace_back(device, st);
  return st;
}

int validate(const lbm_params* p) {
  if (!p) return fail("params is NULL");
  if (p->nx < 1 || p->ny < 2) return fail("grid %dx%d is too small (need nx >= 1, ny >= 2)", p->nx, p->ny);
  return 0;
}

// Tiling for tile_kernel: at most one tile per SM, every haloed tile within one thread block
// ((w + 2K)(h + 2K) <= 1024 threads), K <= the smallest tile's sides; among those the tiling with the
// least work per round (the sum over the K steps of the shrinking haloed tile).  False: the lattice is
// too large for this kernel.
bool plan_tiles(lbm_ctx* ctx) {
  const int nx = ctx->p.nx, ny = ctx->p.ny;
  int sms = 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->slabs[0].device) != cudaSuccess) {
    (void)cudaGetLastError();
    sms = 148;
  }
  // Time model of a round, fitted to the phase clocks of tools/tile_timing.py (profiles/r02_tile_experiments),
  // in us: a hand-off through L2 (store, fence, flag, poll, halo load) ~2.0, plus per step ~0.30 + 0.010 per
  // warp of the haloed tile, or 0.022 per warp once the SM's issue slots are the bound (~160 instructions per cell).
  double best = -1.0;
  int best_tx = 0, best_ty = 0, best_k = 0;
  int best_cpt = 1;
  auto consider = [&](int tx, int ty, int want_k, int cpt) {
    if (tx < 1 || ty < 1 || tx > nx || ty > ny || (long long)tx * ty > sms) return;
    const int w = (nx + tx - 1) / tx, h = (ny + ty - 1) / ty;         // largest tile
    const int k = std::max(1, std::min(want_k, std::min(nx / tx, ny / ty)));   // <= smallest tile's sides
    // one step per hand-off is what the persistent kernel does too, with less redundancy (512^2: 51.3 vs 47.8 GLUPS):
    // unless the tile kernel is forced, a plan needs at least two steps per round
    if (k < 2 && ctx->opt_tile != 1) return;
    const long long cells = (long long)(w + 2 * k) * (h + 2 * k);
    if (cells > 1024LL * cpt) return;
    if (lbm::tile_smem_bytes(cpt, k, w * h) > 200 * 1024) return;
    const double warps = cells / 32.0;
    double per_step = (2.0 + k * std::max(0.30 + 0.010 * warps, 0.022 * warps)) / k;
    if (nx % tx != 0 || ny % ty != 0) per_step *= 1.03;               // ragged tilings: the largest tile sets the pace
    if (best < 0 || per_step < best - 1e-9 || (per_step < best + 1e-9 && tx < best_tx)) {
      best = per_step; best_tx = tx; best_ty = ty; best_k = k; best_cpt = cpt;
    }
  };
  const int k_lo = ctx->opt_tile_steps > 0 ? ctx->opt_tile_steps : 1;
  const int k_hi = ctx->opt_tile_steps > 0 ? ctx->opt_tile_steps : 8;
  for (int cpt = 1; cpt <= 2; cpt++) {
    for (int k = k_lo; k <= k_hi; k++) {
      if (ctx->opt_tile_w > 0 && ctx->opt_tile_h > 0)
        consider((nx + ctx->opt_tile_w - 1) / ctx->opt_tile_w, (ny + ctx->opt_tile_h - 1) / ctx->opt_tile_h, k, cpt);
      else
        for (int ty = 1; ty <= std::min(ny, sms); ty++)
          for (int tx = 1; tx <= std::min(nx, sms / ty); tx++) consider(tx, ty, k, cpt);
    }
  }
  if (best < 0) return false;
  ctx->tiles_x = best_tx;
  ctx->tiles_y = best_ty;
  ctx->tile_K = best_k;
  const int w = (nx + best_tx - 1) / best_tx, h = (ny + best_ty - 1) / best_ty;
  ctx->tile_lw = w + 2 * best_k;
  ctx->tile_lh = h + 2 * best_k;
  ctx->tile_cpt = best_cpt;
  ctx->tile_threads = std::max(64, ((ctx->tile_lw * ctx->tile_lh + best_cpt - 1) / best_cpt + 31) / 32 * 32);
  ctx->tile_smem = lbm::tile_smem_bytes(best_cpt, best_k, w * h);
  return true;
}

void resolve_options(lbm_ctx* ctx) {
  const int nx = ctx->p.nx;
  // L2-resident single-slab lattices run many steps per (cooperative) launch
  const double lattice_bytes = 2.0 * 9.0 * 4.0 * (double)ctx->pitch * (double)(ctx->rows + 2 * GHOST);
  const bool can_persist = ctx->slabs.size() == 1 && ctx->nranks == 1;
  // lattices small enough for the SMs' shared memory: K steps per hand-off on tiles (lbm_tile.cuh).
  // Automatic only while "persistent" is automatic too (persistent = 0 / 1 ask for the other kernels).
  ctx->tile = 0;
  if (can_persist && ctx->p.ny >= 4 && (ctx->opt_tile >= 0 ? ctx->opt_tile != 0 : ctx->opt_persistent < 0))
    ctx->tile = plan_tiles(ctx) ? 1 : 0;
  ctx->persistent = !ctx->tile && can_persist && (ctx->opt_persistent >= 0 ? ctx->opt_persistent != 0
                                                                           : lattice_bytes <= 96.0 * 1024 * 1024);
  int V = ctx->opt_v;
  if (V != 1 && V != 2 && V != 4) {
    V = 4;
    // latency-bound small grids: fewer cells per thread until ~1024 warps are in flight
    // (measured on the check decks: 128x128 -> 1, 256x256 -> 2, 1024x1024 -> 4)
    if (ctx->persistent)
      while (V > 1 && (long long)nx * ctx->rows / (32 * V) < 1024) V >>= 1;
  }
  while (V > 1 && nx % V != 0) V >>= 1;
  ctx->V = V;
  int tpb = ctx->opt_tpb;
  if (tpb != 128 && tpb != 256 && tpb != 512) tpb = ctx->persistent ? 128 : 256;
  ctx->tpb = tpb;
  ctx->segs = (nx + 32 * V - 1) / (32 * V);
  // load path of the one-step kernel (lbm_kernels.cuh): ld.global.nc, or coherent ld.global.cg where ghost rows
  // are rewritten by a neighbour while the kernel runs (every ring of several slabs) or on request
  ctx->streaming = (ctx->opt_streaming == 1 || ctx->ring) ? 1 : 0;
  ctx->tps = (ctx->opt_tps == 768) ? 768 : 1024;
  ctx->packed = (ctx->V > 1) && (ctx->opt_packed >= 0 ? ctx->opt_packed != 0 : 0);   // refined below for fuse2
  long long per_step = 0;
  const int wpb = tpb / 32;
  for (auto& s : ctx->slabs) {
    s.blocks = ((long long)s.rows * ctx->segs + wpb - 1) / wpb;
    s.splits = (int)std::max(1LL, std::min(64LL, s.blocks / 2048));
    per_step = std::max(per_step, s.blocks);
  }
  ctx->per_step = per_step;
  long long chunk = ctx->opt_chunk > 0 ? ctx->opt_chunk : (64LL << 20) / (16 * std::max(1LL, per_step));
  ctx->chunk_steps = (int)std::max(1LL, std::min(chunk, 4096LL));
  if (ctx->tile && ctx->opt_chunk <= 0) ctx->chunk_steps = 4096 / ctx->tile_K * ctx->tile_K;   // whole rounds per launch

  // two time steps per HBM pass (lbm_fuse2.cuh): 128-bit kernel only, lattices streamed from HBM,
  // every slab of the ring at least 4 rows (an even split, so every rank decides alike)
  ctx->f2_warps = 4;                                  // 512-column strips
  ctx->f2_kernel = ctx->opt_f2_tma == 2 ? 2 : 3;      // refined below: the two-deep-stage kernel exists for packed arithmetic only
  const long long total_slabs = (long long)ctx->nranks * (long long)ctx->slabs.size();
  // the smallest slab of the even split (the same number on every rank of a ring, so all decide alike)
  // and this context's own smallest slab (lbm_create_slab takes any row range; lbm_connect rejects a
  // ring whose ranks planned differently)
  long long min_rows = std::max(1LL, ctx->p.ny / total_slabs);
  for (auto& s : ctx->slabs) min_rows = std::min<long long>(min_rows, s.rows);
  const bool can_fuse = !ctx->persistent && !ctx->tile && V == 4 && nx >= 8 && min_rows >= 4;
  // rows per segment: every segment start recomputes two warm-up rows, so long segments are cheaper, but
  // the grid (strips x segments) should fill the ~444 resident blocks of a B200 a few times over
  {
    const long long strips = (nx + 128 * ctx->f2_warps - 1) / (128 * ctx->f2_warps);
    // large slabs, measured best (tools/f2_rows_sweep.py): about 2048 blocks, 32 or 64 rows per segment (refined
    // into long + short segments below)
    int seg = 64;
    while (seg > 16 && min_rows * strips / seg < 2048) seg >>= 1;
    if (seg == 16) {
      // mid-size slabs (fewer than ~64 Ki row-strips): the grid is only a few waves of the 3 x SMs resident
      // blocks, so what matters is that the LAST wave is full: the segment length with the least
      // waves x (rows + warm-up) (tools/sizes_bench.py: 2048^2 with 16-row segments = 512 blocks = 1.15 waves)
      int sms = 148;
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->slabs[0].device) != cudaSuccess) {
        (void)cudaGetLastError();
        sms = 148;
      }
      const long long slots = 3LL * sms;
      long long best_cost = -1;
      for (int cand = 8; cand <= 96 && cand <= min_rows; cand++) {
        const long long blocks = strips * ((min_rows + cand - 1) / cand);
        const long long cost = (blocks + slots - 1) / slots * (cand + 3);
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; seg = cand; }   // ties: the longer segment
      }
    }
    ctx->f2_rows = ctx->opt_f2_rows >= 4 ? ctx->opt_f2_rows : seg;
    // auto: on when the lattice is streamed from HBM and the grid can fill most of the GPU
    // (measured 128.4 vs 95 GLUPS at 16384^2; 104.7 vs 74.7 at 1536^2); smaller lattices keep the one-step kernel
    ctx->fuse2 = can_fuse && (ctx->opt_fuse2 >= 0 ? ctx->opt_fuse2 != 0 : min_rows * strips >= 4096);
  }
  if (ctx->fuse2) {
    // the two-step kernel is issue-bound, not HBM-bound: the packed fp32x2 arithmetic pays there
    if (ctx->opt_packed < 0) ctx->packed = 1;
    if (!ctx->packed || (ctx->opt_f2_mode & 3) == 0) ctx->f2_kernel = 2;   // (nor for the per-pair range check, fuse2_mode 0)
    if (ctx->packed && ctx->opt_tps != 1024) ctx->tps = 768;   // odd tail step: packed needs ~80 registers
    ctx->chunk_steps = std::max(2, ctx->chunk_steps);
    const int tx = 128 * ctx->f2_warps;
    // fuse2p_kernel, automatic tiling: every segment start recomputes two warm-up rows (long segments are
    // cheaper), but the launch ends when the LAST block ends (short segments leave a shorter tail), and blocks
    // are dispatched in index order: so 128-row segments first and 32-row ones for the rows that make up the
    // last ~two rounds of resident blocks (DESIGN.md has the measurements).
    ctx->f2_long = 0;
    for (auto& s : ctx->slabs) {
      s.f2_strips = (nx + tx - 1) / tx;
      s.f2_n_long = 0;
      s.f2_segs_y = (s.rows + ctx->f2_rows - 1) / ctx->f2_rows;
    }
    const bool forced = ctx->opt_f2_long > 0;                       // tests / sweeps: fuse2_long = rows of the long segments
    // automatic only where it was measured: slabs large enough for the 64-row uniform choice above (>= 4096 rows
    // of 32 strips: 128-row segments, then 32-row ones), and the class below it (64 Ki .. 128 Ki row-strips, e.g.
    // the 2048-row slab of the 8-GPU strong-scaling split): 64-row segments, then 16-row ones.  fuse2q_kernel,
    // tools/f2_rows_sweep.py: 16384x2048 164.3 GLUPS (128/32 with a quarter short: 157.4; uniform 32: 161.0),
    // 16384x3072 169.6 (157.9; 164.2), 8192x4096 163.6 (156.5; 160.2), 4096x8192 161.8 (155.0; 159.0)
    const bool auto64 = ctx->opt_f2_long < 0 && ctx->opt_f2_rows < 4 && ctx->f2_rows == 64;
    const bool auto32 = ctx->opt_f2_long < 0 && ctx->opt_f2_rows < 4 && ctx->f2_rows == 32;
    const bool automatic = auto64 || auto32;
    if (ctx->f2_kernel >= 2 && (forced || automatic)) {
      const int seg_short = forced ? ctx->f2_rows : auto32 ? 16 : 32;
      const int seg_long = forced ? ctx->opt_f2_long : auto32 ? 64 : 128;
      int sms = 148;
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->slabs[0].device) != cudaSuccess) {
        (void)cudaGetLastError();
        sms = 148;
      }
      // rows left to the short segments: ~two rounds of resident blocks (3 per SM); forced: a quarter of the slab
      auto rows_short_of = [&](const Slab& s) -> long long {
        const long long want = forced ? (s.rows + 3) / 4 : (2LL * 3 * sms + s.f2_strips - 1) / s.f2_strips * seg_short;
        return (want + seg_short - 1) / seg_short * seg_short;
      };
      bool ok = true;
      for (auto& s : ctx->slabs)
        if (forced ? seg_long >= s.rows : (seg_long * 4 > s.rows || rows_short_of(s) * 2 > s.rows)) ok = false;   // small slab: stay uniform
      if (ok) {
        ctx->f2_long = seg_long;
        ctx->f2_rows = seg_short;
        for (auto& s : ctx->slabs) {
          s.f2_n_long = (int)std::max(0LL, (s.rows - rows_short_of(s)) / seg_long);
          if (forced && ctx->opt_f2_nlong >= 0)   // sweeps: the number of long segments given outright
            s.f2_n_long = (int)std::min<long long>(ctx->opt_f2_nlong, (s.rows - 1) / seg_long);
          const int rest = s.rows - s.f2_n_long * seg_long;
          s.f2_segs_y = s.f2_n_long + (rest + seg_short - 1) / seg_short;
        }
      }
    }
  }
  for (auto& s : ctx->slabs)
    s.pstride = std::max(s.blocks, ctx->fuse2 ? (long long)s.f2_strips * s.f2_segs_y : 0LL);
}

int alloc_slab(lbm_ctx* ctx, Slab& s) {
  if (set_device(s)) return 1;
  const int pitch = ctx->pitch;
  s.layout.pitch = pitch;
  s.layout.rows = s.rows;
  s.layout.plane_stride = (long long)(s.rows + 2 * GHOST) * pitch;
  s.layout.buf_floats = 9 * s.layout.plane_stride;
  s.layout.flags_offset = 2 * s.layout.buf_floats * (long long)sizeof(float);
  s.layout.mask_offset = s.layout.flags_offset + 256;
  const size_t mask_bytes = sizeof(uint32_t) * (size_t)ctx->mask_pitch * (size_t)(s.rows + 2);
  // + tail pad: the two-step kernel's bulk copies read whole TX+8-float rows and may run past a short row
  const size_t arena_bytes = (size_t)s.layout.mask_offset + mask_bytes + 8192;
  // stream-ordered clears: a plain cudaMemset runs on the legacy stream, which the slab's
  // non-blocking stream does not wait for, and could still be clearing when lbm_upload copies
  CK(cudaMalloc(&s.arena, arena_bytes));
  CK(cudaMemsetAsync(s.arena, 0, arena_bytes, s.stream));
  s.mask = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(s.arena) + s.layout.mask_offset) + ctx->mask_pitch;
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaEventCreate(&s.ev_start));
  CK(cudaEventCreate(&s.ev_stop));
  return 0;
}

int ensure_av_capacity(lbm_ctx* ctx, long long need) {
  for (auto& s : ctx->slabs) {
    if (s.av_capacity >= need) continue;
    if (set_device(s)) return 1;
    long long cap = std::max(need, std::max<long long>(s.av_capacity * 2, 1024));
    double *hi = nullptr, *lo = nullptr;
    auto grow = [&]() -> int {
      CK(cudaMalloc(&hi, sizeof(double) * cap));
      CK(cudaMalloc(&lo, sizeof(double) * cap));
      CK(cudaMemsetAsync(hi, 0, sizeof(double) * cap, s.stream));
      CK(cudaMemsetAsync(lo, 0, sizeof(double) * cap, s.stream));
      if (s.av_capacity > 0) {
        CK(cudaMemcpyAsync(hi, s.av_hi, sizeof(double) * s.av_capacity, cudaMemcpyDeviceToDevice, s.stream));
        CK(cudaMemcpyAsync(lo, s.av_lo, sizeof(double) * s.av_capacity, cudaMemcpyDeviceToDevice, s.stream));
      }
      CK(cudaStreamSynchronize(s.stream));
      return 0;
    };
    if (grow()) {   // keep the old buffers, drop the half-made new ones
      cudaFree(hi);
      cudaFree(lo);
      (void)cudaGetLastError();
      return 1;
    }
    if (s.av_hi) CK(cudaFree(s.av_hi));
    if (s.av_lo) CK(cudaFree(s.av_lo));
    s.av_hi = hi;
    s.av_lo = lo;
    s.av_capacity = cap;
  }
  return 0;
}

int ensure_partials(lbm_ctx* ctx) {
  for (auto& s : ctx->slabs) {
    if (s.partials) continue;
    if (set_device(s)) return 1;
    // the persistent kernel has at most one block per row, the step kernel s.blocks blocks
    s.partial_capacity = (long long)ctx->chunk_steps *
                         std::max<long long>(s.pstride, ctx->persistent ? s.rows : ctx->tile ? ctx->tiles_x * ctx->tiles_y : 0);
    CK(cudaMalloc(&s.partials, sizeof(double2) * (size_t)s.partial_capacity));
    CK(cudaMalloc(&s.scratch, sizeof(double2) * (size_t)ctx->chunk_steps * (size_t)s.splits));
    CK(cudaMalloc(&s.tickets, sizeof(unsigned int) * (size_t)ctx->chunk_steps));
    CK(cudaMemsetAsync(s.tickets, 0, sizeof(unsigned int) * (size_t)ctx->chunk_steps, s.stream));

  }
  return 0;
}

// Neighbours inside one process: slab i's up neighbour is slab i+1; the ring's
// ends are closed here when the context is the whole ring.
int wire_local_neighbours(lbm_ctx* ctx) {
  const int n = (int)ctx->slabs.size();
  for (int i = 0; i < n; i++) {
    Slab& s = ctx->slabs[i];
    if (i + 1 < n || ctx->nranks == 1) {
      const Slab& u = ctx->slabs[(i + 1) % n];
      s.up.arena = u.arena;
      s.up.layout = u.layout;
    }
    if (i > 0 || ctx->nranks == 1) {
      const Slab& d = ctx->slabs[(i + n - 1) % n];
      s.down.arena = d.arena;
      s.down.layout = d.layout;
    }
  }
  // peer access between distinct devices
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      const int a = ctx->slabs[i].device, b = ctx->slabs[j].device;
      if (a == b) continue;
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, a, b));
      if (!can) return fail("device %d cannot access device %d (no P2P)", a, b);
      CK(cudaSetDevice(a));
      cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail("cudaDeviceEnablePeerAccess(%d -> %d): %s", a, b, cudaGetErrorString(e));
      (void)cudaGetLastError();
    }
  ctx->connected = (ctx->nranks == 1);
  return 0;
}

RingPlan plan_of(const lbm_ctx* ctx) {
  RingPlan pl{};
  pl.nx = ctx->p.nx;
  pl.ny = ctx->p.ny;
  pl.nranks = ctx->nranks;
  pl.fuse2 = ctx->fuse2;
  pl.f2_kernel = ctx->fuse2 ? ctx->f2_kernel : 0;
  pl.f2_rows = ctx->fuse2 ? ctx->f2_rows : 0;
  pl.f2_long = ctx->fuse2 ? ctx->f2_long : 0;
  pl.V = ctx->V;
  return pl;
}

// A ring wait that timed out left its epoch in the slab's error word (wait_epoch, lbm_kernels.cuh).
int check_ring_health(lbm_ctx* ctx) {
  if (ctx->failed) return fail("the context already failed (a wait timed out, see the first error); destroy it");
  if (!ctx->ring) return 0;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    unsigned long long word = 0;
    CK(cudaMemcpyAsync(&word, s.flags() + 4, sizeof word, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    if (word != 0ULL) {
      ctx->failed = true;
      return fail("ring neighbour of rank %d (rows %d..%d) did not reach epoch %llu within %ld ms: it died, was "
                  "launched with a different step count, or never called lbm_run; results are invalid",
                  ctx->rank, s.y0, s.y0 + s.rows - 1, word, ctx->wait_timeout_ms);
    }
  }
  return 0;
}

int create_common(lbm_ctx** out, const lbm_params* p, int nslabs, const int* devices, int rank, int nranks,
                  int y0, int rows) {
  if (!out) return fail("out is NULL");
  *out = nullptr;
  if (validate(p)) return 1;
  if (nslabs < 1 || nslabs > rows) return fail("cannot split %d rows into %d slabs", rows, nslabs);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail("no CUDA device is visible");
  lbm_ctx* ctx = new lbm_ctx();
  ctx->p = *p;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->y0 = y0;
  ctx->rows = rows;
  ctx->pitch = (p->nx + 31) / 32 * 32;
  ctx->mask_pitch = ctx->pitch / 32;
  ctx->ring = (nslabs > 1) || (nranks > 1);
  ctx->w1 = p->density * p->accel / 9.0f;   // kernels.cl:14
  ctx->w2 = p->density * p->accel / 36.0f;  // kernels.cl:15
  ctx->slabs.resize(nslabs);
  for (int i = 0; i < nslabs; i++) {
    Slab& s = ctx->slabs[i];
    s.device = devices[i];
    if (s.device < 0 || s.device >= ndev) {
      fail("device ordinal %d out of range (have %d)", s.device, ndev);
      lbm_destroy(ctx);
      return 1;
    }
    int sy0, srows;
    lbm_partition_rows(rows, nslabs, i, &sy0, &srows);
    s.y0 = y0 + sy0;
    s.rows = srows;
    int err = 0;
    s.stream = stream_for(ctx, s.device, &err);
    if (err || alloc_slab(ctx, s)) { lbm_destroy(ctx); return 1; }
  }
  if (wire_local_neighbours(ctx)) { lbm_destroy(ctx); return 1; }
  resolve_options(ctx);
  *out = ctx;
  return 0;
}

struct Variant {
  int V, hint, tpb, tps, packed;
};

template <int V, int HINT, int TPB, int TPS, bool PACKED>
void launch_step_t(const lbm::StepArgs& a, long long blocks, cudaStream_t st) {
  lbm::step_kernel<V, HINT, TPB, TPS, PACKED><<<(unsigned)blocks, TPB, 0, st>>>(a);
}

template <int V, int HINT, int TPB, int TPS>
void launch_step_p(const Variant& v, const lbm::StepArgs& a, long long blocks, cudaStream_t st) {
  if constexpr (V > 1) {
    if (v.packed) return launch_step_t<V, HINT, TPB, TPS, true>(a, blocks, st);
  }
  launch_step_t<V, HINT, TPB, TPS, false>(a, blocks, st);
}

template <int V, int HINT, int TPB>
void launch_step_s(const Variant& v, const lbm::StepArgs& a, long long blocks, cudaStream_t st) {
  if (v.tps == 768 && TPB != 512) launch_step_p<V, HINT, TPB, 768>(v, a, blocks, st);
  else launch_step_p<V, HINT, TPB, 1024>(v, a, blocks, st);
}

template <int V, int HINT>
void launch_step_v(const Variant& v, const lbm::StepArgs& a, long long blocks, cudaStream_t st) {
  if (v.tpb == 128) launch_step_s<V, HINT, 128>(v, a, blocks, st);
  else if (v.tpb == 512) launch_step_s<V, HINT, 512>(v, a, blocks, st);
  else launch_step_s<V, HINT, 256>(v, a, blocks, st);
}

template <int HINT>
void launch_step_h(const Variant& v, const lbm::StepArgs& a, long long blocks, cudaStream_t st) {
  if (v.V == 4) launch_step_v<4, HINT>(v, a, blocks, st);
  else if (v.V == 2) launch_step_v<2, HINT>(v, a, blocks, st);
  else launch_step_v<1, HINT>(v, a, blocks, st);
}

void launch_step(const Variant& v, const lbm::StepArgs& a, long long blocks, cudaStream_t st) {
  if (v.hint) launch_step_h<5>(v, a, blocks, st);   // coherent L2 loads (rings; option streaming = 1)
  else launch_step_h<0>(v, a, blocks, st);
}

template <int V, int TPB, bool PACKED>
int persistent_grid_t(int device, int rows, long long* grid, int* rows_per_block) {
  // block b owns rows [b*rpb, (b+1)*rpb): as many blocks as can be co-resident (cooperative launch)
  int per_sm = 0, sms = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbm::persistent_kernel<V, TPB, PACKED>, TPB, 0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const long long resident = std::max(1LL, (long long)per_sm * sms);
  const int rpb = (int)std::max(1LL, (rows + resident - 1) / resident);
  *rows_per_block = rpb;
  *grid = (rows + rpb - 1) / rpb;
  return 0;
}

template <int V, int TPB, bool PACKED>
int launch_persistent_t(const lbm::PersistArgs& pa, long long grid, cudaStream_t st) {
  void* args[] = {const_cast<lbm::PersistArgs*>(&pa)};
  CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lbm::persistent_kernel<V, TPB, PACKED>), dim3((unsigned)grid),
                                 dim3(TPB), args, 0, st));
  return 0;
}

#define LBM_DISPATCH_V_TPB(V_, TPB_, P_, CALL)                              \
  do {                                                                      \
    if ((V_) == 4) {                                                        \
      if (P_) {                                                             \
        if ((TPB_) == 128) return CALL(4, 128, true);                       \
        if ((TPB_) == 512) return CALL(4, 512, true);                       \
        return CALL(4, 256, true);                                          \
      }                                                                     \
      if ((TPB_) == 128) return CALL(4, 128, false);                        \
      if ((TPB_) == 512) return CALL(4, 512, false);                        \
      return CALL(4, 256, false);                                           \
    } else if ((V_) == 2) {                                                 \
      if (P_) {                                                             \
        if ((TPB_) == 128) return CALL(2, 128, true);                       \
        if ((TPB_) == 512) return CALL(2, 512, true);                       \
        return CALL(2, 256, true);                                          \
      }                                                                     \
      if ((TPB_) == 128) return CALL(2, 128, false);                        \
      if ((TPB_) == 512) return CALL(2, 512, false);                        \
      return CALL(2, 256, false);                                           \
    } else {                                                                \
      if ((TPB_) == 128) return CALL(1, 128, false);                        \
      if ((TPB_) == 512) return CALL(1, 512, false);                        \
      return CALL(1, 256, false);                                           \
    }                                                                       \
  } while (0)

int persistent_grid(const Variant& v, int device, int rows, long long* grid, int* rows_per_block) {
#define CALL_(vv, t, p) persistent_grid_t<vv, t, p>(device, rows, grid, rows_per_block)
  LBM_DISPATCH_V_TPB(v.V, v.tpb, v.packed, CALL_);
#undef CALL_
}

int launch_persistent(const Variant& v, const lbm::PersistArgs& pa, long long grid, cudaStream_t st) {
#define CALL_(vv, t, p) launch_persistent_t<vv, t, p>(pa, grid, st)
  LBM_DISPATCH_V_TPB(v.V, v.tpb, v.packed, CALL_);
#undef CALL_
}

template <int W, bool PACKED, bool FULLW, int MODE>
int launch_fuse2p_t(const lbm::Fuse2Args& fa, long long grid, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    CK(cudaFuncSetAttribute(lbm::fuse2p_kernel<W, PACKED, FULLW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            lbm::fuse2p_smem_bytes<W>()));
    CK(cudaFuncSetAttribute(lbm::fuse2p_kernel<W, PACKED, FULLW, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    configured[dev] = true;
  }
  lbm::fuse2p_kernel<W, PACKED, FULLW, MODE><<<(unsigned)grid, 32 * (W + 1), lbm::fuse2p_smem_bytes<W>(), st>>>(fa);
  return 0;
}

// the re-pipelined two-step kernel (lbm_fuse2p.cuh): 512-column strips.  mode bit 0: one reciprocal /
// square-root range check per thread instead of per pair (packed only); bit 1: dry run (bandwidth
// experiments only)
int launch_fuse2p(int packed, bool fullw, int mode, const lbm::Fuse2Args& fa, long long grid, cudaStream_t st) {
#define F2P_(P, F)                                                                  \
  do {                                                                              \
    if (mode & 2) return launch_fuse2p_t<4, P, F, 2>(fa, grid, st);                 \
    if (P && (mode & 1)) return launch_fuse2p_t<4, P, F, 1>(fa, grid, st);          \
    return launch_fuse2p_t<4, P, F, 0>(fa, grid, st);                               \
  } while (0)
  if (packed) { if (fullw) F2P_(true, true); else F2P_(true, false); }
  if (fullw) F2P_(false, true); else F2P_(false, false);
#undef F2P_
}

template <int W, bool PACKED, bool FULLW, int MODE>
int launch_fuse2q_t(const lbm::Fuse2Args& fa, long long grid, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    CK(cudaFuncSetAttribute(lbm::fuse2q_kernel<W, PACKED, FULLW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            lbm::fuse2q_smem_bytes<W>()));
    CK(cudaFuncSetAttribute(lbm::fuse2q_kernel<W, PACKED, FULLW, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    configured[dev] = true;
  }
  lbm::fuse2q_kernel<W, PACKED, FULLW, MODE><<<(unsigned)grid, 32 * (W + 1), lbm::fuse2q_smem_bytes<W>(), st>>>(fa);
  return 0;
}

// the two-deep-stage variant of the re-pipelined two-step kernel (lbm_fuse2q.cuh); packed arithmetic only
int launch_fuse2q(bool fullw, int mode, const lbm::Fuse2Args& fa, long long grid, cudaStream_t st) {
  if (fullw) return (mode & 2) ? launch_fuse2q_t<4, true, true, 2>(fa, grid, st) : launch_fuse2q_t<4, true, true, 1>(fa, grid, st);
  return (mode & 2) ? launch_fuse2q_t<4, true, false, 2>(fa, grid, st) : launch_fuse2q_t<4, true, false, 1>(fa, grid, st);
}

// local row of global row ny-2 in this slab, or -1
int accel_row_of(const lbm_ctx* ctx, const Slab& s) {
  const int g = ctx->p.ny - 2;
  return (g >= s.y0 && g < s.y0 + s.rows) ? g - s.y0 : -1;
}

int run_impl(lbm_ctx* ctx, int nsteps, bool timed, float* ms) {
  if (!ctx) return fail("ctx is NULL");
  if (nsteps < 0) return fail("nsteps must be >= 0");
  if (!ctx->uploaded) return fail("lbm_run before lbm_upload");
  if (ctx->failed) return fail("lbm_run on a failed context (a wait timed out earlier); destroy it");
  if (ctx->nranks > 1 && !ctx->connected) return fail("lbm_run before lbm_connect on a %d-rank ring", ctx->nranks);
  if (ms) *ms = 0.f;
  if (nsteps == 0) return 0;
  if (ensure_av_capacity(ctx, ctx->steps_since_upload + nsteps)) return 1;
  if (ensure_partials(ctx)) return 1;

  if (timed)
    for (auto& s : ctx->slabs) {
      if (set_device(s)) return 1;
      CK(cudaEventRecord(s.ev_start, s.stream));
    }

  // epoch 1 of the run: step 0's accelerate_flow on the resident state
  ctx->epoch++;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    lbm::AccelArgs a{};
    a.cur = s.row0(ctx->cur);
    a.plane_stride = s.layout.plane_stride;
    a.pitch = ctx->pitch;
    a.nx = ctx->p.nx;
    a.rows = s.rows;
    a.mask = s.mask;
    a.mask_pitch = ctx->mask_pitch;
    a.w1 = ctx->w1;
    a.w2 = ctx->w2;
    a.accel_row = accel_row_of(ctx, s);
    a.up_ghost = nb_ghost_below(s.up, ctx->cur);
    a.up_plane_stride = s.up.layout.plane_stride;
    a.down_ghost = nb_ghost_above(s.down, ctx->cur);
    a.down_plane_stride = s.down.layout.plane_stride;
    if (ctx->ring) {
      a.peer_up_flag = nb_flags(s.up) + 1;     // the up neighbour's flag_from_down
      a.peer_down_flag = nb_flags(s.down) + 0; // the down neighbour's flag_from_up
    }
    a.epoch = ctx->epoch;
    if (a.accel_row >= 0 || ctx->ring) {
      lbm::accelerate_kernel<<<1, 1024, 0, s.stream>>>(a);
      ctx->launches++;
    }
  }

  if (ctx->tile) {
    // one cooperative launch per chunk of steps; K steps per hand-off between neighbouring tiles (lbm_tile.cuh)
    Slab& s = ctx->slabs[0];
    if (set_device(s)) return 1;
    const int ntiles = ctx->tiles_x * ctx->tiles_y;
    static bool configured[64] = {};
    if (s.device < 64 && !configured[s.device]) {
      CK(cudaFuncSetAttribute(lbm::tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CK(cudaFuncSetAttribute(lbm::tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured[s.device] = true;
    }
    int per_sm = 0, sms = 0;
    if (ctx->tile_cpt == 2)
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbm::tile_kernel<2>, ctx->tile_threads, ctx->tile_smem));
    else
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbm::tile_kernel<1>, ctx->tile_threads, ctx->tile_smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device));
    if ((long long)per_sm * sms < ntiles)
      return fail("internal: %d tiles cannot be co-resident (%d blocks per SM on %d SMs)", ntiles, per_sm, sms);
    if (s.progress_capacity < ntiles) {
      if (s.progress) CK(cudaFree(s.progress));
      s.progress = nullptr;
      CK(cudaMalloc(&s.progress, sizeof(unsigned int) * 32 * (size_t)ntiles));
      s.progress_capacity = ntiles;
    }
    const long long max_steps = std::max(1LL, std::min<long long>(ctx->chunk_steps, s.partial_capacity / ntiles));
    long long first = ctx->steps_since_upload;
    for (int done = 0; done < nsteps;) {
      const int n = (int)std::min<long long>(max_steps, nsteps - done);
      lbm::TileArgs ta{};
      ta.buf[0] = s.row0(0);
      ta.buf[1] = s.row0(1);
      ta.plane_stride = s.layout.plane_stride;
      ta.pitch = ctx->pitch;
      ta.nx = ctx->p.nx;
      ta.ny = ctx->p.ny;
      ta.mask = s.mask;
      ta.mask_pitch = ctx->mask_pitch;
      ta.omega = ctx->p.omega;
      ta.w1 = ctx->w1;
      ta.w2 = ctx->w2;
      ta.tiles_x = ctx->tiles_x;
      ta.tiles_y = ctx->tiles_y;
      ta.K = ctx->tile_K;
      ta.lw = ctx->tile_lw;
      ta.lh = ctx->tile_lh;
      ta.nsteps = n;
      ta.first_buf = ctx->cur;
      ta.accel_row = ctx->p.ny - 2;
      ta.skip_last_accel = (done + n == nsteps) ? 1 : 0;
      ta.progress = s.progress;
      ta.partials = s.partials;
      if (ctx->opt_tile_debug && !s.tile_timing) {
        const size_t bytes = sizeof(long long) * lbm::TILE_TIMING_ROUNDS * lbm::TILE_TIMING_SLOTS;
        CK(cudaMalloc(&s.tile_timing, bytes));
        CK(cudaMemsetAsync(s.tile_timing, 0, bytes, s.stream));
      }
      ta.timing = ctx->opt_tile_debug ? s.tile_timing : nullptr;
      CK(cudaMemsetAsync(s.progress, 0, sizeof(unsigned int) * 32 * (size_t)ntiles, s.stream));
      void* kargs[] = {&ta};
      CK(cudaLaunchCooperativeKernel(ctx->tile_cpt == 2 ? reinterpret_cast<void*>(lbm::tile_kernel<2>)
                                                        : reinterpret_cast<void*>(lbm::tile_kernel<1>),
                                     dim3((unsigned)ntiles),
                                     dim3((unsigned)ctx->tile_threads), kargs, (size_t)ctx->tile_smem, s.stream));
      lbm::av_finalize_kernel<<<dim3(1, n), 256, 0, s.stream>>>(s.partials, ntiles, s.scratch, s.tickets, s.av_hi, s.av_lo,
                                                                 first);
      ctx->launches += 2;   // (+ one memset node)
      const int rounds = (n + ctx->tile_K - 1) / ctx->tile_K;
      ctx->cur ^= (rounds & 1);
      first += n;
      done += n;
    }
  }

  if (ctx->persistent) {
    // one cooperative launch per chunk of steps; grid barrier between steps (lbm_kernels.cuh)
    Slab& s = ctx->slabs[0];
    if (set_device(s)) return 1;
    const Variant var{ctx->V, ctx->streaming, ctx->tpb, ctx->tps, ctx->packed};
    if (persistent_grid(var, s.device, s.rows, &s.pblocks, &s.rows_per_block)) return 1;
    if (s.progress_capacity < s.pblocks) {
      if (s.progress) CK(cudaFree(s.progress));
      CK(cudaMalloc(&s.progress, sizeof(unsigned int) * 32 * (size_t)s.pblocks));
      s.progress_capacity = s.pblocks;
    }
    if (s.partial_capacity < s.pblocks) return fail("internal: partial buffer smaller than the persistent grid");
    const long long max_steps = std::max(1LL, std::min<long long>(ctx->chunk_steps, s.partial_capacity / s.pblocks));
    auto fill = [&](lbm::StepArgs& a, int src_buf) {
      a.src = s.row0(src_buf);
      a.dst = s.row0(src_buf ^ 1);
      a.plane_stride = s.layout.plane_stride;
      a.pitch = ctx->pitch;
      a.nx = ctx->p.nx;
      a.rows = s.rows;
      a.segs = ctx->segs;
      a.mask = s.mask;
      a.mask_pitch = ctx->mask_pitch;
      a.omega = ctx->p.omega;
      a.w1 = ctx->w1;
      a.w2 = ctx->w2;
      a.accel_row = -1;
      a.up_ghost = nb_ghost_below(s.up, src_buf ^ 1);
      a.up_plane_stride = s.up.layout.plane_stride;
      a.down_ghost = nb_ghost_above(s.down, src_buf ^ 1);
      a.down_plane_stride = s.down.layout.plane_stride;
    };
    long long first = ctx->steps_since_upload;
    for (int done = 0; done < nsteps;) {
      const int n = (int)std::min<long long>(max_steps, nsteps - done);
      lbm::PersistArgs pa{};
      fill(pa.even, ctx->cur);
      fill(pa.odd, ctx->cur ^ 1);
      pa.nsteps = n;
      pa.accel_row = accel_row_of(ctx, s);
      pa.skip_last_accel = (done + n == nsteps) ? 1 : 0;
      pa.rows_per_block = s.rows_per_block;
      pa.progress = s.progress;
      pa.partials = s.partials;
      pa.global_barrier = ctx->opt_sync;
      CK(cudaMemsetAsync(s.progress, 0, sizeof(unsigned int) * 32 * (size_t)s.pblocks, s.stream));
      if (launch_persistent(var, pa, s.pblocks, s.stream)) return 1;
      lbm::av_finalize_kernel<<<dim3(1, n), 256, 0, s.stream>>>(s.partials, s.pblocks, s.scratch, s.tickets, s.av_hi,
                                                                 s.av_lo, first);
      ctx->launches += 2;   // (+ one memset node)
      ctx->cur ^= (n & 1);
      first += n;
      done += n;
    }
  }

  int in_chunk = 0;
  long long chunk_first = ctx->steps_since_upload;
  auto flush_chunk = [&]() -> int {   // per-step Σ|u| of the chunk's steps -> av_hi/av_lo
    if (in_chunk == 0) return 0;
    for (auto& s : ctx->slabs) {
      if (set_device(s)) return 1;
      lbm::av_finalize_kernel<<<dim3(s.splits, in_chunk), 256, 0, s.stream>>>(s.partials, s.pstride, s.scratch, s.tickets,
                                                                               s.av_hi, s.av_lo, chunk_first);
      ctx->launches++;
    }
    chunk_first += in_chunk;
    in_chunk = 0;
    return 0;
  };
  for (int t = 0; !ctx->persistent && !ctx->tile && t < nsteps;) {
    const bool pair = ctx->fuse2 && (nsteps - t >= 2);   // two steps in one launch; an odd tail runs one step
    const int adv = pair ? 2 : 1;
    if (in_chunk + adv > ctx->chunk_steps && flush_chunk()) return 1;
    ctx->epoch++;
    const bool last = (t + adv == nsteps);
    for (auto& s : ctx->slabs) {
      if (set_device(s)) return 1;
      lbm::StepArgs a{};
      a.src = s.row0(ctx->cur);
      a.dst = s.row0(ctx->cur ^ 1);
      a.plane_stride = s.layout.plane_stride;
      a.pitch = ctx->pitch;
      a.nx = ctx->p.nx;
      a.rows = s.rows;
      a.segs = ctx->segs;
      a.mask = s.mask;
      a.mask_pitch = ctx->mask_pitch;
      a.omega = ctx->p.omega;
      a.w1 = ctx->w1;
      a.w2 = ctx->w2;
      a.accel_row = last ? -1 : accel_row_of(ctx, s);
      a.up_ghost = nb_ghost_below(s.up, ctx->cur ^ 1);
      a.up_plane_stride = s.up.layout.plane_stride;
      a.down_ghost = nb_ghost_above(s.down, ctx->cur ^ 1);
      a.down_plane_stride = s.down.layout.plane_stride;
      a.partials = s.partials + (long long)in_chunk * s.pstride;
      if (ctx->ring) {
        unsigned long long* f = s.flags();
        a.flag_from_up = f + 0;
        a.flag_from_down = f + 1;
        a.edge_count = f + 2;
        a.peer_up_flag = nb_flags(s.up) + 1;
        a.peer_down_flag = nb_flags(s.down) + 0;
        a.error_word = f + 4;
        a.wait_timeout_ns = (unsigned long long)std::max(1L, ctx->wait_timeout_ms) * 1000000ULL;
        // completions this launch adds to each edge counter: the warps of the two edge rows (one-step
        // kernel), or the blocks of the segments holding rows 0,1 / rows-2,rows-1 (two-step kernel)
        if (pair) {
          int top_segs = 0, bottom_segs = 0;
          for (int sy = 0; sy < s.f2_segs_y; sy++) {
            int ys, ye;
            lbm::f2_segment_rows(sy, ctx->f2_rows, ctx->f2_long, s.f2_n_long, s.rows, ys, ye);
            if (ys < 2) bottom_segs++;
            if (ye >= s.rows - 1) top_segs++;
          }
          s.edge_expected += (unsigned long long)s.f2_strips * bottom_segs;
          s.edge_expected_top += (unsigned long long)s.f2_strips * top_segs;
        } else {
          const unsigned long long n = (unsigned long long)ctx->segs * (unsigned long long)std::min(GHOST, s.rows);
          s.edge_expected += n;
          s.edge_expected_top += n;
        }
        a.edge_target = s.edge_expected;
        a.edge_target_top = s.edge_expected_top;
      }
      a.epoch = ctx->epoch;
      if (pair) {
        lbm::Fuse2Args fa{};
        fa.s = a;
        fa.y0 = s.y0;
        fa.ny = ctx->p.ny;
        fa.last = last ? 1 : 0;
        fa.strips = s.f2_strips;
        fa.segs_y = s.f2_segs_y;
        fa.seg_rows = ctx->f2_rows;
        fa.seg_long = ctx->f2_long;
        fa.n_long = s.f2_n_long;
        fa.l2_ahead = ctx->opt_f2_l2ahead;
        fa.partials1 = s.partials + (long long)in_chunk * s.pstride;
        fa.partials2 = s.partials + (long long)(in_chunk + 1) * s.pstride;
        fa.per_step = s.pstride;
        const long long f2grid = (long long)s.f2_strips * s.f2_segs_y;
        const int rc = ctx->f2_kernel == 3
                           ? launch_fuse2q(ctx->p.nx % 512 == 0, ctx->opt_f2_mode, fa, f2grid, s.stream)
                           : launch_fuse2p(ctx->packed, ctx->p.nx % 512 == 0, ctx->opt_f2_mode, fa, f2grid, s.stream);
        if (rc) return 1;
      } else {
        if (s.pstride > s.blocks)   // (tiny grids only) the step kernel writes s.blocks partials: clear the rest
          CK(cudaMemsetAsync(a.partials + s.blocks, 0, sizeof(double2) * (size_t)(s.pstride - s.blocks), s.stream));
        launch_step(Variant{ctx->V, ctx->streaming, ctx->tpb, ctx->tps, ctx->packed}, a, s.blocks, s.stream);
      }
      ctx->launches++;
    }
    ctx->cur ^= 1;
    in_chunk += adv;
    t += adv;
    if ((in_chunk == ctx->chunk_steps || t == nsteps) && flush_chunk()) return 1;
  }
  CK(cudaGetLastError());
  ctx->steps_done += nsteps;
  ctx->steps_since_upload += nsteps;

  if (timed) {
    float worst = 0.f;
    for (auto& s : ctx->slabs) {
      if (set_device(s)) return 1;
      CK(cudaEventRecord(s.ev_stop, s.stream));
    }
    for (auto& s : ctx->slabs) {
      if (set_device(s)) return 1;
      CK(cudaEventSynchronize(s.ev_stop));
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, s.ev_start, s.ev_stop));
      worst = std::max(worst, t);
    }
    if (ms) *ms = worst;
  }
  return 0;
}

int sync_all(lbm_ctx* ctx) {
  for (auto& ds : ctx->streams) {
    CK(cudaSetDevice(ds.first));
    CK(cudaStreamSynchronize(ds.second));
  }
  return 0;
}

int push_halos(lbm_ctx* ctx) {
  // Every slab copies its top GHOST rows (all nine planes) into the up neighbour's ghost rows
  // below, its bottom GHOST rows into the down neighbour's ghost rows above, and the matching
  // obstacle-mask rows (one each way).
  const size_t width = sizeof(float) * (size_t)ctx->p.nx;
  const size_t mwidth = sizeof(uint32_t) * (size_t)ctx->mask_pitch;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    for (int d = 1; d <= GHOST && d <= s.rows; d++) {
      // my row rows-d -> the up neighbour's ghost row -d; my row d-1 -> the down neighbour's ghost row rows+d-1
      const float* top = s.row0(ctx->cur) + (long long)(s.rows - d) * ctx->pitch;
      const float* bot = s.row0(ctx->cur) + (long long)(d - 1) * ctx->pitch;
      float* up = nb_ghost_below(s.up, ctx->cur) - (long long)(d - 1) * ctx->pitch;
      float* down = nb_ghost_above(s.down, ctx->cur) + (long long)(d - 1) * ctx->pitch;
      CK(cudaMemcpy2DAsync(up, sizeof(float) * s.up.layout.plane_stride, top, sizeof(float) * s.layout.plane_stride,
                           width, 9, cudaMemcpyDefault, s.stream));
      CK(cudaMemcpy2DAsync(down, sizeof(float) * s.down.layout.plane_stride, bot, sizeof(float) * s.layout.plane_stride,
                           width, 9, cudaMemcpyDefault, s.stream));
    }
    CK(cudaMemcpyAsync(nb_mask_ghost_below(s.up), s.mask + (size_t)(s.rows - 1) * ctx->mask_pitch, mwidth,
                       cudaMemcpyDefault, s.stream));
    CK(cudaMemcpyAsync(nb_mask_ghost_above(s.down), s.mask, mwidth, cudaMemcpyDefault, s.stream));
  }
  return sync_all(ctx);
}

}  // namespace

// ---------------------------------------------------------------------------
// exported functions
// ---------------------------------------------------------------------------

extern "C" {

const char* lbm_last_error(void) { return g_error.c_str(); }
int lbm_abi_version(void) { return LBM_ABI_VERSION; }

int lbm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

void lbm_partition_rows(int ny, int nparts, int part, int* y0, int* rows) {
  const int base = ny / nparts, rem = ny % nparts;
  const int r = base + (part < rem ? 1 : 0);
  const int start = part * base + (part < rem ? part : rem);
  if (y0) *y0 = start;
  if (rows) *rows = r;
}

int lbm_create_on(lbm_ctx** out, const lbm_params* p, int nslabs, const int* devices) {
  if (!devices) return fail("devices is NULL");
  if (validate(p)) return 1;
  DeviceGuard guard;
  return create_common(out, p, nslabs, devices, 0, 1, 0, p->ny);
}

int lbm_create(lbm_ctx** out, const lbm_params* p, int ngpus) {
  if (validate(p)) return 1;
  DeviceGuard guard;
  std::vector<int> devices;
  if (const char* env = getenv("LBM_DEVICES")) {
    for (const char* c = env; *c;) {
      char* end = nullptr;
      long v = strtol(c, &end, 10);
      if (end == c) break;
      devices.push_back((int)v);
      c = (*end == ',') ? end + 1 : end;
    }
  }
  if (ngpus <= 0) {
    const char* env = getenv("LBM_NGPUS");
    ngpus = env ? atoi(env) : (devices.empty() ? 1 : (int)devices.size());
    if (ngpus <= 0) ngpus = 1;
  }
  if (devices.empty())
    for (int i = 0; i < ngpus; i++) devices.push_back(i);
  if ((int)devices.size() < ngpus) return fail("LBM_DEVICES lists %zu devices, %d needed", devices.size(), ngpus);
  return create_common(out, p, ngpus, devices.data(), 0, 1, 0, p->ny);
}

int lbm_create_slab(lbm_ctx** out, const lbm_params* p, int device, int rank, int nranks, int y0, int rows) {
  if (validate(p)) return 1;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("bad rank %d of %d", rank, nranks);
  if (rows < 1 || y0 < 0 || y0 + rows > p->ny) return fail("bad row range [%d, %d) of %d", y0, y0 + rows, p->ny);
  DeviceGuard guard;
  return create_common(out, p, 1, &device, rank, nranks, y0, rows);
}

size_t lbm_export_size(void) { return sizeof(Blob); }

int lbm_export(lbm_ctx* ctx, void* blob) {
  if (!ctx || !blob) return fail("lbm_export: NULL argument");
  if (ctx->slabs.size() != 1) return fail("lbm_export needs a one-slab context (lbm_create_slab)");
  DeviceGuard guard;
  Slab& s = ctx->slabs[0];
  if (set_device(s)) return 1;
  Blob b{};
  b.magic = BLOB_MAGIC;
  CK(cudaIpcGetMemHandle(&b.handle, s.arena));
  b.layout = s.layout;
  b.device = s.device;
  b.rank = ctx->rank;
  b.plan = plan_of(ctx);
  memcpy(blob, &b, sizeof b);
  return 0;
}

int lbm_connect(lbm_ctx* ctx, const void* blob_down, const void* blob_up) {
  if (!ctx || !blob_down || !blob_up) return fail("lbm_connect: NULL argument");
  if (ctx->slabs.size() != 1) return fail("lbm_connect needs a one-slab context (lbm_create_slab)");
  if (ctx->nranks == 1) return 0;
  DeviceGuard guard;
  Slab& s = ctx->slabs[0];
  if (set_device(s)) return 1;
  Blob d, u;
  memcpy(&d, blob_down, sizeof d);
  memcpy(&u, blob_up, sizeof u);
  if (d.magic != BLOB_MAGIC || u.magic != BLOB_MAGIC) return fail("lbm_connect: not an lbm_export blob");
  if (d.layout.pitch != ctx->pitch || u.layout.pitch != ctx->pitch) return fail("lbm_connect: neighbour pitch differs");
  // every rank must have planned the same kernels: they wait on each other's epoch flags
  const RingPlan mine = plan_of(ctx);
  for (const Blob* b : {&d, &u}) {
    const RingPlan& o = b->plan;
    if (o.nx != mine.nx || o.ny != mine.ny || o.nranks != mine.nranks)
      return fail("lbm_connect: rank %d describes a %dx%d grid on %d ranks, this rank %dx%d on %d", b->rank, o.nx, o.ny,
                  o.nranks, mine.nx, mine.ny, mine.nranks);
    if (o.fuse2 != mine.fuse2 || o.f2_kernel != mine.f2_kernel || o.f2_rows != mine.f2_rows || o.f2_long != mine.f2_long ||
        o.V != mine.V)
      return fail("lbm_connect: rank %d planned other kernels (two-step %d/%d rows %d/%d, %d cells per thread) than rank %d "
                  "(two-step %d/%d rows %d/%d, %d cells per thread): give every rank the lbm_partition_rows split and "
                  "the same options",
                  b->rank, o.fuse2, o.f2_kernel, o.f2_long, o.f2_rows, o.V, ctx->rank, mine.fuse2, mine.f2_kernel,
                  mine.f2_long, mine.f2_rows, mine.V);
  }
  // a second lbm_connect replaces the first: close what it opened
  if (s.up.ipc && s.up.arena) CK(cudaIpcCloseMemHandle(s.up.arena));
  if (s.down.ipc && s.down.arena) CK(cudaIpcCloseMemHandle(s.down.arena));
  s.up = Neighbour{};
  s.down = Neighbour{};
  ctx->connected = false;
  void* pd = nullptr;
  CK(cudaIpcOpenMemHandle(&pd, d.handle, cudaIpcMemLazyEnablePeerAccess));
  s.down.arena = static_cast<float*>(pd);
  s.down.layout = d.layout;
  s.down.ipc = true;
  if (memcmp(&d.handle, &u.handle, sizeof d.handle) == 0) {  // two-rank ring: one neighbour, both sides
    s.up.arena = s.down.arena;
    s.up.layout = d.layout;
    s.up.ipc = false;
  } else {
    void* pu = nullptr;
    CK(cudaIpcOpenMemHandle(&pu, u.handle, cudaIpcMemLazyEnablePeerAccess));
    s.up.arena = static_cast<float*>(pu);
    s.up.layout = u.layout;
    s.up.ipc = true;
  }
  ctx->connected = true;
  return 0;
}

void lbm_destroy(lbm_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard guard;
  for (auto& ds : ctx->streams) {
    cudaSetDevice(ds.first);
    cudaStreamSynchronize(ds.second);
  }
  for (auto& s : ctx->slabs) {
    cudaSetDevice(s.device);
    if (s.up.ipc && s.up.arena) cudaIpcCloseMemHandle(s.up.arena);
    if (s.down.ipc && s.down.arena) cudaIpcCloseMemHandle(s.down.arena);
    if (s.arena) cudaFree(s.arena);
    if (s.stage) cudaFree(s.stage);
    for (float* f : s.fs_stage)
      if (f) cudaFree(f);
    if (s.partials) { cudaFree(s.partials); cudaFree(s.scratch); cudaFree(s.tickets); }
    if (s.progress) cudaFree(s.progress);
    if (s.tile_timing) cudaFree(s.tile_timing);
    if (s.av_hi) cudaFree(s.av_hi);
    if (s.av_lo) cudaFree(s.av_lo);
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_stop) cudaEventDestroy(s.ev_stop);
  }
  for (auto& ds : ctx->streams) {
    cudaSetDevice(ds.first);
    cudaStreamDestroy(ds.second);
  }
  (void)cudaGetLastError();
  delete ctx;
}

}  // extern "C"

namespace {

// Uploads the lattice and the obstacle map: either the reference's int-per-cell map (packed to the
// bit mask on the device) or an already packed mask (one 32-bit word per 32 cells of a row).
int upload_impl(lbm_ctx* ctx, const float* cells_soa, const int* obstacles, const unsigned int* mask_words) {
  DeviceGuard guard;
  if (sync_all(ctx)) return 1;
  resolve_options(ctx);
  for (auto& s : ctx->slabs)
    if (s.partials) {  // geometry options may have changed
      if (set_device(s)) return 1;
      CK(cudaFree(s.partials));
      CK(cudaFree(s.scratch));
      CK(cudaFree(s.tickets));
      s.partials = nullptr;
    }
  const int nx = ctx->p.nx;
  const size_t host_plane = (size_t)nx * ctx->rows;
  ctx->cur = 0;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    const size_t row_off = (size_t)(s.y0 - ctx->y0) * nx;
    for (int k = 0; k < 9; k++) {
      float* dst = s.row0(0) + k * s.layout.plane_stride;
      const float* src = cells_soa + k * host_plane + row_off;
      if (ctx->pitch == nx)   // rows are contiguous on both sides: one flat copy per plane (full PCIe rate)
        CK(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)nx * s.rows, cudaMemcpyHostToDevice, s.stream));
      else
        CK(cudaMemcpy2DAsync(dst, sizeof(float) * ctx->pitch, src, sizeof(float) * nx, sizeof(float) * nx, s.rows,
                             cudaMemcpyHostToDevice, s.stream));
    }
    if (mask_words) {
      // the device mask has the same shape (mask_pitch = ceil(nx/32) words per row): one flat copy
      CK(cudaMemcpyAsync(s.mask, mask_words + (size_t)(s.y0 - ctx->y0) * ctx->mask_pitch,
                         sizeof(uint32_t) * (size_t)ctx->mask_pitch * s.rows, cudaMemcpyHostToDevice, s.stream));
    } else {
      // obstacle ints -> bit mask, staged through a bounded device buffer kept for later uploads; copies and
      // pack kernels are stream-ordered, so the buffer can be reused without host synchronisation
      // (rows sit on gridDim.y: at most 65535 per launch)
      const int stage_rows =
          (int)std::max<long long>(1, std::min<long long>(std::min(s.rows, 65535), (64LL << 20) / std::max(1, nx)));
      if (!s.stage) CK(cudaMalloc(&s.stage, sizeof(int) * (size_t)stage_rows * nx));
      for (int r0 = 0; r0 < s.rows; r0 += stage_rows) {
        const int nr = std::min(stage_rows, s.rows - r0);
        CK(cudaMemcpyAsync(s.stage, obstacles + row_off + (size_t)r0 * nx, sizeof(int) * (size_t)nr * nx,
                           cudaMemcpyHostToDevice, s.stream));
        dim3 grid(ctx->mask_pitch, nr);
        lbm::pack_obstacles_kernel<<<grid, 32, 0, s.stream>>>(s.stage, nx, s.mask + (size_t)r0 * ctx->mask_pitch,
                                                              ctx->mask_pitch);
        ctx->launches++;
      }
    }
    if (s.av_hi) CK(cudaMemsetAsync(s.av_hi, 0, sizeof(double) * s.av_capacity, s.stream));
    if (s.av_lo) CK(cudaMemsetAsync(s.av_lo, 0, sizeof(double) * s.av_capacity, s.stream));
  }
  CK(cudaGetLastError());
  if (sync_all(ctx)) return 1;
  ctx->steps_since_upload = 0;
  ctx->uploaded = true;
  if (ctx->nranks == 1) return push_halos(ctx);
  return 0;
}

}  // namespace

extern "C" {

int lbm_upload(lbm_ctx* ctx, const float* cells_soa, const int* obstacles) {
  if (!ctx || !cells_soa || !obstacles) return fail("lbm_upload: NULL argument");
  return upload_impl(ctx, cells_soa, obstacles, nullptr);
}

int lbm_upload_packed(lbm_ctx* ctx, const float* cells_soa, const unsigned int* mask_words) {
  if (!ctx || !cells_soa || !mask_words) return fail("lbm_upload_packed: NULL argument");
  return upload_impl(ctx, cells_soa, nullptr, mask_words);
}

size_t lbm_mask_words_per_row(int nx) { return (size_t)((nx + 31) / 32); }

void lbm_pack_obstacles(const int* obstacles, int nx, int rows, unsigned int* mask_words) {
  const size_t wpr = lbm_mask_words_per_row(nx);
  for (int r = 0; r < rows; r++) {
    const int* row = obstacles + (size_t)r * nx;
    unsigned int* out = mask_words + (size_t)r * wpr;
    for (size_t w = 0; w < wpr; w++) {
      unsigned int word = 0;
      const int x0 = (int)w * 32, n = std::min(32, nx - x0);
      for (int b = 0; b < n; b++) word |= (row[x0 + b] != 0 ? 1u : 0u) << b;
      out[w] = word;
    }
  }
}

int lbm_halo_push(lbm_ctx* ctx) {
  if (!ctx) return fail("ctx is NULL");
  if (!ctx->uploaded) return fail("lbm_halo_push before lbm_upload");
  if (ctx->nranks > 1 && !ctx->connected) return fail("lbm_halo_push before lbm_connect");
  DeviceGuard guard;
  return push_halos(ctx);
}

int lbm_download_cells(lbm_ctx* ctx, float* cells_soa) {
  if (!ctx || !cells_soa) return fail("lbm_download_cells: NULL argument");
  if (!ctx->uploaded) return fail("lbm_download_cells before lbm_upload");
  DeviceGuard guard;
  if (check_ring_health(ctx)) return 1;
  const int nx = ctx->p.nx;
  const size_t host_plane = (size_t)nx * ctx->rows;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    const size_t row_off = (size_t)(s.y0 - ctx->y0) * nx;
    for (int k = 0; k < 9; k++) {
      float* dst = cells_soa + k * host_plane + row_off;
      const float* src = s.row0(ctx->cur) + k * s.layout.plane_stride;
      if (ctx->pitch == nx)
        CK(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)nx * s.rows, cudaMemcpyDeviceToHost, s.stream));
      else
        CK(cudaMemcpy2DAsync(dst, sizeof(float) * nx, src, sizeof(float) * ctx->pitch, sizeof(float) * nx, s.rows,
                             cudaMemcpyDeviceToHost, s.stream));
    }
  }
  return sync_all(ctx);
}

int lbm_download_final_state(lbm_ctx* ctx, float* u_x, float* u_y, float* u, float* pressure) {
  if (!ctx) return fail("ctx is NULL");
  if (!ctx->uploaded) return fail("lbm_download_final_state before lbm_upload");
  DeviceGuard guard;
  if (check_ring_health(ctx)) return 1;
  float* host[4] = {u_x, u_y, u, pressure};
  const int nx = ctx->p.nx;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    // bounded staging kept by the slab (freed by lbm_destroy): at most ~64 Mi cells per field at a time,
    // and at most 65535 rows per launch (rows sit on gridDim.y); the copies are stream-ordered behind the
    // kernel of the same chunk, so one synchronisation per slab suffices
    const int chunk =
        (int)std::max<long long>(1, std::min<long long>(std::min(s.rows, 65535), (64LL << 20) / std::max(1, nx)));
    const long long need = (long long)chunk * nx;
    for (int f = 0; f < 4; f++) {
      if (!host[f] || (s.fs_stage[f] && s.fs_capacity >= need)) continue;
      if (s.fs_stage[f]) CK(cudaFree(s.fs_stage[f]));
      s.fs_stage[f] = nullptr;
      CK(cudaMalloc(&s.fs_stage[f], sizeof(float) * (size_t)need));
    }
    s.fs_capacity = std::max(s.fs_capacity, need);
    for (int r0 = 0; r0 < s.rows; r0 += chunk) {
      const int nr = std::min(chunk, s.rows - r0);
      dim3 grid((nx + 255) / 256, nr);
      lbm::final_state_kernel<<<grid, 256, 0, s.stream>>>(s.row0(ctx->cur), s.layout.plane_stride, ctx->pitch, nx, r0,
                                                          s.mask, ctx->mask_pitch, ctx->p.density,
                                                          host[0] ? s.fs_stage[0] : nullptr, host[1] ? s.fs_stage[1] : nullptr,
                                                          host[2] ? s.fs_stage[2] : nullptr, host[3] ? s.fs_stage[3] : nullptr);
      ctx->launches++;
      const size_t off = ((size_t)(s.y0 - ctx->y0) + r0) * nx;
      for (int f = 0; f < 4; f++)
        if (host[f])
          CK(cudaMemcpyAsync(host[f] + off, s.fs_stage[f], sizeof(float) * (size_t)nr * nx, cudaMemcpyDeviceToHost, s.stream));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s.stream));
  }
  return 0;
}

int lbm_download_av_sums(lbm_ctx* ctx, double* hi, double* lo, int n) {
  if (!ctx || !hi || !lo) return fail("lbm_download_av_sums: NULL argument");
  if (n < 0 || n > ctx->steps_since_upload) return fail("asked for %d averages, %lld steps run", n, ctx->steps_since_upload);
  if (ctx->slabs.size() != 1) return fail("lbm_download_av_sums needs a one-slab context");
  DeviceGuard guard;
  if (n == 0) return 0;
  if (check_ring_health(ctx)) return 1;
  Slab& s = ctx->slabs[0];
  if (set_device(s)) return 1;
  CK(cudaMemcpyAsync(hi, s.av_hi, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(lo, s.av_lo, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return 0;
}

void lbm_combine_av_sums(const double* hi, const double* lo, int nparts, int n, int stride, float free_cells_inv,
                         float* av) {
  for (int i = 0; i < n; i++) {
    double H = 0.0, L = 0.0;
    for (int part = 0; part < nparts; part++) {  // fixed part order; TwoSum keeps the rounding error
      const double x = hi[(size_t)part * stride + i];
      volatile double s = H + x;
      volatile double bb = s - H;
      volatile double err = (H - (s - bb)) + (x - bb);
      H = s;
      L = (L + lo[(size_t)part * stride + i]) + err;
    }
    av[i] = (float)((H + L) * (double)free_cells_inv);  // kernels.cl:202 scales by FREE_CELLS_INV
  }
}

int lbm_download_av_vels(lbm_ctx* ctx, float* av, int n) {
  if (!ctx || !av) return fail("lbm_download_av_vels: NULL argument");
  if (ctx->nranks != 1) return fail("lbm_download_av_vels on a multi-process ring: use lbm_download_av_sums");
  if (n < 0 || n > ctx->steps_since_upload) return fail("asked for %d averages, %lld steps run", n, ctx->steps_since_upload);
  if (n == 0) return 0;
  DeviceGuard guard;
  if (check_ring_health(ctx)) return 1;
  const int parts = (int)ctx->slabs.size();
  std::vector<double> hi((size_t)parts * n), lo((size_t)parts * n);
  for (int i = 0; i < parts; i++) {
    Slab& s = ctx->slabs[i];
    if (set_device(s)) return 1;
    CK(cudaMemcpyAsync(hi.data() + (size_t)i * n, s.av_hi, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(lo.data() + (size_t)i * n, s.av_lo, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
  }
  if (sync_all(ctx)) return 1;
  lbm_combine_av_sums(hi.data(), lo.data(), parts, n, n, ctx->p.free_cells_inv, av);
  return 0;
}

void* lbm_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    (void)cudaGetLastError();
    fail("cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

void lbm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int lbm_device_numa_node(int device) {
  // the PCI function's NUMA node as the kernel reports it (sysfs); -1: unknown / single node
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  for (char* c = bus; *c; c++)
    if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* fp = fopen(path, "r");
  if (!fp) return -1;
  int node = -1;
  if (fscanf(fp, "%d", &node) != 1) node = -1;
  fclose(fp);
  return node;
}

void* lbm_host_alloc_on(size_t bytes, int device) {
  // Pinned host memory on the NUMA node the GPU hangs off: on a two-socket box eight ranks moving
  // ~20 GB each through whatever node their pages happened to land on share one socket's memory
  // controllers and the inter-socket link (round 1: download 0.18 s on 1 GPU, 0.85 s on 8).  Linux
  // places pages on the node of the thread that first touches them, so: bind the calling thread to
  // the device's node, allocate + touch, restore the affinity.  No libnuma needed.
  cpu_set_t old_set, node_set;
  bool bound = false;
  const int node = lbm_device_numa_node(device);
  if (node >= 0 && sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    if (FILE* fp = fopen(path, "r")) {
      CPU_ZERO(&node_set);
      int a = 0, b = 0, n = 0;
      while (fscanf(fp, "%d", &a) == 1) {   // "0-31,64-95"
        b = a;
        int c = fgetc(fp);
        if (c == '-') {
          if (fscanf(fp, "%d", &b) != 1) b = a;
          c = fgetc(fp);
        }
        for (int i = a; i <= b && i < CPU_SETSIZE; i++)
          if (CPU_ISSET(i, &old_set)) { CPU_SET(i, &node_set); n++; }
        if (c != ',') break;
      }
      fclose(fp);
      if (n > 0 && sched_setaffinity(0, sizeof node_set, &node_set) == 0) bound = true;
    }
  }
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  void* p = nullptr;
  const bool ok = cudaSetDevice(device) == cudaSuccess && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess;
  if (ok) {   // first touch (cudaHostAlloc has usually populated the pages already; this makes it certain)
    volatile char* c = static_cast<volatile char*>(p);
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    for (size_t i = 0; i < bytes; i += page) c[i] = 0;
  }
  if (prev >= 0) (void)cudaSetDevice(prev);
  if (bound) sched_setaffinity(0, sizeof old_set, &old_set);
  if (!ok) {
    (void)cudaGetLastError();
    fail("cudaHostAlloc(%zu) for device %d failed", bytes, device);
    return nullptr;
  }
  return p;
}

int lbm_run(lbm_ctx* ctx, int nsteps) {
  DeviceGuard guard;
  return run_impl(ctx, nsteps, false, nullptr);
}

int lbm_run_timed(lbm_ctx* ctx, int nsteps, float* ms) {
  DeviceGuard guard;
  if (run_impl(ctx, nsteps, true, ms)) return 1;
  return check_ring_health(ctx);
}

int lbm_sync(lbm_ctx* ctx) {
  if (!ctx) return fail("ctx is NULL");
  DeviceGuard guard;
  if (sync_all(ctx)) return 1;
  return check_ring_health(ctx);
}

int lbm_set_option(lbm_ctx* ctx, const char* key, long value) {
  if (!ctx || !key) return fail("lbm_set_option: NULL argument");
  if (!strcmp(key, "wait_timeout_ms")) {   // not a planning option: allowed at any time
    if (value < 1) return fail("wait_timeout_ms must be >= 1");
    ctx->wait_timeout_ms = value;
    return 0;
  }
  // everything else changes which kernels run on which grid; the ranks of a connected ring have
  // already compared their plans (lbm_connect), so it is too late there
  if (ctx->nranks > 1 && ctx->connected)
    return fail("lbm_set_option('%s') after lbm_connect: set options before lbm_export on every rank", key);
  DeviceGuard guard;
  if (!strcmp(key, "cells_per_thread")) ctx->opt_v = (int)value;
  else if (!strcmp(key, "threads_per_block")) ctx->opt_tpb = (int)value;
  else if (!strcmp(key, "streaming")) ctx->opt_streaming = (int)value;
  else if (!strcmp(key, "persistent")) ctx->opt_persistent = (int)value;
  else if (!strcmp(key, "chunk_steps")) ctx->opt_chunk = (int)value;
  else if (!strcmp(key, "tile")) ctx->opt_tile = (int)value;
  else if (!strcmp(key, "tile_debug")) ctx->opt_tile_debug = value ? 1 : 0;   // development: record phase clocks of tile 0
  else if (!strcmp(key, "tile_steps")) ctx->opt_tile_steps = (int)std::max(0L, std::min(64L, value));
  else if (!strcmp(key, "tile_w")) ctx->opt_tile_w = (int)std::max(0L, value);
  else if (!strcmp(key, "tile_h")) ctx->opt_tile_h = (int)std::max(0L, value);
  else if (!strcmp(key, "global_barrier")) ctx->opt_sync = value ? 1 : 0;
  else if (!strcmp(key, "threads_per_sm")) ctx->opt_tps = (int)value;
  else if (!strcmp(key, "packed")) ctx->opt_packed = (int)value;
  else if (!strcmp(key, "fuse2")) ctx->opt_fuse2 = (int)value;
  else if (!strcmp(key, "fuse2_rows")) ctx->opt_f2_rows = (int)value;
  else if (!strcmp(key, "fuse2_tma")) {
    if (value != 2 && value != 3) return fail("fuse2_tma must be 2 (fuse2p_kernel) or 3 (fuse2q_kernel, the default)");
    ctx->opt_f2_tma = (int)value;
  }
  else if (!strcmp(key, "fuse2_nlong")) ctx->opt_f2_nlong = (int)value;  // with fuse2_long > 0: how many long segments per strip
  else if (!strcmp(key, "fuse2_long")) ctx->opt_f2_long = (int)value;   // -1 auto, 0 uniform segments, n: rows of the long ones
  else if (!strcmp(key, "fuse2_mode")) ctx->opt_f2_mode = (int)(value & 3);
  else if (!strcmp(key, "fuse2_l2_ahead")) ctx->opt_f2_l2ahead = (int)std::max(0L, std::min(64L, value));
  else return fail("unknown option '%s'", key);
  if (sync_all(ctx)) return 1;
  resolve_options(ctx);
  for (auto& s : ctx->slabs)  // partial geometry may have changed
    if (s.partials) {
      if (set_device(s)) return 1;
      CK(cudaFree(s.partials));
      CK(cudaFree(s.scratch));
      CK(cudaFree(s.tickets));
      s.partials = nullptr;
    }
  return 0;
}

int lbm_debug_pad_nonzero(lbm_ctx* ctx, long long* count) {
  // Out-of-bounds canary (compute-sanitizer is not available on the GPU pool): the arena is
  // cleared at creation and no kernel may ever write the pad columns [nx, pitch) of any row
  // (ghost rows included) of either buffer; counts the floats there that are no longer 0.
  if (!ctx || !count) return fail("lbm_debug_pad_nonzero: NULL argument");
  *count = 0;
  const int pad = ctx->pitch - ctx->p.nx;
  if (pad == 0) return 0;
  if (sync_all(ctx)) return 1;
  for (auto& s : ctx->slabs) {
    if (set_device(s)) return 1;
    const size_t nrows = (size_t)2 * 9 * (s.rows + 2 * GHOST);
    std::vector<float> host(nrows * pad);
    CK(cudaMemcpy2D(host.data(), sizeof(float) * pad, s.arena + ctx->p.nx, sizeof(float) * ctx->pitch,
                    sizeof(float) * pad, nrows, cudaMemcpyDeviceToHost));
    for (float v : host)
      if (v != 0.0f) (*count)++;
  }
  return 0;
}

int lbm_debug_tile_timing(lbm_ctx* ctx, long long* clocks, int rounds, int slots) {
  // development aid (option "tile_debug" = 1): SM clock stamps of tile 0 / thread 0 of the LAST tile_kernel launch,
  // rounds 8..71: slot 0 round start, 1 neighbours flags seen, 2 halo in shared memory, 3.. after each step barrier,
  // 15 own last step done
  if (!ctx || !clocks) return fail("lbm_debug_tile_timing: NULL argument");
  if (rounds != lbm::TILE_TIMING_ROUNDS || slots != lbm::TILE_TIMING_SLOTS)
    return fail("lbm_debug_tile_timing: expected %d x %d", lbm::TILE_TIMING_ROUNDS, lbm::TILE_TIMING_SLOTS);
  DeviceGuard guard;
  Slab& s = ctx->slabs[0];
  if (!s.tile_timing) return fail("lbm_debug_tile_timing: option tile_debug was not set (or no tile_kernel launch yet)");
  if (set_device(s)) return 1;
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaMemcpy(clocks, s.tile_timing, sizeof(long long) * rounds * slots, cudaMemcpyDeviceToHost));
  return 0;
}

int lbm_debug_fastmath_mismatches(unsigned long long* rcp_bad, unsigned long long* sqrt_bad) {
  // rcp_rn_fast / sqrt_rn_fast + their range tests (lbm_kernels.cuh) against __frcp_rn / __fsqrt_rn
  // over ALL 2^32 float bit patterns, on the current device.
  if (!rcp_bad || !sqrt_bad) return fail("lbm_debug_fastmath_mismatches: NULL argument");
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
  CK(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
  lbm::fastmath_check_kernel<<<148 * 8, 256>>>(d);
  CK(cudaGetLastError());
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
  *rcp_bad = h[0];
  *sqrt_bad = h[1];
  return 0;
}

int lbm_get_info(lbm_ctx* ctx, lbm_info* info) {
  if (!ctx || !info) return fail("lbm_get_info: NULL argument");
  memset(info, 0, sizeof *info);
  info->abi_version = LBM_ABI_VERSION;
  info->nslabs = (int)ctx->slabs.size();
  info->rank = ctx->rank;
  info->nranks = ctx->nranks;
  info->y0 = ctx->y0;
  info->rows = ctx->rows;
  info->pitch = ctx->pitch;
  info->cells_per_thread = ctx->V;
  info->threads_per_block = ctx->tpb;
  info->streaming = ctx->streaming;
  info->steps_per_launch = (ctx->persistent || ctx->tile) ? ctx->chunk_steps : (ctx->fuse2 ? 2 : 1);
  info->steps_done = ctx->steps_done;
  info->kernel_launches = ctx->launches;
  info->partials_per_step = ctx->per_step;
  if (ctx->tile)
    snprintf(info->kernel_name, sizeof info->kernel_name, "tile_kernel<K=%d,tiles=%dx%d,halo=%dx%d%s>", ctx->tile_K,
             ctx->tiles_x, ctx->tiles_y, ctx->tile_lw, ctx->tile_lh, ctx->tile_cpt == 2 ? ",2/thread" : "");
  else if (ctx->persistent)
    snprintf(info->kernel_name, sizeof info->kernel_name, "persistent_kernel<V=%d,tpb=%d,packed=%d>", ctx->V, ctx->tpb,
             ctx->packed);
  else if (ctx->fuse2)
  {
    char rows[16];
    if (ctx->f2_long > 0) snprintf(rows, sizeof rows, "%d/%d", ctx->f2_long, ctx->f2_rows);   // long / short segments
    else snprintf(rows, sizeof rows, "%d", ctx->f2_rows);
    snprintf(info->kernel_name, sizeof info->kernel_name, "%s<W=%d,packed=%d,rows=%s>",
             ctx->f2_kernel == 3 ? "fuse2q_kernel" : "fuse2p_kernel", ctx->f2_warps,
             ctx->packed, rows);
  }
  else
    snprintf(info->kernel_name, sizeof info->kernel_name, "step_kernel<V=%d,hint=%d,tpb=%d,tps=%d,packed=%d>", ctx->V,
             ctx->streaming, ctx->tpb, ctx->tps, ctx->packed);
  return 0;
}

}  // extern "C"

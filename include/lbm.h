/*
 * include/lbm.h — C-ABI of the B200 D2Q9-BGK time-step library (liblbm_b200.so).
 *
 * This is the drop-in boundary: the reference has no plugin ABI, its boundary is
 * the host<->device split inside d2q9-bgk.c, i.e. the OpenCL calls its C host makes
 * around the time loop.  Each entry point below replaces one of those call sites
 * (file:line in ag14774/OpenCL-Lattice-Boltzmann).  Plain C types only.
 *
 * Conventions
 *   - every int-returning function returns 0 on success, non-zero on failure;
 *     lbm_last_error() then gives the message (the reference's convention is
 *     "print to stderr and exit", checkError d2q9-bgk.c:858-866 — the C host
 *     keeps that by calling die() on a non-zero status).
 *   - host arrays are owned by the caller; device memory by the context.
 *   - host cell layout is the reference's SoA (d2q9-bgk.c:73):
 *       cells[sp*nx*ny + ii*nx + jj], sp in 0..8, ii = row (y), jj = column (x)
 *     obstacles[ii*nx + jj] is an int, 0 = fluid, 1 = blocked (d2q9-bgk.c:553-587).
 *   - one host thread drives a context (the reference is single-threaded with one
 *     in-order queue, d2q9-bgk.c:616); lbm_run is asynchronous, lbm_sync and the
 *     downloads block.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_NSPEEDS 9
#define LBM_ABI_VERSION 2

/* t_param, d2q9-bgk.c:81-92: same fields, same order, same types. */
typedef struct lbm_params {
  float density;        /* density per link */
  float accel;          /* density redistribution */
  float omega;          /* relaxation parameter */
  float free_cells_inv; /* 1 / number of non-blocked cells of the WHOLE grid, d2q9-bgk.c:591 */
  int nx;               /* cells in x (global) */
  int ny;               /* cells in y (global) */
  int maxIters;         /* capacity hint for the per-step average array */
  int reynolds_dim;
} lbm_params;

/* Opaque device state; replaces t_ocl (d2q9-bgk.c:97-119). */
typedef struct lbm_ctx lbm_ctx;

/* What a context is doing, for logs and bench.py. */
typedef struct lbm_info {
  int abi_version;
  int nslabs;            /* row slabs held by this context */
  int rank, nranks;      /* position in a multi-process ring (0,1 when single process) */
  int y0, rows;          /* global rows [y0, y0+rows) held by this context */
  int pitch;             /* device row pitch in floats */
  int cells_per_thread;  /* 1, 2 or 4 */
  int threads_per_block;
  int streaming;         /* cache-hint mode of the lattice loads/stores (lbm_kernels.cuh) */
  int steps_per_launch;  /* 1: one kernel per step; >1: persistent multi-step kernel */
  long long steps_done;
  long long kernel_launches; /* launches of this library's kernels since creation */
  long long partials_per_step;
  char kernel_name[64];
} lbm_info;

/* ---- creation / destruction --------------------------------------------- */

/* Replaces the OpenCL half of initialise() (d2q9-bgk.c:600-710: device pick,
 * program build, 5 buffers).  ngpus row-slabs on devices 0..ngpus-1 of THIS
 * process; ngpus == 0 reads the environment (LBM_NGPUS, default 1; LBM_DEVICES
 * = comma list of device ordinals; the reference used OCL_DEVICE,
 * d2q9-bgk.c:920-929). */
int lbm_create(lbm_ctx **out, const lbm_params *p, int ngpus);

/* Same, with explicit slab placement: nslabs slabs, slab i on devices[i]
 * (ordinals may repeat: several slabs on one device exercise the multi-slab
 * path on a single GPU). */
int lbm_create_on(lbm_ctx **out, const lbm_params *p, int nslabs, const int *devices);

/* One-process-per-GPU form: this process holds global rows [y0, y0+rows) on
 * `device` as member `rank` of a periodic ring of `nranks` processes.  Must be
 * followed by lbm_export / lbm_connect (ring neighbours' blobs) before upload
 * when nranks > 1. */
int lbm_create_slab(lbm_ctx **out, const lbm_params *p, int device, int rank, int nranks,
                    int y0, int rows);

/* Even split used by every caller: rows of part `part` of `nparts`. */
void lbm_partition_rows(int ny, int nparts, int part, int *y0, int *rows);

/* Peer plumbing for lbm_create_slab contexts (CUDA IPC handle + layout of the
 * lattice arena).  Blobs are opaque, lbm_export_size() bytes, and are exchanged
 * by the caller (torch.distributed all_gather in bench.py). */
size_t lbm_export_size(void);
int lbm_export(lbm_ctx *ctx, void *blob);
int lbm_connect(lbm_ctx *ctx, const void *blob_down, const void *blob_up);

/* Replaces finalise()'s device half (d2q9-bgk.c:729-741). */
void lbm_destroy(lbm_ctx *ctx);

/* ---- data movement ------------------------------------------------------- */

/* Replaces the two blocking clEnqueueWriteBuffer calls (d2q9-bgk.c:200-209).
 * cells_soa: 9 planes of rows*nx floats, obstacles: rows*nx ints, where rows is
 * the context's own row count (ny for lbm_create contexts).  Packs the obstacle
 * bit mask and scatters slabs.  Blocking.  For nranks > 1 follow with
 * lbm_halo_push on every rank, then a host barrier. */
int lbm_upload(lbm_ctx *ctx, const float *cells_soa, const int *obstacles);

/* Same upload with the obstacle map already packed: bit (x & 31) of word
 * [row * lbm_mask_words_per_row(nx) + (x >> 5)] is 1 for a blocked cell.  Optional fast path for
 * callers that re-upload the same geometry (the int map is 10 % of the reference-shaped upload,
 * d2q9-bgk.c:205-209, and is packed 32:1 on the device anyway); lbm_pack_obstacles builds the
 * words from the reference's int map on the host. */
int lbm_upload_packed(lbm_ctx *ctx, const float *cells_soa, const unsigned int *mask_words);
size_t lbm_mask_words_per_row(int nx);
void lbm_pack_obstacles(const int *obstacles, int nx, int rows, unsigned int *mask_words);

/* Pushes this context's edge rows into the ring neighbours' ghost rows
 * (single-process contexts do this inside lbm_upload).  Blocking. */
int lbm_halo_push(lbm_ctx *ctx);

/* Replaces clEnqueueReadBuffer(ocl.cells) (d2q9-bgk.c:251-254); always reads
 * the CURRENT buffer (the reference reads ocl.cells whatever the parity). */
int lbm_download_cells(lbm_ctx *ctx, float *cells_soa);

/* Output stage: the four per-cell fields write_values() prints
 * (d2q9-bgk.c:789-831: u_x, u_y, |u|, pressure; obstacle cells 0,0,0,density/3), computed on
 * the device from the current state with the host code's exact fp32 arithmetic and copied to
 * rows*nx host floats each.  Any pointer may be NULL to skip that field. */
int lbm_download_final_state(lbm_ctx *ctx, float *u_x, float *u_y, float *u, float *pressure);

/* Replaces clEnqueueReadBuffer(ocl.avgs) (d2q9-bgk.c:257-260): the first n
 * per-step averages since the last lbm_upload.  Single-process contexts only
 * (all slabs local); multi-process callers use lbm_download_av_sums. */
int lbm_download_av_vels(lbm_ctx *ctx, float *av, int n);

/* Per-step sum of |u| over this context's fluid cells as an unevaluated
 * double-double (hi + lo), n steps.  Ranks combine them in rank order with
 * lbm_combine_av_sums — the deterministic cross-GPU reduction. */
int lbm_download_av_sums(lbm_ctx *ctx, double *hi, double *lo, int n);
void lbm_combine_av_sums(const double *hi, const double *lo, int nparts, int n, int stride,
                         float free_cells_inv, float *av);

/* Pinned host memory for the arrays above (optional; plain malloc works too).
 * lbm_host_alloc_on places the pages on the NUMA node `device` is attached to (what makes the
 * uploads / downloads of eight ranks on a two-socket box run at full PCIe rate each);
 * lbm_device_numa_node reports that node (-1: unknown). */
void *lbm_host_alloc(size_t bytes);
void *lbm_host_alloc_on(size_t bytes, int device);
void lbm_host_free(void *p);
int lbm_device_numa_node(int device);

/* ---- the time loop -------------------------------------------------------- */

/* Replaces the loop d2q9-bgk.c:221-238 (accelerate_flow :282-303, timestep
 * :306-336, reduce :339-393): enqueues nsteps time steps; step t =
 * accelerate(row ny-2) -> fused propagate+rebound+collision -> av[t].
 * Asynchronous.  In a ring every rank must call it with the same nsteps. */
int lbm_run(lbm_ctx *ctx, int nsteps);

/* Replaces clFinish (d2q9-bgk.c:239).  On a ring it also reports a neighbour that never
 * answered: the in-kernel waits are bounded (option "wait_timeout_ms", default 20 000), a
 * timed-out wait marks the context failed and lbm_sync / the downloads return non-zero with a
 * message — loud and fatal like checkError (d2q9-bgk.c:858-866), never a silent hang. */
int lbm_sync(lbm_ctx *ctx);

/* lbm_run + lbm_sync bracketed by CUDA events on the launching stream;
 * *ms = device time of the nsteps steps (max over this context's slabs). */
int lbm_run_timed(lbm_ctx *ctx, int nsteps, float *ms);

/* ---- misc ----------------------------------------------------------------- */

/* Tuning knobs, before lbm_upload (on a multi-process ring: before lbm_export — lbm_connect
 * compares the ranks' kernel plans and later changes are refused): "cells_per_thread" (0 = auto, 1, 2, 4),
 * "threads_per_block", "threads_per_sm" (register bound: 512, 768, 1024), "streaming" (0: default
 * caching, 1: .cs hints), "persistent" (-1 auto, 0, 1), "global_barrier", "chunk_steps",
 * "tile" (-1 auto, 0, 1: the multi-step tile kernel for lattices that fit the SMs' shared memory),
 * "tile_steps" (time steps per hand-off), "tile_w", "tile_h" (tile size), "fuse2" (-1 auto, 0, 1: two
 * time steps per launch), "fuse2_tma" (which two-step kernel: 3 fuse2q_kernel, two-deep stage = default; 2 fuse2p_kernel,
 * its one-deep predecessor, also used for scalar arithmetic and fuse2_mode 0), "fuse2_rows", "fuse2_long" (-1 auto, 0 uniform row segments, n: rows of the leading long
 * segments), "fuse2_mode" (bit 0: one reciprocal/sqrt range check per thread; bit 1: dry run without
 * arithmetic for bandwidth experiments — garbage results), "wait_timeout_ms" (ring waits; any time).
 * Unknown key -> non-zero. */
int lbm_set_option(lbm_ctx *ctx, const char *key, long value);
int lbm_get_info(lbm_ctx *ctx, lbm_info *info);
/* Debug canary: number of non-zero floats in the pad columns [nx, pitch) of every row of both
 * lattice buffers (cleared at creation, never written by a correct kernel). */
int lbm_debug_pad_nonzero(lbm_ctx *ctx, long long *count);
/* Debug check of the two-step kernel's branch-light reciprocal / square root (lbm_kernels.cuh
 * rcp_rn_fast, sqrt_rn_fast) against the correctly rounded __frcp_rn / __fsqrt_rn over all 2^32
 * float bit patterns on the current device; both counts must come back 0. */
int lbm_debug_fastmath_mismatches(unsigned long long *rcp_bad, unsigned long long *sqrt_bad);
/* Development aid for the small-deck tile kernel (option "tile_debug" = 1): SM clock stamps of tile 0 over
 * 64 rounds of the last launch, clocks[64][16] (slot meanings in csrc/lbm_tile.cuh). */
int lbm_debug_tile_timing(lbm_ctx *ctx, long long *clocks, int rounds, int slots);
int lbm_device_count(void);
int lbm_abi_version(void);

/* Message of the last failure on this thread ("" if none); checkError/die
 * equivalent (d2q9-bgk.c:858-874). */
const char *lbm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */

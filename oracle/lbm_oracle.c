/*
 * oracle/lbm_oracle.c — CPU restatement of the reference's D2Q9-BGK time step.
 *
 * TEST INFRASTRUCTURE ONLY (see lbm_oracle.h).  Plain C99, IEEE arithmetic,
 * built with -ffp-contract=off so that every fp32 operation below is exactly
 * the correctly rounded add / mul / fma / div / sqrt written, in the order
 * written; the CUDA path is required to reproduce the resulting lattice bit
 * for bit.
 *
 * Contraction: OpenCL C contracts a*b+c into a fused multiply-add by default
 * (FP_CONTRACT is ON; the reference even adds -cl-fast-relaxed-math for
 * 128x128, d2q9-bgk.c:642-645, and calls mad() itself, kernels.cl:36-42), so
 * the collision below states the contraction explicitly: every product of
 * kernels.cl:143-197 that feeds an addition is one fmaf().  Nothing is left to
 * a compiler's discretion on either side (gcc: -ffp-contract=off; nvcc:
 * explicitly rounded intrinsics).
 *
 * Citations are file:line in the reference tree (ag14774/OpenCL-Lattice-Boltzmann).
 *
 * Reference built-ins are restated with their exactly rounded counterparts:
 *   native_recip(x) -> 1.0f / x          (kernels.cl:129)
 *   native_sqrt(x)  -> sqrtf(x)          (kernels.cl:198)
 *   native_divide   -> /                 (kernels.cl:14-15)
 *   mad(mask, w, c) -> mask * w + c       (kernels.cl:36-42; mask is 0 or 1, so the
 *                                          product is exact and fusing changes nothing)
 */
#include "lbm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NSPEEDS 9

int oracle_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* fp32: kernels.cl                                                          */
/* ------------------------------------------------------------------------ */

/* kernels.cl:69 — lookup[k][0]: the slot an obstacle cell writes value k to. */
static const int opposite[NSPEEDS] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

/* kernels.cl:116-198 for one cell.  t0..t8 are the nine gathered values, lmask is
 * obstacle^1.  v<k> is the value the reference stores to plane lookup[k][lmask];
 * speed is the cell's term of tot_u.  (Scalars and a struct of scalars rather than
 * arrays: the row loop of row_update_f32 must stay vectorisable.) */
typedef struct {
  float v0, v1, v2, v3, v4, v5, v6, v7, v8, speed;
} cell_result;

static inline __attribute__((always_inline)) cell_result collide_f32(float t0, float t1, float t2, float t3, float t4,
                                                                     float t5, float t6, float t7, float t8,
                                                                     int lmask, float omega)
{
  const float ic_sq = 3.0f;                       /* kernels.cl:63 */
  const float w0 = 0.4444444444444444444444f;     /* kernels.cl:65 */
  const float w1 = 0.1111111111111111111111f;     /* kernels.cl:66 */
  const float w2 = 0.0277777777777777777778f;     /* kernels.cl:67 */
  cell_result out;

  /* kernels.cl:119-127 */
  float densvec = t0;
  densvec += t1;
  densvec += t2;
  densvec += t3;
  densvec += t4;
  densvec += t5;
  densvec += t6;
  densvec += t7;
  densvec += t8;

  float densinv = 1.0f / densvec;                 /* kernels.cl:129 */

  /* kernels.cl:131-141 (momentum, not velocity) */
  float u_x = t1 + t5;
  u_x += t8;
  u_x -= t3;
  u_x -= t6;
  u_x -= t7;

  float u_y = t2 + t5;
  u_y += t6;
  u_y -= t4;
  u_y -= t7;
  u_y -= t8;

  float u_sq = fmaf(u_x, u_x, u_y * u_y);         /* kernels.cl:143, contracted */

  /* kernels.cl:146-154; uvec[3,4,7,8] are the negatives of uvec[1,2,5,6] */
  const float uvec1 = u_x;
  const float uvec2 = u_y;
  const float uvec5 = u_x + u_y;
  const float uvec6 = -u_x + u_y;

  /* kernels.cl:176-185: `0.5f * densinv*ic_sq` groups as ((0.5f*densinv)*ic_sq) */
  const float half_inv = 0.5f * densinv * ic_sq;
  const float relax = (float)lmask * omega;       /* kernels.cl:187-197: lmask*OMEGA */

  /* d_equ[0] = w0 * (densvec - half_inv*u_sq); v[0] = t[0] + relax*(d_equ[0] - t[0]) */
  {
    const float y = fmaf(-half_inv, u_sq, densvec);
    const float r = fmaf(w0, y, -t0);
    out.v0 = fmaf(relax, r, t0);
  }
  /* the four direction pairs (k, opposite k) = (1,3) (2,4) (5,7) (6,8) */
#define ORACLE_PAIR(kp, km, w)                                                                          \
  do {                                                                                                  \
    const float u = uvec##kp;                                                                           \
    /* ic_sqtimesu = u*ic_sq (kernels.cl:156-164); ic_sqtimesu_sq - u_sq = (u*ic_sq)*u - u_sq (:166-185) */ \
    const float s = fmaf(u * ic_sq, u, -u_sq);                                                          \
    /* d_equ[k] = w * (densvec + ic_sqtimesu[k] + half_inv * s), for +u and for -u */                   \
    const float yp = fmaf(half_inv, s, fmaf(u, ic_sq, densvec));                                        \
    const float ym = fmaf(half_inv, s, fmaf(u, -ic_sq, densvec));                                       \
    /* v[k] = t[k] + relax * (d_equ[k] - t[k]) */                                                       \
    out.v##kp = fmaf(relax, fmaf(w, yp, -t##kp), t##kp);                                                \
    out.v##km = fmaf(relax, fmaf(w, ym, -t##km), t##km);                                                \
  } while (0)
  ORACLE_PAIR(1, 3, w1);
  ORACLE_PAIR(2, 4, w1);
  ORACLE_PAIR(5, 7, w2);
  ORACLE_PAIR(6, 8, w2);
#undef ORACLE_PAIR

  out.speed = (float)lmask * sqrtf(u_sq) * densinv;    /* kernels.cl:198 */
  return out;
}

void oracle_f32_accelerate(const oracle_params *p, float *cells, const int *obstacles)
{
  const size_t nx = p->nx, plane = (size_t)p->nx * p->ny;
  const float w1 = p->density * p->accel / 9.0f;  /* kernels.cl:14 */
  const float w2 = p->density * p->accel / 36.0f; /* kernels.cl:15 */
  const size_t row = (size_t)(p->ny - 2) * nx;    /* kernels.cl:18 */
  float *f1 = cells + 1 * plane + row, *f3 = cells + 3 * plane + row, *f5 = cells + 5 * plane + row;
  float *f6 = cells + 6 * plane + row, *f7 = cells + 7 * plane + row, *f8 = cells + 8 * plane + row;

  for (size_t jj = 0; jj < nx; jj++) {
    float res1 = f3[jj], res2 = f6[jj], res3 = f7[jj];          /* kernels.cl:23-25 */
    int mask = obstacles[row + jj] ^ 1;                          /* kernels.cl:29 */
    mask &= (res1 - w1 > 0.0f) ? 1 : 0;                          /* kernels.cl:30-33 */
    mask &= (res2 - w2 > 0.0f) ? 1 : 0;
    mask &= (res3 - w2 > 0.0f) ? 1 : 0;
    const float m = (float)mask;
    f1[jj] = m * w1 + f1[jj];                                    /* kernels.cl:36-38 */
    f5[jj] = m * w2 + f5[jj];
    f8[jj] = m * w2 + f8[jj];
    f3[jj] = m * -w1 + res1;                                     /* kernels.cl:40-42 */
    f6[jj] = m * -w2 + res2;
    f7[jj] = m * -w2 + res3;
  }
}

/* One cell of kernels.cl:91-198 at column xx with the periodic neighbours x_w, x_e. */
static inline void cell_update_f32(int xx, int x_w, int x_e, float omega, const float *const s[NSPEEDS],
                                   float *const d[NSPEEDS], const int *obst, float *speed)
{
  const int lmask = obst[xx] ^ 1;                                /* kernels.cl:113 */
  /* kernels.cl:104-112 */
  const cell_result r = collide_f32(s[0][xx], s[1][x_w], s[2][xx], s[3][x_e], s[4][xx], s[5][x_w], s[6][x_e],
                                    s[7][x_e], s[8][x_w], lmask, omega);
  const float v[NSPEEDS] = {r.v0, r.v1, r.v2, r.v3, r.v4, r.v5, r.v6, r.v7, r.v8};
  speed[xx] = r.speed;
  if (lmask) {
    for (int k = 0; k < NSPEEDS; k++) d[k][xx] = v[k];            /* lookup[k][1] = k */
  } else {
    for (int k = 0; k < NSPEEDS; k++) d[opposite[k]][xx] = v[k];  /* lookup[k][0] */
  }
}

/* One row of kernels.cl:91-198.  s_*: source rows (already offset to the
 * row the pull reads: own row for 0,1,3; south row for 2,5,6; north row for
 * 4,7,8), d[k]: destination rows, speed[x] gets the cell's tot_u term.
 *
 * ORACLE_SCALAR_ROWS (liboracle.so): the cell-by-cell form above for every column.
 * Otherwise (the -mavx2 / -mavx512f builds, the CPU baseline): the same cells with the same
 * operations in the same order, written so that the compiler can run 8 / 16 columns at a time:
 * the two wrap columns (kernels.cl:100-102) cell by cell, the interior with x_w = xx-1, x_e = xx+1
 * and the store slot chosen by a select (`d[opposite[k]] = v[k]` for every k is `d[k] =
 * v[opposite[k]]` for every k, the table being an involution).  Vector add / mul / fma / div / sqrt
 * are the same correctly rounded IEEE operations lane by lane, so the two forms agree bit for bit
 * (tests/test_oracle_goldens.py compares the builds). */
#ifndef ORACLE_SCALAR_ROWS
static void row_interior_f32(int lo, int hi, float omega, const float *const s[NSPEEDS], float *const d[NSPEEDS],
                             const int *obst, float *speed)
{
  /* columns [lo, hi), 1 <= lo, hi <= nx-1 */
  const float *restrict s0 = s[0], *restrict s1 = s[1], *restrict s2 = s[2], *restrict s3 = s[3],
              *restrict s4 = s[4], *restrict s5 = s[5], *restrict s6 = s[6], *restrict s7 = s[7],
              *restrict s8 = s[8];
  float *restrict d0 = d[0], *restrict d1 = d[1], *restrict d2 = d[2], *restrict d3 = d[3], *restrict d4 = d[4],
        *restrict d5 = d[5], *restrict d6 = d[6], *restrict d7 = d[7], *restrict d8 = d[8];
  const int *restrict ob = obst;
  float *restrict sp = speed;
#pragma omp simd
  for (int xx = lo; xx < hi; xx++) {
    const int lmask = ob[xx] ^ 1;
    const cell_result r = collide_f32(s0[xx], s1[xx - 1], s2[xx], s3[xx + 1], s4[xx], s5[xx - 1], s6[xx + 1],
                                      s7[xx + 1], s8[xx - 1], lmask, omega);
    sp[xx] = r.speed;
    d0[xx] = r.v0;
    d1[xx] = lmask ? r.v1 : r.v3;
    d2[xx] = lmask ? r.v2 : r.v4;
    d3[xx] = lmask ? r.v3 : r.v1;
    d4[xx] = lmask ? r.v4 : r.v2;
    d5[xx] = lmask ? r.v5 : r.v7;
    d6[xx] = lmask ? r.v6 : r.v8;
    d7[xx] = lmask ? r.v7 : r.v5;
    d8[xx] = lmask ? r.v8 : r.v6;
  }
}
#endif

static void row_update_f32(int nx, float omega, const float *const s[NSPEEDS], float *const d[NSPEEDS],
                           const int *obst, float *speed)
{
#ifdef ORACLE_SCALAR_ROWS
  for (int xx = 0; xx < nx; xx++) {
    const int x_e = (xx + 1 >= nx) ? xx + 1 - nx : xx + 1;      /* kernels.cl:100-101 */
    const int x_w = (xx == 0) ? nx - 1 : xx - 1;                /* kernels.cl:102 */
    cell_update_f32(xx, x_w, x_e, omega, s, d, obst, speed);
  }
#else
  cell_update_f32(0, nx - 1, (nx > 1) ? 1 : 0, omega, s, d, obst, speed);
  if (nx > 1) cell_update_f32(nx - 1, nx - 2, 0, omega, s, d, obst, speed);
  row_interior_f32(1, nx - 1, omega, s, d, obst, speed);
#endif
}

static int is_pow2(long n) { return n > 0 && (n & (n - 1)) == 0; }

/* kernels.cl:202-229 for one work-group: 64 items, each the sum of cells jj and
 * jj + nx/2 (kernels.cl:97-99) scaled by FREE_CELLS_INV, then the local tree. */
static float workgroup_partial(const float *speed, int nx, int group_x, float free_cells_inv)
{
  float local_avgs[64];
  for (int item = 0; item < 64; item++) {
    const int jj = group_x * 64 + item;
    float tot_u = 0.0f;
    tot_u += speed[jj];
    tot_u += speed[jj + nx / 2];
    local_avgs[item] = tot_u * free_cells_inv;
  }
  for (int s = 32; s >= 1; s >>= 1)
    for (int i = 0; i < s; i++) local_avgs[i] += local_avgs[i + s];
  return local_avgs[0];
}

/* kernels.cl:234-290 driven by d2q9-bgk.c:350-377, for one row of partial_avgs
 * (one time step); width must be a power of two. */
static float reduce_row(float *partial, long width)
{
  long global_size = width;
  float local[256];
  if (global_size == 1) return partial[0];
  for (;;) {
    global_size /= 2;
    const long lsz = global_size >= 256 ? 256 : global_size;
    const long ngroups = global_size / lsz;
    for (long g = 0; g < ngroups; g++) {
      const long k0 = 2 * g * lsz;
      for (long l = 0; l < lsz; l++) local[l] = partial[k0 + l] + partial[k0 + l + lsz];
      for (long s = lsz / 2; s >= 1; s >>= 1)
        for (long i = 0; i < s; i++) local[i] += local[i + s];
      partial[g] = local[0]; /* groups are consumed in increasing g, so g < k0 + ... is safe */
    }
    if (ngroups == 1) return partial[0];
    global_size = ngroups;
  }
}

float oracle_f32_timestep(const oracle_params *p, const float *src, float *dst,
                          const int *obstacles, int reference_order)
{
  const int nx = p->nx, ny = p->ny;
  const size_t plane = (size_t)nx * ny;
  const int ref_tree = reference_order && nx % 128 == 0 && is_pow2((long)(nx / 128) * ny);
  const int groups_x = ref_tree ? nx / 128 : 1;
  float *partial = (float *)malloc(sizeof(float) * (size_t)groups_x * ny);

#pragma omp parallel
  {
    float *speed = (float *)malloc(sizeof(float) * (size_t)nx);
#pragma omp for schedule(static)
    for (int ii = 0; ii < ny; ii++) {
      const int y_n = (ii + 1 == ny) ? 0 : ii + 1;            /* kernels.cl:91-92 */
      const int y_s = (ii == 0) ? ny - 1 : ii - 1;            /* kernels.cl:93 */
      const float *s[NSPEEDS];
      float *d[NSPEEDS];
      for (int k = 0; k < NSPEEDS; k++) d[k] = dst + k * plane + (size_t)ii * nx;
      s[0] = src + 0 * plane + (size_t)ii * nx;
      s[1] = src + 1 * plane + (size_t)ii * nx;
      s[3] = src + 3 * plane + (size_t)ii * nx;
      s[2] = src + 2 * plane + (size_t)y_s * nx;
      s[5] = src + 5 * plane + (size_t)y_s * nx;
      s[6] = src + 6 * plane + (size_t)y_s * nx;
      s[4] = src + 4 * plane + (size_t)y_n * nx;
      s[7] = src + 7 * plane + (size_t)y_n * nx;
      s[8] = src + 8 * plane + (size_t)y_n * nx;
      row_update_f32(nx, p->omega, s, d, obstacles + (size_t)ii * nx, speed);
      if (ref_tree) {
        /* groupID = group_id_Y * num_groups_X + group_id_X, kernels.cl:208 */
        for (int gx = 0; gx < groups_x; gx++)
          partial[(size_t)ii * groups_x + gx] = workgroup_partial(speed, nx, gx, p->free_cells_inv);
      } else {
        float row = 0.0f;
        for (int xx = 0; xx < nx; xx++) row += speed[xx];
        partial[ii] = row;
      }
    }
    free(speed);
  }

  float av;
  if (ref_tree) {
    av = reduce_row(partial, (long)groups_x * ny);
  } else {
    float tot = 0.0f;
    for (int ii = 0; ii < ny; ii++) tot += partial[ii];
    av = tot * p->free_cells_inv;
  }
  free(partial);
  return av;
}

void oracle_f32_run(const oracle_params *p, float *cells, float *scratch, const int *obstacles,
                    int nsteps, float *av_vels, int reference_order)
{
  /* d2q9-bgk.c:214-238 */
  float *buf[2] = {cells, scratch};
  int curr_read = 0;
  for (int tt = 0; tt < nsteps; tt++) {
    oracle_f32_accelerate(p, buf[curr_read], obstacles);
    const float av = oracle_f32_timestep(p, buf[curr_read], buf[curr_read ^ 1], obstacles, reference_order);
    if (av_vels) av_vels[tt] = av;
    curr_read ^= 1;
  }
  if (curr_read == 1) memcpy(cells, scratch, sizeof(float) * NSPEEDS * (size_t)p->nx * p->ny);
}

float oracle_f32_av_velocity(const oracle_params *p, const float *cells, const int *obstacles)
{
  /* d2q9-bgk.c:396-442 */
  const size_t plane = (size_t)p->nx * p->ny;
  float tot_u = 0.0f;
  for (size_t c = 0; c < plane; c++) {
    if (obstacles[c]) continue;
    float local_density = 0.0f;
    for (int kk = 0; kk < NSPEEDS; kk++) local_density += cells[kk * plane + c];
    const float u_x = (cells[1 * plane + c] + cells[5 * plane + c] + cells[8 * plane + c]
                       - cells[3 * plane + c] - cells[6 * plane + c] - cells[7 * plane + c]) / local_density;
    const float u_y = (cells[2 * plane + c] + cells[5 * plane + c] + cells[6 * plane + c]
                       - cells[4 * plane + c] - cells[7 * plane + c] - cells[8 * plane + c]) / local_density;
    /* `sqrt` on a float argument promotes to double, d2q9-bgk.c:437 */
    tot_u += sqrt((u_x * u_x) + (u_y * u_y));
  }
  return tot_u * p->free_cells_inv;
}

float oracle_f32_reynolds(const oracle_params *p, const float *cells, const int *obstacles)
{
  /* d2q9-bgk.c:747-752 */
  const float viscosity = 1.0f / 6.0f * (2.0f / p->omega - 1.0f);
  return oracle_f32_av_velocity(p, cells, obstacles) * p->reynolds_dim / viscosity;
}

float oracle_f32_total_density(const oracle_params *p, const float *cells)
{
  /* d2q9-bgk.c:754-770 — row, column, speed order */
  const size_t plane = (size_t)p->nx * p->ny;
  float total = 0.0f;
  for (size_t c = 0; c < plane; c++)
    for (int kk = 0; kk < NSPEEDS; kk++) total += cells[kk * plane + c];
  return total;
}

void oracle_f32_final_state(const oracle_params *p, const float *cells, const int *obstacles,
                            float *u_x_out, float *u_y_out, float *u_out, float *pressure_out)
{
  /* d2q9-bgk.c:789-831 */
  const float c_sq = 1.0f / 3.0f;
  const size_t plane = (size_t)p->nx * p->ny;
  for (size_t c = 0; c < plane; c++) {
    if (obstacles[c]) {
      u_x_out[c] = u_y_out[c] = u_out[c] = 0.0f;
      pressure_out[c] = p->density * c_sq;
      continue;
    }
    float local_density = 0.0f;
    for (int kk = 0; kk < NSPEEDS; kk++) local_density += cells[kk * plane + c];
    const float u_x = (cells[1 * plane + c] + cells[5 * plane + c] + cells[8 * plane + c]
                       - cells[3 * plane + c] - cells[6 * plane + c] - cells[7 * plane + c]) / local_density;
    const float u_y = (cells[2 * plane + c] + cells[5 * plane + c] + cells[6 * plane + c]
                       - cells[4 * plane + c] - cells[7 * plane + c] - cells[8 * plane + c]) / local_density;
    u_x_out[c] = u_x;
    u_y_out[c] = u_y;
    u_out[c] = sqrt((u_x * u_x) + (u_y * u_y));
    pressure_out[c] = local_density * c_sq;
  }
}

/* ---- row-slab form (for the world_size-2 CPU tests) ---------------------- */

void oracle_f32_slab_accelerate(const oracle_params *p, float *slab, const int *obstacles,
                                int rows_local, int accel_row_local)
{
  if (accel_row_local < 0 || accel_row_local >= rows_local) return;
  const size_t nx = p->nx, plane = (size_t)p->nx * (rows_local + 2);
  const float w1 = p->density * p->accel / 9.0f;
  const float w2 = p->density * p->accel / 36.0f;
  const size_t row = (size_t)(accel_row_local + 1) * nx;
  const int *ob = obstacles + (size_t)accel_row_local * nx;
  for (size_t jj = 0; jj < nx; jj++) {
    float res1 = slab[3 * plane + row + jj], res2 = slab[6 * plane + row + jj], res3 = slab[7 * plane + row + jj];
    int mask = ob[jj] ^ 1;
    mask &= (res1 - w1 > 0.0f) ? 1 : 0;
    mask &= (res2 - w2 > 0.0f) ? 1 : 0;
    mask &= (res3 - w2 > 0.0f) ? 1 : 0;
    const float m = (float)mask;
    slab[1 * plane + row + jj] = m * w1 + slab[1 * plane + row + jj];
    slab[5 * plane + row + jj] = m * w2 + slab[5 * plane + row + jj];
    slab[8 * plane + row + jj] = m * w2 + slab[8 * plane + row + jj];
    slab[3 * plane + row + jj] = m * -w1 + res1;
    slab[6 * plane + row + jj] = m * -w2 + res2;
    slab[7 * plane + row + jj] = m * -w2 + res3;
  }
}

void oracle_f32_slab_timestep(const oracle_params *p, const float *src, float *dst,
                              const int *obstacles, int rows_local, float *row_sums)
{
  const int nx = p->nx;
  const size_t plane = (size_t)nx * (rows_local + 2);
  float *speed = (float *)malloc(sizeof(float) * (size_t)nx);
  for (int r = 1; r <= rows_local; r++) {
    const float *s[NSPEEDS];
    float *d[NSPEEDS];
    for (int k = 0; k < NSPEEDS; k++) d[k] = dst + k * plane + (size_t)r * nx;
    s[0] = src + 0 * plane + (size_t)r * nx;
    s[1] = src + 1 * plane + (size_t)r * nx;
    s[3] = src + 3 * plane + (size_t)r * nx;
    s[2] = src + 2 * plane + (size_t)(r - 1) * nx;
    s[5] = src + 5 * plane + (size_t)(r - 1) * nx;
    s[6] = src + 6 * plane + (size_t)(r - 1) * nx;
    s[4] = src + 4 * plane + (size_t)(r + 1) * nx;
    s[7] = src + 7 * plane + (size_t)(r + 1) * nx;
    s[8] = src + 8 * plane + (size_t)(r + 1) * nx;
    row_update_f32(nx, p->omega, s, d, obstacles + (size_t)(r - 1) * nx, speed);
    if (row_sums) {
      float row = 0.0f;
      for (int xx = 0; xx < nx; xx++) row += speed[xx];
      row_sums[r - 1] = row;
    }
  }
  free(speed);
}

/* ------------------------------------------------------------------------ */
/* fp64: the original serial equations.                                      */
/*                                                                           */
/* The reference's goldens (check/ *.dat) were produced by the double        */
/* precision serial predecessor of d2q9-bgk.c: the unoptimised serial stage  */
/* matches them with total difference 0 (profiles/0initial/128x128/          */
/* check.txt:2-10).  Its equations survive in the host-side av_velocity and  */
/* write_values (d2q9-bgk.c:411-436, :802-831: velocity = momentum / density */
/* with the west-going sum subtracted as a group) and are, term for term,    */
/* the classic BGK equilibrium that kernels.cl:176-185 encodes in momentum   */
/* form: d_equ = w rho (1 + e.u/c^2 + (e.u)^2/(2 c^4) - u^2/(2 c^2)).         */
/* Step order: accelerate, propagate, rebound, collision, then av_velocity   */
/* on the post-collision state (d2q9-bgk.c:126-131 names the four stages).   */
/* ------------------------------------------------------------------------ */

static void accelerate_f64(const oracle_params64 *p, double *cells, const int *obstacles)
{
  const size_t nx = p->nx, plane = (size_t)p->nx * p->ny;
  const double w1 = p->density * p->accel / 9.0;
  const double w2 = p->density * p->accel / 36.0;
  const size_t row = (size_t)(p->ny - 2) * nx;
  for (size_t jj = 0; jj < nx; jj++) {
    const size_t c = row + jj;
    if (!obstacles[c] && (cells[3 * plane + c] - w1) > 0.0 && (cells[6 * plane + c] - w2) > 0.0
        && (cells[7 * plane + c] - w2) > 0.0) {
      cells[1 * plane + c] += w1;
      cells[5 * plane + c] += w2;
      cells[8 * plane + c] += w2;
      cells[3 * plane + c] -= w1;
      cells[6 * plane + c] -= w2;
      cells[7 * plane + c] -= w2;
    }
  }
}

static double step_f64(const oracle_params64 *p, const double *src, double *dst, const int *obstacles)
{
  const int nx = p->nx, ny = p->ny;
  const size_t plane = (size_t)nx * ny;
  const double c_sq = 1.0 / 3.0;
  const double w0 = 4.0 / 9.0, w1 = 1.0 / 9.0, w2 = 1.0 / 36.0;
  double *row_tot = (double *)malloc(sizeof(double) * (size_t)ny);
  long free_cells = 0;

#pragma omp parallel for schedule(static) reduction(+ : free_cells)
  for (int ii = 0; ii < ny; ii++) {
    const int y_n = (ii + 1) % ny;
    const int y_s = (ii == 0) ? (ny - 1) : (ii - 1);
    double tot = 0.0;
    for (int jj = 0; jj < nx; jj++) {
      const int x_e = (jj + 1) % nx;
      const int x_w = (jj == 0) ? (nx - 1) : (jj - 1);
      const size_t c = (size_t)ii * nx + jj;
      double t[NSPEEDS];
      /* propagate, as a pull */
      t[0] = src[0 * plane + (size_t)ii * nx + jj];
      t[1] = src[1 * plane + (size_t)ii * nx + x_w];
      t[2] = src[2 * plane + (size_t)y_s * nx + jj];
      t[3] = src[3 * plane + (size_t)ii * nx + x_e];
      t[4] = src[4 * plane + (size_t)y_n * nx + jj];
      t[5] = src[5 * plane + (size_t)y_s * nx + x_w];
      t[6] = src[6 * plane + (size_t)y_s * nx + x_e];
      t[7] = src[7 * plane + (size_t)y_n * nx + x_e];
      t[8] = src[8 * plane + (size_t)y_n * nx + x_w];
      if (obstacles[c]) {
        /* rebound: mirror */
        dst[0 * plane + c] = t[0];
        dst[1 * plane + c] = t[3];
        dst[2 * plane + c] = t[4];
        dst[3 * plane + c] = t[1];
        dst[4 * plane + c] = t[2];
        dst[5 * plane + c] = t[7];
        dst[6 * plane + c] = t[8];
        dst[7 * plane + c] = t[5];
        dst[8 * plane + c] = t[6];
        continue;
      }
      /* collision */
      double local_density = 0.0;
      for (int kk = 0; kk < NSPEEDS; kk++) local_density += t[kk];
      const double u_x = (t[1] + t[5] + t[8] - (t[3] + t[6] + t[7])) / local_density;
      const double u_y = (t[2] + t[5] + t[6] - (t[4] + t[7] + t[8])) / local_density;
      const double u_sq = u_x * u_x + u_y * u_y;
      double u[NSPEEDS], d_equ[NSPEEDS], o[NSPEEDS];
      u[1] = u_x;
      u[2] = u_y;
      u[3] = -u_x;
      u[4] = -u_y;
      u[5] = u_x + u_y;
      u[6] = -u_x + u_y;
      u[7] = -u_x - u_y;
      u[8] = u_x - u_y;
      d_equ[0] = w0 * local_density * (1.0 - u_sq / (2.0 * c_sq));
      for (int kk = 1; kk < NSPEEDS; kk++) {
        const double w = kk <= 4 ? w1 : w2;
        d_equ[kk] = w * local_density
                    * (1.0 + u[kk] / c_sq + (u[kk] * u[kk]) / (2.0 * c_sq * c_sq) - u_sq / (2.0 * c_sq));
      }
      for (int kk = 0; kk < NSPEEDS; kk++) {
        o[kk] = t[kk] + p->omega * (d_equ[kk] - t[kk]);
        dst[kk * plane + c] = o[kk];
      }
      /* av_velocity on the post-collision cell */
      double dens2 = 0.0;
      for (int kk = 0; kk < NSPEEDS; kk++) dens2 += o[kk];
      const double v_x = (o[1] + o[5] + o[8] - (o[3] + o[6] + o[7])) / dens2;
      const double v_y = (o[2] + o[5] + o[6] - (o[4] + o[7] + o[8])) / dens2;
      tot += sqrt((v_x * v_x) + (v_y * v_y));
      free_cells++;
    }
    row_tot[ii] = tot;
  }
  double tot_u = 0.0;
  for (int ii = 0; ii < ny; ii++) tot_u += row_tot[ii];
  free(row_tot);
  return tot_u / (double)free_cells;
}

void oracle_f64_run(const oracle_params64 *p, double *cells, double *scratch, const int *obstacles,
                    int nsteps, double *av_vels)
{
  double *buf[2] = {cells, scratch};
  int curr = 0;
  for (int tt = 0; tt < nsteps; tt++) {
    accelerate_f64(p, buf[curr], obstacles);
    const double av = step_f64(p, buf[curr], buf[curr ^ 1], obstacles);
    if (av_vels) av_vels[tt] = av;
    curr ^= 1;
  }
  if (curr == 1) memcpy(cells, scratch, sizeof(double) * NSPEEDS * (size_t)p->nx * p->ny);
}

double oracle_f64_av_velocity(const oracle_params64 *p, const double *cells, const int *obstacles)
{
  const size_t plane = (size_t)p->nx * p->ny;
  double tot_u = 0.0;
  long free_cells = 0;
  for (size_t c = 0; c < plane; c++) {
    if (obstacles[c]) continue;
    double d = 0.0;
    for (int kk = 0; kk < NSPEEDS; kk++) d += cells[kk * plane + c];
    const double u_x = (cells[1 * plane + c] + cells[5 * plane + c] + cells[8 * plane + c]
                        - (cells[3 * plane + c] + cells[6 * plane + c] + cells[7 * plane + c])) / d;
    const double u_y = (cells[2 * plane + c] + cells[5 * plane + c] + cells[6 * plane + c]
                        - (cells[4 * plane + c] + cells[7 * plane + c] + cells[8 * plane + c])) / d;
    tot_u += sqrt((u_x * u_x) + (u_y * u_y));
    free_cells++;
  }
  return tot_u / (double)free_cells;
}

void oracle_f64_pressure(const oracle_params64 *p, const double *cells, const int *obstacles, double *pressure)
{
  const double c_sq = 1.0 / 3.0;
  const size_t plane = (size_t)p->nx * p->ny;
  for (size_t c = 0; c < plane; c++) {
    if (obstacles[c]) {
      pressure[c] = p->density * c_sq;
    } else {
      double d = 0.0;
      for (int kk = 0; kk < NSPEEDS; kk++) d += cells[kk * plane + c];
      pressure[c] = d * c_sq;
    }
  }
}

/* d2q9-bgk.c:789-831 in double, for regenerating the final_state goldens the
 * reference repository stripped (.MISSING_LARGE_BLOBS). */
void oracle_f64_final_state(const oracle_params64 *p, const double *cells, const int *obstacles,
                            double *u_x_out, double *u_y_out, double *u_out, double *pressure_out)
{
  const double c_sq = 1.0 / 3.0;
  const size_t plane = (size_t)p->nx * p->ny;
  for (size_t c = 0; c < plane; c++) {
    if (obstacles[c]) {
      u_x_out[c] = u_y_out[c] = u_out[c] = 0.0;
      pressure_out[c] = p->density * c_sq;
      continue;
    }
    double d = 0.0;
    for (int kk = 0; kk < NSPEEDS; kk++) d += cells[kk * plane + c];
    const double u_x = (cells[1 * plane + c] + cells[5 * plane + c] + cells[8 * plane + c]
                        - (cells[3 * plane + c] + cells[6 * plane + c] + cells[7 * plane + c])) / d;
    const double u_y = (cells[2 * plane + c] + cells[5 * plane + c] + cells[6 * plane + c]
                        - (cells[4 * plane + c] + cells[7 * plane + c] + cells[8 * plane + c])) / d;
    u_x_out[c] = u_x;
    u_y_out[c] = u_y;
    u_out[c] = sqrt((u_x * u_x) + (u_y * u_y));
    pressure_out[c] = d * c_sq;
  }
}

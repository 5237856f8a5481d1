/*
 * oracle/lbm_oracle.h — CPU restatement of the reference's D2Q9-BGK time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (the C host, the C-ABI
 * library, the CUDA kernels) includes, links or calls this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker / the CPU baseline.
 *
 * Parity status: PINNED.  The fp64 variant reproduces every golden the
 * reference ships for this path (check/128x128.*, check/128x256.*,
 * check/256x256.av_vels.dat, check/1024x1024.av_vels.dat) — see
 * tests/test_oracle_goldens.py; the fp32 variant (the arithmetic of
 * kernels.cl, multiply-adds contracted explicitly as OpenCL does by default)
 * passes the reference checker's 1 % gate against the same files.
 *
 * Data layout is the reference's (d2q9-bgk.c:73): SoA, nine planes of ny*nx
 * values, index sp*nx*ny + ii*nx + jj, ii = row (y), jj = column (x).
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors t_param, d2q9-bgk.c:81-92 */
typedef struct {
  float density;
  float accel;
  float omega;
  float free_cells_inv;
  int nx;
  int ny;
  int maxIters;
  int reynolds_dim;
} oracle_params;

/* ---- fp32: the arithmetic and operation order of kernels.cl ------------- */

/* kernels.cl:9-53 — body force on row ny-2, in place. */
void oracle_f32_accelerate(const oracle_params *p, float *cells, const int *obstacles);

/* kernels.cl:56-231 — fused pull-propagate + rebound + collision.  Returns the
 * step's average speed.  reference_order != 0 and nx % 128 == 0 reproduces the
 * reference's work-group tree (64 items x 2 cells, kernels.cl:202-229) and the
 * reduce kernel's tree (kernels.cl:234-290); otherwise the per-row sums are
 * added sequentially in row order. */
float oracle_f32_timestep(const oracle_params *p, const float *src, float *dst,
                          const int *obstacles, int reference_order);

/* d2q9-bgk.c:221-238 — nsteps of accelerate + timestep with ping-pong; the
 * final state is left in `cells` whatever the parity of nsteps; `scratch` is a
 * second buffer of the same size; av_vels gets nsteps values. */
void oracle_f32_run(const oracle_params *p, float *cells, float *scratch,
                    const int *obstacles, int nsteps, float *av_vels, int reference_order);

/* d2q9-bgk.c:396-442 — host av_velocity on a final state. */
float oracle_f32_av_velocity(const oracle_params *p, const float *cells, const int *obstacles);

/* d2q9-bgk.c:747-752 */
float oracle_f32_reynolds(const oracle_params *p, const float *cells, const int *obstacles);

/* d2q9-bgk.c:754-770 */
float oracle_f32_total_density(const oracle_params *p, const float *cells);

/* d2q9-bgk.c:789-831 — per-cell u_x, u_y, |u|, pressure (each ny*nx). */
void oracle_f32_final_state(const oracle_params *p, const float *cells, const int *obstacles,
                            float *u_x, float *u_y, float *u, float *pressure);

/* Row-slab form used by the world_size-2 CPU tests: `src`/`dst` hold
 * rows_local + 2 rows per plane (ghost row below = index 0, ghost row above =
 * index rows_local + 1, pitch nx).  Rows 1..rows_local are updated from src
 * into dst; obstacles is rows_local x nx.  accel_row_local is the local index
 * (0-based, without ghost) of global row ny-2, or -1 when another slab owns it.
 * row_sums (rows_local floats, may be NULL) receives each row's sum of speeds
 * (not yet scaled by free_cells_inv), summed left to right. */
void oracle_f32_slab_accelerate(const oracle_params *p, float *slab, const int *obstacles,
                                int rows_local, int accel_row_local);
void oracle_f32_slab_timestep(const oracle_params *p, const float *src, float *dst,
                              const int *obstacles, int rows_local, float *row_sums);

/* ---- fp64: the original serial equations (what produced the check/ goldens) ---- */

typedef struct {
  double density;
  double accel;
  double omega;
  int nx;
  int ny;
  int maxIters;
  int reynolds_dim;
} oracle_params64;

/* accelerate_flow -> propagate -> rebound -> collision -> av_velocity (after
 * collision), all in double, velocity form.  cells/scratch: SoA 9*ny*nx. */
void oracle_f64_run(const oracle_params64 *p, double *cells, double *scratch,
                    const int *obstacles, int nsteps, double *av_vels);
double oracle_f64_av_velocity(const oracle_params64 *p, const double *cells, const int *obstacles);
void oracle_f64_pressure(const oracle_params64 *p, const double *cells, const int *obstacles,
                         double *pressure);

void oracle_f64_final_state(const oracle_params64 *p, const double *cells, const int *obstacles,
                            double *u_x, double *u_y, double *u, double *pressure);

int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

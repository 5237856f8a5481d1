/*
 * host/lbm_io.c — see lbm_io.h.  Formats, messages and host maths follow the
 * reference host program (file:line cited per function); the code is new.
 */
#include "lbm_io.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NSPEEDS LBM_NSPEEDS

void die(const char *message, const int line, const char *file)
{
  fprintf(stderr, "Error at line %d of file %s:\n", line, file);
  fprintf(stderr, "%s\n", message);
  fflush(stderr);
  exit(EXIT_FAILURE);
}

void usage(const char *exe)
{
  fprintf(stderr, "Usage: %s <paramfile> <obstaclefile>\n", exe);
  exit(EXIT_FAILURE);
}

/* one "%d\n" / "%f\n" field of the params file, d2q9-bgk.c:466-492 */
static void read_int(FILE *fp, int *dst, const char *what)
{
  if (fscanf(fp, "%d\n", dst) != 1) {
    char msg[128];
    snprintf(msg, sizeof msg, "could not read param file: %s", what);
    die(msg, __LINE__, __FILE__);
  }
}

static void read_float(FILE *fp, float *dst, const char *what)
{
  if (fscanf(fp, "%f\n", dst) != 1) {
    char msg[128];
    snprintf(msg, sizeof msg, "could not read param file: %s", what);
    die(msg, __LINE__, __FILE__);
  }
}

static void *host_alloc(size_t bytes, int *pinned)
{
  /* pinned memory makes the timed upload/download run at full PCIe speed; fall
   * back to malloc when the library cannot provide it */
  void *p = getenv("LBM_NO_PINNED") ? NULL : lbm_host_alloc(bytes);
  *pinned = (p != NULL);
  return p ? p : malloc(bytes);
}

static int g_cells_pinned = 0, g_obstacles_pinned = 0;

void load_deck(const char *paramfile, const char *obstaclefile, lbm_params *params, float **cells_ptr,
               int **obstacles_ptr, float **av_vels_ptr)
{
  char message[1024];
  FILE *fp = fopen(paramfile, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input parameter file: %s", paramfile);
    die(message, __LINE__, __FILE__);
  }
  read_int(fp, &params->nx, "nx");
  read_int(fp, &params->ny, "ny");
  read_int(fp, &params->maxIters, "maxIters");
  read_int(fp, &params->reynolds_dim, "reynolds_dim");
  read_float(fp, &params->density, "density");
  read_float(fp, &params->accel, "accel");
  read_float(fp, &params->omega, "omega");
  fclose(fp);

  const size_t ncells = (size_t)params->nx * (size_t)params->ny;
  float *cells = (float *)host_alloc(sizeof(float) * NSPEEDS * ncells, &g_cells_pinned);
  if (cells == NULL) die("cannot allocate memory for cells", __LINE__, __FILE__);
  int *obstacles = (int *)host_alloc(sizeof(int) * ncells, &g_obstacles_pinned);
  if (obstacles == NULL) die("cannot allocate column memory for obstacles", __LINE__, __FILE__);

  /* uniform initial densities, obstacle cells included (d2q9-bgk.c:529-550) */
  const float w0 = params->density * 4.0f / 9.0f;
  const float w1 = params->density / 9.0f;
  const float w2 = params->density / 36.0f;
  for (int sp = 0; sp < NSPEEDS; sp++) {
    const float w = (sp == 0) ? w0 : (sp <= 4 ? w1 : w2);
    float *plane = cells + (size_t)sp * ncells;
    for (size_t c = 0; c < ncells; c++) plane[c] = w;
  }
  memset(obstacles, 0, sizeof(int) * ncells);

  fp = fopen(obstaclefile, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input obstacles file: %s", obstaclefile);
    die(message, __LINE__, __FILE__);
  }
  long free_cells = (long)ncells;
  int xx, yy, blocked, retval;
  while ((retval = fscanf(fp, "%d %d %d\n", &xx, &yy, &blocked)) != EOF) {
    /* checks and messages of d2q9-bgk.c:574-580 */
    if (retval != 3) die("expected 3 values per line in obstacle file", __LINE__, __FILE__);
    if (xx < 0 || xx > params->nx - 1) die("obstacle x-coord out of range", __LINE__, __FILE__);
    if (yy < 0 || yy > params->ny - 1) die("obstacle y-coord out of range", __LINE__, __FILE__);
    if (blocked != 1) die("obstacle blocked value should be 1", __LINE__, __FILE__);
    int *cell = &obstacles[(size_t)yy * params->nx + xx];
    if (!*cell) free_cells--; /* duplicates are not counted twice, d2q9-bgk.c:583-585 */
    *cell = blocked;
  }
  fclose(fp);
  params->free_cells_inv = 1.0f / free_cells; /* d2q9-bgk.c:591 */

  float *av_vels = (float *)malloc(sizeof(float) * (size_t)(params->maxIters > 0 ? params->maxIters : 1));
  if (av_vels == NULL) die("cannot allocate memory for av_vels", __LINE__, __FILE__);

  *cells_ptr = cells;
  *obstacles_ptr = obstacles;
  *av_vels_ptr = av_vels;
}

void free_deck(float *cells, int *obstacles, float *av_vels)
{
  if (g_cells_pinned) lbm_host_free(cells); else free(cells);
  if (g_obstacles_pinned) lbm_host_free(obstacles); else free(obstacles);
  free(av_vels);
}

/* density and velocity of one cell, the sums of d2q9-bgk.c:411-434 / :802-825 */
static void cell_moments(const float *cells, size_t ncells, size_t c, float *density, float *u_x, float *u_y)
{
  float local_density = 0.0f;
  for (int kk = 0; kk < NSPEEDS; kk++) local_density += cells[(size_t)kk * ncells + c];
  *u_x = (cells[1 * ncells + c] + cells[5 * ncells + c] + cells[8 * ncells + c]
          - cells[3 * ncells + c] - cells[6 * ncells + c] - cells[7 * ncells + c]) / local_density;
  *u_y = (cells[2 * ncells + c] + cells[5 * ncells + c] + cells[6 * ncells + c]
          - cells[4 * ncells + c] - cells[7 * ncells + c] - cells[8 * ncells + c]) / local_density;
  *density = local_density;
}

float av_velocity(const lbm_params *params, const float *cells, const int *obstacles)
{
  const size_t ncells = (size_t)params->nx * (size_t)params->ny;
  float tot_u = 0.0f;
  for (size_t c = 0; c < ncells; c++) {
    if (obstacles[c]) continue;
    float d, u_x, u_y;
    cell_moments(cells, ncells, c, &d, &u_x, &u_y);
    tot_u += sqrt((u_x * u_x) + (u_y * u_y)); /* double sqrt of a float, as d2q9-bgk.c:437 */
  }
  return tot_u * params->free_cells_inv;
}

float calc_reynolds(const lbm_params *params, const float *cells, const int *obstacles)
{
  const float viscosity = 1.0f / 6.0f * (2.0f / params->omega - 1.0f);
  return av_velocity(params, cells, obstacles) * params->reynolds_dim / viscosity;
}

/* One text row of final_state.dat into buf; returns the bytes written.  Line format of
 * d2q9-bgk.c:835: x y u_x u_y |u| pressure obstacle. */
static size_t format_row(char *buf, int ii, int nx, const float *u_x, const float *u_y, const float *u,
                         const float *pressure, const int *obstacles)
{
  size_t n = 0;
  for (int jj = 0; jj < nx; jj++)
    n += (size_t)sprintf(buf + n, "%d %d %.12E %.12E %.12E %.12E %d\n", jj, ii, u_x[jj], u_y[jj], u[jj],
                         pressure[jj], obstacles[jj]);
  return n;
}

#define ROW_BYTES_PER_CELL 128 /* > 2*11 + 4*20 + 7 */

int write_fields(const lbm_params *params, const float *u_x, const float *u_y, const float *u,
                 const float *pressure, const int *obstacles, const float *av_vels)
{
  const int nx = params->nx, ny = params->ny;
  const char *mode = getenv("LBM_FINAL_STATE");
  if (mode == NULL) mode = "text";

  if (strcmp(mode, "none") != 0) {
    FILE *fp = fopen(FINALSTATEFILE, "w");
    if (fp == NULL) die("could not open file output file", __LINE__, __FILE__);
    if (strcmp(mode, "binary") == 0) {
      /* for grids whose text file would be tens of GB: header "LBMFS1 nx ny\n" then u_x, u_y, |u|,
       * pressure (fp32) and the obstacle map (int32), each ny*nx, row-major */
      const size_t ncells = (size_t)nx * (size_t)ny;
      fprintf(fp, "LBMFS1 %d %d\n", nx, ny);
      fwrite(u_x, sizeof(float), ncells, fp);
      fwrite(u_y, sizeof(float), ncells, fp);
      fwrite(u, sizeof(float), ncells, fp);
      fwrite(pressure, sizeof(float), ncells, fp);
      fwrite(obstacles, sizeof(int), ncells, fp);
    } else {
      /* rows are formatted in parallel (exact printf semantics), written in order */
      int block = (int)((64u << 20) / ((size_t)nx * ROW_BYTES_PER_CELL));
      if (block < 1) block = 1;
      if (block > ny) block = ny;
      char *buf = (char *)malloc((size_t)block * nx * ROW_BYTES_PER_CELL);
      size_t *len = (size_t *)malloc(sizeof(size_t) * (size_t)block);
      if (buf == NULL || len == NULL) die("cannot allocate memory for the output buffer", __LINE__, __FILE__);
      for (int i0 = 0; i0 < ny; i0 += block) {
        const int nb = (ny - i0 < block) ? ny - i0 : block;
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nb; b++) {
          const size_t off = (size_t)(i0 + b) * nx;
          len[b] = format_row(buf + (size_t)b * nx * ROW_BYTES_PER_CELL, i0 + b, nx, u_x + off, u_y + off, u + off,
                              pressure + off, obstacles + off);
        }
        for (int b = 0; b < nb; b++) fwrite(buf + (size_t)b * nx * ROW_BYTES_PER_CELL, 1, len[b], fp);
      }
      free(buf);
      free(len);
    }
    fclose(fp);
  }

  FILE *fp = fopen(AVVELSFILE, "w");
  if (fp == NULL) die("could not open file output file", __LINE__, __FILE__);
  for (int ii = 0; ii < params->maxIters; ii++) fprintf(fp, "%d:\t%.12E\n", ii, av_vels[ii]); /* :850 */
  fclose(fp);
  return EXIT_SUCCESS;
}

int write_values(const lbm_params *params, const float *cells, const int *obstacles, const float *av_vels)
{
  /* d2q9-bgk.c:772-856 with the fields computed on the host, as the reference does */
  const float c_sq = 1.0f / 3.0f;
  const size_t ncells = (size_t)params->nx * (size_t)params->ny;
  float *f = (float *)malloc(sizeof(float) * 4 * ncells);
  if (f == NULL) die("cannot allocate memory for the output fields", __LINE__, __FILE__);
  float *u_x = f, *u_y = f + ncells, *u = f + 2 * ncells, *pressure = f + 3 * ncells;
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < ncells; c++) {
    u_x[c] = u_y[c] = u[c] = 0.0f;
    if (obstacles[c]) {
      pressure[c] = params->density * c_sq; /* d2q9-bgk.c:794-798 */
    } else {
      float local_density;
      cell_moments(cells, ncells, c, &local_density, &u_x[c], &u_y[c]);
      u[c] = sqrt((u_x[c] * u_x[c]) + (u_y[c] * u_y[c]));
      pressure[c] = local_density * c_sq;   /* d2q9-bgk.c:829-831 */
    }
  }
  const int rc = write_fields(params, u_x, u_y, u, pressure, obstacles, av_vels);
  free(f);
  return rc;
}

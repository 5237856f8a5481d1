/*
 * host/d2q9_bgk_main.c — the `d2q9-bgk <paramfile> <obstaclefile>` program.
 *
 * Same command line, input formats, output files and stdout lines as the
 * reference's main() (ag14774/OpenCL-Lattice-Boltzmann d2q9-bgk.c:165-280); the
 * device side is the C-ABI of include/lbm.h instead of OpenCL:
 *
 *   initialise() file half   :457-597   -> load_deck()            (lbm_io.c)
 *   initialise() OpenCL half :600-710   -> lbm_create()
 *   clEnqueueWriteBuffer x2  :200-209   -> lbm_upload()
 *   time loop                :221-238   -> lbm_run(maxIters)
 *   clFinish                 :239       -> lbm_sync()
 *   clEnqueueReadBuffer x2   :251-260   -> lbm_download_cells(), lbm_download_av_vels()
 *   finalise()               :715-744   -> lbm_destroy(), free_deck()
 *
 * The timed region is the reference's: upload + loop + sync + download (:196-263).
 * Environment: LBM_NGPUS=N splits the rows into N slabs on devices 0..N-1
 * (LBM_DEVICES=a,b,.. picks ordinals; the reference had OCL_DEVICE, :920-929);
 * LBM_QUIET=1 suppresses the extra throughput lines after the reference's five;
 * LBM_FINAL_STATE=text|binary|none picks the final_state.dat form (text = the reference's).
 */
#include <stdio.h>
#include <stdlib.h>
#include <sys/resource.h>
#include <sys/time.h>

#include "lbm.h"
#include "lbm_io.h"

static double wall_seconds(void)
{
  struct timeval t;
  gettimeofday(&t, NULL);
  return t.tv_sec + (t.tv_usec / 1000000.0);
}

/* checkError(), d2q9-bgk.c:858-866: message to stderr, exit(EXIT_FAILURE) */
static void check(int status, const char *op, const int line)
{
  if (status != 0) {
    fprintf(stderr, "LBM error during '%s' on line %d: %s\n", op, line, lbm_last_error());
    fflush(stderr);
    exit(EXIT_FAILURE);
  }
}

int main(int argc, char *argv[])
{
  lbm_params params;
  lbm_ctx *ctx = NULL;
  float *cells = NULL, *av_vels = NULL;
  int *obstacles = NULL;
  struct rusage ru;

  if (argc != 3) usage(argv[0]);

  load_deck(argv[1], argv[2], &params, &cells, &obstacles, &av_vels);
  check(lbm_create(&ctx, &params, 0), "creating device context", __LINE__);

  const double tic = wall_seconds();
  check(lbm_upload(ctx, cells, obstacles), "writing cells and obstacles data", __LINE__);
  float loop_ms = 0.0f;
  check(lbm_run_timed(ctx, params.maxIters, &loop_ms), "running time steps", __LINE__);
  check(lbm_sync(ctx), "waiting for queue", __LINE__);
  check(lbm_download_cells(ctx, cells), "reading cells data", __LINE__);
  check(lbm_download_av_vels(ctx, av_vels, params.maxIters), "reading av_vels data", __LINE__);
  const double toc = wall_seconds();

  getrusage(RUSAGE_SELF, &ru);
  const double usrtim = ru.ru_utime.tv_sec + (ru.ru_utime.tv_usec / 1000000.0);
  const double systim = ru.ru_stime.tv_sec + (ru.ru_stime.tv_usec / 1000000.0);

  /* d2q9-bgk.c:271-275, verbatim formats */
  printf("==done==\n");
  printf("Reynolds number:\t\t%.12E\n", calc_reynolds(&params, cells, obstacles));
  printf("Elapsed time:\t\t\t%.6lf (s)\n", toc - tic);
  printf("Elapsed user CPU time:\t\t%.6lf (s)\n", usrtim);
  printf("Elapsed system CPU time:\t%.6lf (s)\n", systim);

  if (!getenv("LBM_QUIET")) {
    lbm_info info;
    check(lbm_get_info(ctx, &info), "querying context", __LINE__);
    const double updates = (double)params.nx * params.ny * (double)params.maxIters;
    const double mlups_loop = loop_ms > 0.0f ? updates / (loop_ms * 1e-3) / 1e6 : 0.0;
    printf("Device loop time:\t\t%.6lf (s)\n", loop_ms * 1e-3);
    printf("MLUPS (device loop):\t\t%.1f\n", mlups_loop);
    printf("MLUPS (elapsed):\t\t%.1f\n", updates / (toc - tic) / 1e6);
    printf("Effective bandwidth:\t\t%.1f GB/s (72 B per cell update)\n", mlups_loop * 72.0 / 1e3);
    printf("Kernel:\t\t\t\t%s x %d slab(s), %lld launches\n", info.kernel_name, info.nslabs,
           info.kernel_launches);
  }

  /* write_values(), d2q9-bgk.c:276.  The per-cell fields come from the device output stage
   * (bit-identical to the host maths); LBM_HOST_FIELDS=1 computes them on the host instead. */
  if (getenv("LBM_HOST_FIELDS")) {
    write_values(&params, cells, obstacles, av_vels);
  } else {
    const size_t ncells = (size_t)params.nx * (size_t)params.ny;
    float *fields = (float *)malloc(sizeof(float) * 4 * ncells);
    if (fields == NULL) die("cannot allocate memory for the output fields", __LINE__, __FILE__);
    check(lbm_download_final_state(ctx, fields, fields + ncells, fields + 2 * ncells, fields + 3 * ncells),
          "reading final state fields", __LINE__);
    write_fields(&params, fields, fields + ncells, fields + 2 * ncells, fields + 3 * ncells, obstacles, av_vels);
    free(fields);
  }
  lbm_destroy(ctx);
  free_deck(cells, obstacles, av_vels);
  return EXIT_SUCCESS;
}

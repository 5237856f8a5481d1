/*
 * host/lbm_io.h — file formats and host-only maths of the d2q9-bgk CLI.
 *
 * The drop-in surface of the reference host program (ag14774/OpenCL-Lattice-
 * Boltzmann, d2q9-bgk.c): the 7-line .params file (:457-495), the "x y 1"
 * obstacle list (:553-591), the uniform initial state (:529-550), the two output
 * files (:772-856), the Reynolds print (:747-752) and the die()/usage() error
 * convention (:868-880).  No device code here.
 */
#ifndef LBM_IO_H
#define LBM_IO_H

#include "lbm.h"

#define FINALSTATEFILE "final_state.dat" /* d2q9-bgk.c:69 */
#define AVVELSFILE "av_vels.dat"         /* d2q9-bgk.c:70 */

/* prints "Error at line L of file F:\nmessage\n" to stderr and exits (d2q9-bgk.c:868-874) */
void die(const char *message, const int line, const char *file);
/* prints "Usage: exe <paramfile> <obstaclefile>" to stderr and exits (d2q9-bgk.c:876-880) */
void usage(const char *exe);

/* file half of initialise(): params, host lattice (pinned when the library can
 * provide it), obstacle map, av_vels array; sets params->free_cells_inv */
void load_deck(const char *paramfile, const char *obstaclefile, lbm_params *params, float **cells,
               int **obstacles, float **av_vels);
void free_deck(float *cells, int *obstacles, float *av_vels);

/* av_velocity() on the host (d2q9-bgk.c:396-442) and calc_reynolds() (:747-752) */
float av_velocity(const lbm_params *params, const float *cells, const int *obstacles);
float calc_reynolds(const lbm_params *params, const float *cells, const int *obstacles);

/* write_values() (d2q9-bgk.c:772-856): final_state.dat and av_vels.dat in the CWD, fields computed
 * on the host from the cells as the reference does */
int write_values(const lbm_params *params, const float *cells, const int *obstacles, const float *av_vels);

/* Same files from fields already computed (lbm_download_final_state).  Rows are formatted in
 * parallel.  LBM_FINAL_STATE=text (default) | binary | none selects the final_state.dat form. */
int write_fields(const lbm_params *params, const float *u_x, const float *u_y, const float *u,
                 const float *pressure, const int *obstacles, const float *av_vels);

#endif

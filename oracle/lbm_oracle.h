/*
 * lbm_oracle.h -- CPU restatement of the reference's D2Q9-BGK hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and there only as the checker or as the
 * timed CPU baseline.  The shipped solver (liblbm_b200.so) never links,
 * loads or falls back to this code.
 *
 * Every function cites the reference lines it follows (paths relative to the
 * upstream repo thorbenlouw/lbm-graphcore).  The oracle is PINNED: tests/
 * check it (a) text-identically against the unmodified reference program
 * main/LastChance.cpp built into oracle/_ref/ with the same -ffp-contract=off
 * flag, (b) against the reference's shipped double-precision goldens
 * (check/(128x128|128x256|256x256|1024x1024).*.dat) under check.py's 1 % rule,
 * and (c) against the known-answer vectors of test/codelets/main.cpp.
 *
 * Layout conventions are the reference's: cells are AoS float[ny][nx][9]
 * (main/include/LatticeBoltzmannUtils.hpp:18,125-135), obstacles are one byte
 * per cell, row-major, index ii + jj*nx (main/include/LbmParams.hpp:88-90),
 * speeds numbered 0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW, 7 SW, 8 SE
 * (main/LastChance.cpp:7-13).
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_NSPEEDS 9

/* main/LastChance.cpp:65-77 (t_param), main/include/LbmParams.hpp:16-24 */
typedef struct {
  int nx;
  int ny;
  int max_iters;
  int reynolds_dim;
  float density;
  float accel;
  float omega;
} oracle_params;

/* how tot_u is accumulated over the grid within one timestep */
enum {
  ORACLE_ACCUM_FLOAT = 0,  /* literal reference: one fp32 scalar, row-major (LastChance.cpp:190,262) */
  ORACLE_ACCUM_DOUBLE = 1  /* same terms, fp64 accumulator (fp32 saturates at 16384^2)   */
};

/* ---- file formats --------------------------------------------------------------------------- */

/* main/LastChance.cpp:352-387: seven lines nx ny maxIters reynolds_dim density accel omega.
 * Returns 0, or a negative code naming the field that failed (-1 open, -2..-8 fields). */
int oracle_load_params(const char *path, oracle_params *p);

/* main/LastChance.cpp:452-484: lines "x y 1"; range-checked; duplicates idempotent.
 * obst must hold ny*nx bytes and is zeroed first.  Returns #lines read or <0 on error. */
long oracle_load_obstacles(const char *path, int nx, int ny, uint8_t *obst);

/* main/LastChance.cpp:486-493 */
long oracle_count_free_cells(const uint8_t *obst, int nx, int ny);

/* main/LastChance.cpp:429-450 == lbm::Cells::initialise LatticeBoltzmannUtils.hpp:137-157 */
void oracle_init_cells(const oracle_params *p, float *cells);

/* ---- per-cell stages (exposed for the known-answer tests) ------------------------------------ */

/* main/LastChance.cpp:213-223: out = {s0,s3,s4,s1,s2,s7,s8,s5,s6} */
void oracle_rebound_cell(const float s[9], float out[9]);

/* main/LastChance.cpp:225-251: BGK relaxation of one (already propagated) cell.
 * Returns u_sq (pre-collision velocity squared) so the caller can accumulate sqrtf(u_sq). */
float oracle_collide_cell(const float s[9], float omega, float out[9]);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* One-off guarded accelerate of row ny-2 before step 0: main/LastChance.cpp:158-183 */
void oracle_accelerate_flow(const oracle_params *p, float *cells, const uint8_t *obst);

/* One fused timestep src -> dst over global rows [row_begin,row_end):
 * pull-propagate (LastChance.cpp:196-211), rebound (:213-223), collision (:225-251),
 * unguarded next-step accelerate on row ny-2 (:198,:254-261), tot_u (:262).
 * Returns tot_u over those rows (float mode: the fp32 running sum widened to double).
 * nthreads>1 parallelises rows with OpenMP (double mode only). */
double oracle_timestep_rows(const oracle_params *p, const float *src, float *dst, const uint8_t *obst,
                            int row_begin, int row_end, int accum_mode, int nthreads);

/* The whole loop of main/LastChance.cpp:156-268: optional first accelerate, then `iters` steps
 * ping-ponging cells <-> tmp (:187-188).  av_vels[t] = tot_u / (float)free_cells (:266).
 * Returns a pointer to whichever of cells/tmp holds the final state. */
float *oracle_run(const oracle_params *p, float *cells, float *tmp, const uint8_t *obst, int iters,
                  float *av_vels, int accum_mode, int nthreads, int do_first_accel);

/* Same loop, single-threaded, recording av_vels under BOTH accumulation modes in one pass (the cells
 * do not depend on the mode).  Note the fp64 variant sums row partials (row-major rows, each row
 * accumulated left to right) -- within 1e-15 relative of any other fp64 order. */
float *oracle_run_both(const oracle_params *p, float *cells, float *tmp, const uint8_t *obst, int iters,
                       float *av_vels_f32sum, float *av_vels_f64sum, int do_first_accel);

/* main/LastChance.cpp:290-339 */
float oracle_av_velocity(const oracle_params *p, const float *cells, const uint8_t *obst);
/* main/LastChance.cpp:529-534 == lbm::reynoldsNumber LatticeBoltzmannUtils.hpp:202-205 */
float oracle_reynolds(const oracle_params *p, float av_velocity);
/* main/LastChance.cpp:536-552 */
float oracle_total_density(const oracle_params *p, const float *cells);

/* ---- writers (reference C format, "%.12E") --------------------------------------------------- */

/* main/LastChance.cpp:566-617.  NOTE: the reference prints the obstacle flag of the TRANSPOSED
 * cell (obstacles[ii*nx + jj], :614) in column 7; we print the flag of the cell itself, as the
 * C++ writer lbm::writeResults does (LatticeBoltzmannUtils.hpp:273).  check.py reads columns
 * 0,1,5 only.  transposed_flag=1 reproduces the reference column byte for byte. */
int oracle_write_final_state(const char *path, const oracle_params *p, const float *cells,
                             const uint8_t *obst, int transposed_flag);
/* main/LastChance.cpp:619-632 */
int oracle_write_av_vels(const char *path, const float *av_vels, int n);

#ifdef __cplusplus
}
#endif
#endif

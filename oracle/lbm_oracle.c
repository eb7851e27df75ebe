/*
 * lbm_oracle.c -- CPU restatement of the reference's D2Q9-BGK timestep.
 *
 * TEST INFRASTRUCTURE ONLY (see lbm_oracle.h).  Build with -ffp-contract=off so
 * every fp32 operation below is a single IEEE-754 rounding in source order: the
 * CUDA kernels are compiled with -fmad=false and IEEE div/sqrt and reproduce
 * these cells bit for bit.
 *
 * Follows main/LastChance.cpp:156-268 of thorbenlouw/lbm-graphcore operation
 * by operation (same association order, same constants); twin statements of the
 * same arithmetic are main/LbmCpu.cpp:121-218 (timestep2) and
 * main/codelets/D2Q9Codelets.cpp:94-191 (lbmKernel).
 */
#include "lbm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CELL(base, nx, ii, jj) ((base) + ((size_t)(ii) + (size_t)(jj) * (size_t)(nx)) * ORACLE_NSPEEDS)

/* ---------------------------------------------------------------------------------------------- */
/* file formats                                                                                   */
/* ---------------------------------------------------------------------------------------------- */

int oracle_load_params(const char *path, oracle_params *p) {
  FILE *fp = fopen(path, "r");
  if (!fp) return -1;
  int rc = 0;
  /* main/LastChance.cpp:361-387: one value per line, ints then floats */
  if (fscanf(fp, "%d\n", &p->nx) != 1) rc = -2;
  else if (fscanf(fp, "%d\n", &p->ny) != 1) rc = -3;
  else if (fscanf(fp, "%d\n", &p->max_iters) != 1) rc = -4;
  else if (fscanf(fp, "%d\n", &p->reynolds_dim) != 1) rc = -5;
  else if (fscanf(fp, "%f\n", &p->density) != 1) rc = -6;
  else if (fscanf(fp, "%f\n", &p->accel) != 1) rc = -7;
  else if (fscanf(fp, "%f\n", &p->omega) != 1) rc = -8;
  fclose(fp);
  return rc;
}

long oracle_load_obstacles(const char *path, int nx, int ny, uint8_t *obst) {
  memset(obst, 0, (size_t)nx * (size_t)ny); /* LastChance.cpp:452-459 */
  FILE *fp = fopen(path, "r");
  if (!fp) return -1;
  long lines = 0;
  int xx, yy, blocked, got;
  /* LastChance.cpp:471-484 */
  while ((got = fscanf(fp, "%d %d %d\n", &xx, &yy, &blocked)) != EOF) {
    if (got != 3) { fclose(fp); return -2; }
    if (xx < 0 || xx > nx - 1) { fclose(fp); return -3; }
    if (yy < 0 || yy > ny - 1) { fclose(fp); return -4; }
    if (blocked != 1) { fclose(fp); return -5; }
    obst[(size_t)xx + (size_t)yy * (size_t)nx] = 1;
    ++lines;
  }
  fclose(fp);
  return lines;
}

long oracle_count_free_cells(const uint8_t *obst, int nx, int ny) {
  long n = 0;
  for (size_t i = 0; i < (size_t)nx * (size_t)ny; ++i) n += !obst[i];
  return n;
}

void oracle_init_cells(const oracle_params *p, float *cells) {
  /* LastChance.cpp:429-431 */
  const float w0 = p->density * 4.f / 9.f;
  const float w1 = p->density / 9.f;
  const float w2 = p->density / 36.f;
  const size_t n = (size_t)p->nx * (size_t)p->ny;
  for (size_t c = 0; c < n; ++c) {
    float *s = cells + c * ORACLE_NSPEEDS;
    s[0] = w0;
    s[1] = s[2] = s[3] = s[4] = w1;
    s[5] = s[6] = s[7] = s[8] = w2;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* per-cell stages                                                                                */
/* ---------------------------------------------------------------------------------------------- */

static inline void rebound_cell(const float s[9], float out[9]) {
  /* LastChance.cpp:215-223: every population is sent back where it came from */
  static const int opposite[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
  for (int k = 0; k < 9; ++k) out[k] = s[opposite[k]];
}

static inline float collide_cell(const float s[9], float omega, float out[9]) {
  const float one_minus_omega = 1 - omega; /* LastChance.cpp:387 */
  /* LastChance.cpp:228 -- strictly left to right */
  const float rho = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7] + s[8];
  /* :230, :232 */
  const float u_x = (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / rho;
  const float u_y = (s[2] + s[5] + s[6] - (s[4] + s[7] + s[8])) / rho;
  /* :235-242 */
  const float u_sq = u_x * u_x + u_y * u_y;
  const float c_sq = 1.00f - u_sq * 1.50f;
  const float ld0 = 4.00f / 9.00f * rho * omega;
  const float ld1 = rho / 9.00f * omega;
  const float ld2 = rho / 36.00f * omega;
  const float u_s = u_x + u_y;
  const float u_d = -u_x + u_y;
  /* :244-252 */
  out[0] = s[0] * one_minus_omega + ld0 * c_sq;
  out[1] = s[1] * one_minus_omega + ld1 * ((4.50f * u_x) * (2.00f / 3.00f + u_x) + c_sq);
  out[2] = s[2] * one_minus_omega + ld1 * ((4.50f * u_y) * (2.00f / 3.00f + u_y) + c_sq);
  out[3] = s[3] * one_minus_omega + ld1 * ((-4.50f * u_x) * (2.00f / 3.00f - u_x) + c_sq);
  out[4] = s[4] * one_minus_omega + ld1 * ((-4.50f * u_y) * (2.00f / 3.00f - u_y) + c_sq);
  out[5] = s[5] * one_minus_omega + ld2 * ((4.50f * u_s) * (2.00f / 3.00f + u_s) + c_sq);
  out[6] = s[6] * one_minus_omega + ld2 * ((4.50f * u_d) * (2.00f / 3.00f + u_d) + c_sq);
  out[7] = s[7] * one_minus_omega + ld2 * ((-4.50f * u_s) * (2.00f / 3.00f - u_s) + c_sq);
  out[8] = s[8] * one_minus_omega + ld2 * ((-4.50f * u_d) * (2.00f / 3.00f - u_d) + c_sq);
  return u_sq;
}

void oracle_rebound_cell(const float s[9], float out[9]) { rebound_cell(s, out); }
float oracle_collide_cell(const float s[9], float omega, float out[9]) { return collide_cell(s, omega, out); }

/* ---------------------------------------------------------------------------------------------- */
/* hot path                                                                                       */
/* ---------------------------------------------------------------------------------------------- */

void oracle_accelerate_flow(const oracle_params *p, float *cells, const uint8_t *obst) {
  /* LastChance.cpp:158-159 */
  const float w1 = p->density * p->accel / 9.f;
  const float w2 = p->density * p->accel / 36.f;
  const int jj = p->ny - 2; /* :166 */
  if (jj < 0) return; /* ny == 1: the reference would index row -1 (undefined); defined here as a no-op */
  for (int ii = 0; ii < p->nx; ++ii) {
    float *s = CELL(cells, p->nx, ii, jj);
    /* :169-172 guard: free cell and no population would go non-positive */
    if (!obst[(size_t)ii + (size_t)jj * (size_t)p->nx] && (s[3] - w1) > 0.f && (s[6] - w2) > 0.f &&
        (s[7] - w2) > 0.f) {
      s[1] += w1; /* :175-181 */
      s[5] += w2;
      s[8] += w2;
      s[3] -= w1;
      s[6] -= w2;
      s[7] -= w2;
    }
  }
}

/* one row of the fused sweep; returns the row's fp64 partial and threads the fp32 running sum */
static double sweep_row(const oracle_params *p, const float *src, float *dst, const uint8_t *obst, int jj,
                        float w1, float w2, float *tot_f) {
  const int nx = p->nx, ny = p->ny;
  /* LastChance.cpp:196-198 */
  const int y_n = (jj + 1) % ny;
  const int y_s = (jj == 0) ? (jj + ny - 1) : (jj - 1);
  const float accel = (jj == ny - 2) ? 1.00f : 0.00f;
  double tot_d = 0.0;
  float tf = *tot_f;
  for (int ii = 0; ii < nx; ++ii) {
    const int x_e = (ii + 1) % nx;                      /* :199 */
    const int x_w = (ii == 0) ? (ii + nx - 1) : (ii - 1); /* :200 */
    float s[9], out[9];
    /* :203-211 pull: population k arrives from the neighbour opposite to direction k */
    s[0] = CELL(src, nx, ii, jj)[0];
    s[1] = CELL(src, nx, x_w, jj)[1];
    s[2] = CELL(src, nx, ii, y_s)[2];
    s[3] = CELL(src, nx, x_e, jj)[3];
    s[4] = CELL(src, nx, ii, y_n)[4];
    s[5] = CELL(src, nx, x_w, y_s)[5];
    s[6] = CELL(src, nx, x_e, y_s)[6];
    s[7] = CELL(src, nx, x_e, y_n)[7];
    s[8] = CELL(src, nx, x_w, y_n)[8];
    float *d = CELL(dst, nx, ii, jj);
    if (obst[(size_t)ii + (size_t)jj * (size_t)nx]) { /* :213 */
      rebound_cell(s, out);
      for (int k = 0; k < 9; ++k) d[k] = out[k];
    } else {
      const float u_sq = collide_cell(s, p->omega, out);
      /* :253-261 store + unguarded accelerate of the NEXT step on row ny-2 */
      d[0] = out[0];
      d[1] = out[1] + accel * w1;
      d[2] = out[2];
      d[3] = out[3] - accel * w1;
      d[4] = out[4];
      d[5] = out[5] + accel * w2;
      d[6] = out[6] - accel * w2;
      d[7] = out[7] - accel * w2;
      d[8] = out[8] + accel * w2;
      const float u = sqrtf(u_sq); /* :262 */
      tf += u;
      tot_d += (double)u;
    }
  }
  *tot_f = tf;
  return tot_d;
}

double oracle_timestep_rows(const oracle_params *p, const float *src, float *dst, const uint8_t *obst,
                            int row_begin, int row_end, int accum_mode, int nthreads) {
  const float w1 = p->density * p->accel / 9.f;
  const float w2 = p->density * p->accel / 36.f;
  if (accum_mode == ORACLE_ACCUM_FLOAT) {
    float tot_f = 0.00f; /* :190 */
    for (int jj = row_begin; jj < row_end; ++jj) (void)sweep_row(p, src, dst, obst, jj, w1, w2, &tot_f);
    return (double)tot_f;
  }
  double tot = 0.0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for reduction(+ : tot) schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int jj = row_begin; jj < row_end; ++jj) {
    float unused = 0.f;
    tot += sweep_row(p, src, dst, obst, jj, w1, w2, &unused);
  }
  return tot;
}

float *oracle_run(const oracle_params *p, float *cells, float *tmp, const uint8_t *obst, int iters,
                  float *av_vels, int accum_mode, int nthreads, int do_first_accel) {
  const long free_cells = oracle_count_free_cells(obst, p->nx, p->ny);
  if (do_first_accel) oracle_accelerate_flow(p, cells, obst);
  float *cur = cells, *nxt = tmp;
  for (int t = 0; t < iters; ++t) {
    const double tot = oracle_timestep_rows(p, cur, nxt, obst, 0, p->ny, accum_mode, nthreads);
    /* :266 -- fp32 quotient of the (rounded) total by the fp32 free-cell count */
    av_vels[t] = (float)tot / (float)free_cells;
    float *sw = cur; cur = nxt; nxt = sw; /* :187-188 */
  }
  return cur;
}

float *oracle_run_both(const oracle_params *p, float *cells, float *tmp, const uint8_t *obst, int iters,
                       float *av_vels_f32sum, float *av_vels_f64sum, int do_first_accel) {
  const float w1 = p->density * p->accel / 9.f;
  const float w2 = p->density * p->accel / 36.f;
  const long free_cells = oracle_count_free_cells(obst, p->nx, p->ny);
  if (do_first_accel) oracle_accelerate_flow(p, cells, obst);
  float *cur = cells, *nxt = tmp;
  for (int t = 0; t < iters; ++t) {
    float tot_f = 0.00f;
    double tot_d = 0.0;
    for (int jj = 0; jj < p->ny; ++jj) tot_d += sweep_row(p, cur, nxt, obst, jj, w1, w2, &tot_f);
    av_vels_f32sum[t] = tot_f / (float)free_cells;
    av_vels_f64sum[t] = (float)tot_d / (float)free_cells;
    float *sw = cur; cur = nxt; nxt = sw;
  }
  return cur;
}

float oracle_av_velocity(const oracle_params *p, const float *cells, const uint8_t *obst) {
  int tot_cells = 0;
  float tot_u = 0.f;
  for (int jj = 0; jj < p->ny; ++jj) {
    for (int ii = 0; ii < p->nx; ++ii) {
      if (obst[(size_t)ii + (size_t)jj * (size_t)p->nx]) continue; /* :304 */
      const float *s = CELL(cells, p->nx, ii, jj);
      float rho = 0.f;
      for (int k = 0; k < 9; ++k) rho += s[k]; /* :307-312 */
      const float u_x = (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / rho; /* :315-321 */
      const float u_y = (s[2] + s[5] + s[6] - (s[4] + s[7] + s[8])) / rho; /* :323-329 */
      tot_u += sqrtf((u_x * u_x) + (u_y * u_y));                          /* :331 */
      ++tot_cells;
    }
  }
  return tot_u / (float)tot_cells; /* :338 */
}

float oracle_reynolds(const oracle_params *p, float av_velocity) {
  const float viscosity = 1.f / 6.f * (2.f / p->omega - 1.f); /* :531 */
  return av_velocity * p->reynolds_dim / viscosity;           /* :533 */
}

float oracle_total_density(const oracle_params *p, const float *cells) {
  float total = 0.f;
  const size_t n = (size_t)p->nx * (size_t)p->ny * ORACLE_NSPEEDS;
  for (size_t i = 0; i < n; ++i) total += cells[i]; /* :540-549, same visiting order */
  return total;
}

/* ---------------------------------------------------------------------------------------------- */
/* writers                                                                                        */
/* ---------------------------------------------------------------------------------------------- */

int oracle_write_final_state(const char *path, const oracle_params *p, const float *cells,
                             const uint8_t *obst, int transposed_flag) {
  FILE *fp = fopen(path, "w");
  if (!fp) return -1;
  const float c_sq = 1.f / 3.f; /* :557 */
  for (int jj = 0; jj < p->ny; ++jj) {
    for (int ii = 0; ii < p->nx; ++ii) {
      float u_x, u_y, u, pressure;
      const size_t idx = (size_t)ii + (size_t)jj * (size_t)p->nx;
      if (obst[idx]) { /* :576-580 */
        u_x = u_y = u = 0.f;
        pressure = p->density * c_sq;
      } else { /* :583-609 */
        const float *s = CELL(cells, p->nx, ii, jj);
        float rho = 0.f;
        for (int k = 0; k < 9; ++k) rho += s[k];
        u_x = (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / rho;
        u_y = (s[2] + s[5] + s[6] - (s[4] + s[7] + s[8])) / rho;
        u = sqrtf((u_x * u_x) + (u_y * u_y));
        pressure = rho * c_sq;
      }
      /* the reference's index ii*nx+jj can leave the array when ny < nx; fall back to the cell itself */
      const size_t tidx = (size_t)ii * (size_t)p->nx + (size_t)jj;
      const int flag = (transposed_flag && tidx < (size_t)p->nx * (size_t)p->ny) ? obst[tidx] : obst[idx];
      fprintf(fp, "%d %d %.12E %.12E %.12E %.12E %d\n", ii, jj, u_x, u_y, u, pressure, flag); /* :614 */
    }
  }
  fclose(fp);
  return 0;
}

int oracle_write_av_vels(const char *path, const float *av_vels, int n) {
  FILE *fp = fopen(path, "w");
  if (!fp) return -1;
  for (int i = 0; i < n; ++i) fprintf(fp, "%d:\t%.12E\n", i, av_vels[i]); /* :628-631 */
  fclose(fp);
  return 0;
}

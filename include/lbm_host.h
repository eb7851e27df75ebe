/*
 * lbm_host.h -- C ABI of liblbm_host.so: the host-side problem I/O of the B200 LBM solver
 * (no CUDA).  It exposes, for tests and for bindings, the C++ helpers of
 * lbm-graphcore_b200/csrc/lbm_host.hpp, which mirror the reference's header-only
 * main/include/LbmParams.hpp and main/include/LatticeBoltzmannUtils.hpp.
 * All functions return 0 on success and a negative value on failure unless stated.
 */
#ifndef LBM_HOST_H
#define LBM_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lbm::Params::fromFile -- main/include/LbmParams.hpp:28-58 */
int lbm_host_load_params(const char *path, long long *nx, long long *ny, long long *max_iters,
                         long long *reynolds_dim, float *density, float *accel, float *omega);

/* lbm::Obstacles::fromFile -- main/include/LbmParams.hpp:92-123.  obstacles: ny*nx bytes, zeroed first.
 * Returns the number of solid cells, or negative on failure (unreadable file, coordinate out of range). */
long long lbm_host_load_obstacles(const char *path, long long nx, long long ny, uint8_t *obstacles);

/* lbm::Cells::initialise -- main/include/LatticeBoltzmannUtils.hpp:137-157.  cells: ny*nx*9 floats. */
int lbm_host_init_cells(long long nx, long long ny, float density, float *cells);

/* lbm::Cells::averageVelocity -- LatticeBoltzmannUtils.hpp:160-194 */
float lbm_host_average_velocity(long long nx, long long ny, const float *cells, const uint8_t *obstacles);
/* lbm::Cells::totalDensity -- LatticeBoltzmannUtils.hpp:196-198 */
float lbm_host_total_density(long long nx, long long ny, const float *cells);
/* lbm::reynoldsNumber -- LatticeBoltzmannUtils.hpp:202-205 */
float lbm_host_reynolds_number(float omega, long long reynolds_dim, float average_velocity);

/* lbm::writeAverageVelocities -- LatticeBoltzmannUtils.hpp:208-219 */
int lbm_host_write_av_vels(const char *path, const float *av_vels, long long n);
/* lbm::writeResults -- LatticeBoltzmannUtils.hpp:221-281 */
int lbm_host_write_final_state(const char *path, long long nx, long long ny, float density, const float *cells,
                               const uint8_t *obstacles);

/* The same file from fields the device already evaluated (lbm_download_fields of lbm_b200.h): fields is
 * float[ny][nx][4] = u_x, u_y, u, pressure.  Lines are formatted on `threads` host threads (0 = all cores). */
int lbm_host_write_final_state_fields(const char *path, long long nx, long long ny, const float *fields,
                                      const uint8_t *obstacles, int threads);

/* the synthetic channel of SURVEY.md 8(d): rows [row_begin, row_begin+row_count) of an nx*ny grid */
int lbm_host_synthetic_obstacles(long long nx, long long ny, uint64_t seed, float solid_prob, long long row_begin,
                                 long long row_count, uint8_t *obstacles);

#ifdef __cplusplus
}
#endif
#endif /* LBM_HOST_H */

// lbm_host.cpp -- C ABI (include/lbm_host.h) over the header-only host library lbm_host.hpp.
#include "../../include/lbm_host.h"

#include "lbm_host.hpp"

namespace {
lbm::Params make_params(long long nx, long long ny, float density) {
  lbm::Params p;
  p.nx = static_cast<size_t>(nx);
  p.ny = static_cast<size_t>(ny);
  p.density = density;
  return p;
}
// borrow caller memory as Cells/Obstacles without copying twice: the C ABI works on raw arrays, so
// the few functions below re-state the loops over raw pointers via the same per-cell helpers.
}  // namespace

extern "C" {

int lbm_host_load_params(const char *path, long long *nx, long long *ny, long long *max_iters,
                         long long *reynolds_dim, float *density, float *accel, float *omega) {
  if (!path) return -1;
  const auto p = lbm::Params::fromFile(path);
  if (!p) return -2;
  if (nx) *nx = static_cast<long long>(p->nx);
  if (ny) *ny = static_cast<long long>(p->ny);
  if (max_iters) *max_iters = static_cast<long long>(p->maxIters);
  if (reynolds_dim) *reynolds_dim = static_cast<long long>(p->reynolds_dim);
  if (density) *density = p->density;
  if (accel) *accel = p->accel;
  if (omega) *omega = p->omega;
  return 0;
}

long long lbm_host_load_obstacles(const char *path, long long nx, long long ny, uint8_t *obstacles) {
  if (!path || !obstacles || nx < 1 || ny < 1) return -1;
  const auto o = lbm::Obstacles::fromFile(static_cast<size_t>(nx), static_cast<size_t>(ny), path);
  if (!o) return -2;
  std::memcpy(obstacles, o->getData(), static_cast<size_t>(nx) * static_cast<size_t>(ny));
  return static_cast<long long>(static_cast<size_t>(nx) * static_cast<size_t>(ny) - o->freeCells());
}

int lbm_host_init_cells(long long nx, long long ny, float density, float *cells) {
  if (!cells || nx < 1 || ny < 1) return -1;
  lbm::Cells c(static_cast<size_t>(nx), static_cast<size_t>(ny));
  c.initialise(make_params(nx, ny, density));
  std::memcpy(cells, c.data.data(), c.data.size() * sizeof(float));
  return 0;
}

float lbm_host_average_velocity(long long nx, long long ny, const float *cells, const uint8_t *obstacles) {
  lbm::Cells c(static_cast<size_t>(nx), static_cast<size_t>(ny));
  std::memcpy(c.data.data(), cells, c.data.size() * sizeof(float));
  lbm::Obstacles o(static_cast<size_t>(nx), static_cast<size_t>(ny));
  std::memcpy(o.getData(), obstacles, static_cast<size_t>(nx) * static_cast<size_t>(ny));
  return c.averageVelocity(make_params(nx, ny, 0.f), o);
}

float lbm_host_total_density(long long nx, long long ny, const float *cells) {
  lbm::Cells c(static_cast<size_t>(nx), static_cast<size_t>(ny));
  std::memcpy(c.data.data(), cells, c.data.size() * sizeof(float));
  return c.totalDensity();
}

float lbm_host_reynolds_number(float omega, long long reynolds_dim, float average_velocity) {
  lbm::Params p;
  p.omega = omega;
  p.reynolds_dim = static_cast<size_t>(reynolds_dim);
  return lbm::reynoldsNumber(p, average_velocity);
}

int lbm_host_write_av_vels(const char *path, const float *av_vels, long long n) {
  if (!path || (!av_vels && n > 0) || n < 0) return -1;
  return lbm::writeAverageVelocities(path, std::vector<float>(av_vels, av_vels + n)) ? 0 : -2;
}

int lbm_host_write_final_state(const char *path, long long nx, long long ny, float density, const float *cells,
                               const uint8_t *obstacles) {
  if (!path || !cells || !obstacles || nx < 1 || ny < 1) return -1;
  lbm::Cells c(static_cast<size_t>(nx), static_cast<size_t>(ny));
  std::memcpy(c.data.data(), cells, c.data.size() * sizeof(float));
  lbm::Obstacles o(static_cast<size_t>(nx), static_cast<size_t>(ny));
  std::memcpy(o.getData(), obstacles, static_cast<size_t>(nx) * static_cast<size_t>(ny));
  return lbm::writeResults(path, make_params(nx, ny, density), o, c) ? 0 : -2;
}

int lbm_host_write_final_state_fields(const char *path, long long nx, long long ny, const float *fields,
                                      const uint8_t *obstacles, int threads) {
  if (!path || !fields || !obstacles || nx < 1 || ny < 1 || threads < 0) return -1;
  lbm::Obstacles o(static_cast<size_t>(nx), static_cast<size_t>(ny));
  std::memcpy(o.getData(), obstacles, static_cast<size_t>(nx) * static_cast<size_t>(ny));
  return lbm::writeResultsFromFields(path, make_params(nx, ny, 0.f), o, fields, static_cast<unsigned>(threads)) ? 0 : -2;
}

int lbm_host_synthetic_obstacles(long long nx, long long ny, uint64_t seed, float solid_prob, long long row_begin,
                                 long long row_count, uint8_t *obstacles) {
  if (!obstacles || nx < 1 || ny < 1 || row_begin < 0 || row_count < 0 || row_begin + row_count > ny) return -1;
  for (long long r = 0; r < row_count; ++r)
    for (long long x = 0; x < nx; ++x)
      obstacles[r * nx + x] = lbm::syntheticSolid(seed, static_cast<size_t>(x), static_cast<size_t>(row_begin + r),
                                                  static_cast<size_t>(nx), static_cast<size_t>(ny), solid_prob)
                                  ? 1
                                  : 0;
  return 0;
}

}  // extern "C"

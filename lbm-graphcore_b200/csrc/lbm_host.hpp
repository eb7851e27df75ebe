// lbm_host.hpp -- host-side problem I/O for the B200 LBM solver (C++17, no CUDA).
//
// The reference keeps this in two header-only files, main/include/LbmParams.hpp (Params, Obstacles)
// and main/include/LatticeBoltzmannUtils.hpp (Cells, reynoldsNumber, writeAverageVelocities,
// writeResults).  This file offers the same names with the same meaning and the same file formats,
// written from scratch: parsing is strict where the reference has undefined behaviour, and the
// writers format into a large buffer instead of flushing an ofstream once per line (the reference's
// writer dominates its own end-to-end time at 1024x1024).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <optional>
#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace lbm {

constexpr unsigned NumSpeeds = 9u;
// speed numbering of main/LastChance.cpp:7-13 / LatticeBoltzmannUtils.hpp:20-22
enum SpeedIndexes { Middle, East, North, West, South, NorthEast, NorthWest, SouthWest, SouthEast };

// ---- Params (main/include/LbmParams.hpp:16-65) -----------------------------------------------------
struct Params {
  size_t nx = 0, ny = 0, maxIters = 0, reynolds_dim = 0;
  float density = 0.f, accel = 0.f, omega = 0.f;

  // Seven lines: nx, ny, maxIters, reynolds_dim (unsigned), density, accel, omega (float)
  // (LbmParams.hpp:28-58).  The reference lets stoul/stof throw on a short or malformed file; here
  // that is a std::nullopt with a message on stderr, like the could-not-open case (:55-56).
  static std::optional<Params> fromFile(const std::string &filename) {
    std::ifstream file(filename);
    if (!file.is_open()) {
      std::cerr << "Could not read parameters from " << filename << std::endl;
      return std::nullopt;
    }
    static const char *names[7] = {"nx", "ny", "maxIters", "reynolds_dim", "density", "accel", "omega"};
    Params p;
    size_t *uints[4] = {&p.nx, &p.ny, &p.maxIters, &p.reynolds_dim};
    float *floats[3] = {&p.density, &p.accel, &p.omega};
    for (int field = 0; field < 7; ++field) {
      std::string line;
      bool ok = static_cast<bool>(std::getline(file, line));
      if (ok) {
        const char *begin = line.c_str();
        char *end = nullptr;
        if (field < 4) {
          while (*begin == ' ' || *begin == '\t') ++begin;
          ok = (*begin >= '0' && *begin <= '9');  // stoul would accept "-1" and wrap; a size cannot be negative
          if (ok) *uints[field] = std::strtoul(begin, &end, 10);
        } else {
          *floats[field - 4] = std::strtof(begin, &end);
          ok = end != begin;
        }
      }
      if (!ok) {
        std::cerr << "Could not parse parameter '" << names[field] << "' (line " << field + 1 << ") of " << filename
                  << std::endl;
        return std::nullopt;
      }
    }
    return p;
  }
};

// ---- Obstacles (main/include/LbmParams.hpp:67-128) -------------------------------------------------
class Obstacles {
 public:
  size_t nx = 0, ny = 0;

  Obstacles(size_t nx_, size_t ny_) : nx(nx_), ny(ny_), data_(nx_ * ny_, 0) {}

  // one byte per cell, non-zero = solid; same interchange as the reference's bool[] (LbmParams.hpp:75-77)
  uint8_t *getData() { return data_.data(); }
  const uint8_t *getData() const { return data_.data(); }
  bool at(size_t x, size_t y) const { return data_[y * nx + x] != 0; }  // LbmParams.hpp:88-90
  void set(size_t x, size_t y, bool solid = true) { data_[y * nx + x] = solid ? 1 : 0; }

  size_t freeCells() const {  // total_free_cells, main/LastChance.cpp:486-493
    size_t n = 0;
    for (uint8_t b : data_) n += (b == 0);
    return n;
  }

  // Lines "x y 1" (LbmParams.hpp:92-123).  As in the reference a line whose third field is not 1 ends
  // parsing with "Malformed line: obstacle must be 1" and the cells read so far are kept (:101-106).
  // Unlike the reference (which writes out of bounds, :115) coordinates are range-checked like
  // main/LastChance.cpp:476-480; an out-of-range coordinate is an error.  Duplicates are idempotent.
  static std::optional<Obstacles> fromFile(size_t nx, size_t ny, const std::string &filename) {
    std::ifstream file(filename);
    if (!file.is_open()) {
      std::cerr << "Could not read parameters from " << filename << std::endl;  // sic, LbmParams.hpp:120
      return std::nullopt;
    }
    Obstacles obstacles(nx, ny);
    std::string line;
    while (std::getline(file, line)) {
      long x = 0, y = 0, flag = 0;
      const int got = std::sscanf(line.c_str(), "%ld %ld %ld", &x, &y, &flag);
      if (got != 3 || flag != 1) {
        if (got > 0 || line.find_first_not_of(" \t\r") != std::string::npos)
          std::cerr << "Malformed line: obstacle must be 1" << std::endl;
        break;
      }
      if (x < 0 || y < 0 || static_cast<size_t>(x) >= nx || static_cast<size_t>(y) >= ny) {
        std::cerr << "Obstacle (" << x << "," << y << ") is outside the " << nx << "x" << ny << " grid" << std::endl;
        return std::nullopt;
      }
      obstacles.set(static_cast<size_t>(x), static_cast<size_t>(y));
    }
    return obstacles;
  }

 private:
  std::vector<uint8_t> data_;
};

// ---- synthetic channel, host twin of lbm_init_synthetic ------------------------------
inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline bool syntheticSolid(uint64_t seed, size_t x, size_t y, size_t nx, size_t ny, float solid_prob) {
  if (y == 0 || y == ny - 1) return true;
  if (y + 2 == ny) return false;
  const uint32_t threshold24 = static_cast<uint32_t>(static_cast<double>(solid_prob) * 16777216.0);
  return static_cast<uint32_t>(splitmix64(seed + y * nx + x) >> 40) < threshold24;
}

// ---- Cells (main/include/LatticeBoltzmannUtils.hpp:125-199) ------------------------------------------
class Cells {
 public:
  size_t nx, ny;
  std::vector<float> data;  // AoS [ny][nx][9], LatticeBoltzmannUtils.hpp:18

  Cells(size_t nx_, size_t ny_) : nx(nx_), ny(ny_), data(nx_ * ny_ * NumSpeeds, 0.f) {}

  float &at(size_t x, size_t y, unsigned speed) { return data[(x + y * nx) * NumSpeeds + speed]; }
  float at(size_t x, size_t y, unsigned speed) const { return data[(x + y * nx) * NumSpeeds + speed]; }

  // every cell (obstacles too) starts at rest equilibrium (LatticeBoltzmannUtils.hpp:137-157)
  void initialise(const Params &params) {
    const float w0 = params.density * 4.f / 9.f;
    const float w1 = params.density / 9.f;
    const float w2 = params.density / 36.f;
    const float proto[NumSpeeds] = {w0, w1, w1, w1, w1, w2, w2, w2, w2};
    for (size_t c = 0; c < nx * ny; ++c) std::memcpy(&data[c * NumSpeeds], proto, sizeof proto);
  }

  // LatticeBoltzmannUtils.hpp:160-194 (masked form: obstacle cells contribute 0 and are not counted)
  float averageVelocity(const Params &params, const Obstacles &obstacles) const {
    int tot_cells = 0;
    float tot_u = 0.f;
    for (size_t jj = 0; jj < params.ny; ++jj) {
      for (size_t ii = 0; ii < params.nx; ++ii) {
        const int mask = 1 - static_cast<int>(obstacles.at(ii, jj));
        const float *s = &data[(ii + jj * nx) * NumSpeeds];
        float local_density = 0.f;
        for (unsigned k = 0; k < NumSpeeds; ++k) local_density += s[k];
        const float u_x = static_cast<float>(mask) * (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / local_density;
        const float u_y = static_cast<float>(mask) * ((s[2] + s[5] + s[6]) - (s[4] + s[7] + s[8])) / local_density;
        tot_u += sqrtf((u_x * u_x) + (u_y * u_y));
        tot_cells += mask;
      }
    }
    return tot_u / static_cast<float>(tot_cells);
  }

  float totalDensity() const {  // LatticeBoltzmannUtils.hpp:196-198
    float total = 0.0f;
    for (float v : data) total += v;
    return total;
  }
};

// LatticeBoltzmannUtils.hpp:202-205
inline float reynoldsNumber(const Params &params, float average_velocity) {
  const float viscosity = 1.f / 6.f * (2.f / params.omega - 1.f);
  return average_velocity * params.reynolds_dim / viscosity;
}

namespace detail {
// append-only text buffer flushed in large writes
class TextSink {
 public:
  explicit TextSink(const std::string &path) : fp_(std::fopen(path.c_str(), "w")) { buf_.reserve(kFlush + 512); }
  ~TextSink() { close(); }
  bool ok() const { return fp_ != nullptr && !failed_; }
  char *reserve(size_t n) {
    if (buf_.size() + n > kFlush) flush();
    const size_t old = buf_.size();
    buf_.resize(old + n);
    return buf_.data() + old;
  }
  void commit(size_t reserved, size_t used) { buf_.resize(buf_.size() - reserved + used); }
  bool close() {
    if (fp_) {
      flush();
      if (std::fclose(fp_) != 0) failed_ = true;
      fp_ = nullptr;
    }
    return !failed_;
  }

 private:
  static constexpr size_t kFlush = 1u << 22;
  void flush() {
    if (fp_ && !buf_.empty() && std::fwrite(buf_.data(), 1, buf_.size(), fp_) != buf_.size()) failed_ = true;
    buf_.clear();
  }
  FILE *fp_;
  bool failed_ = false;
  std::vector<char> buf_;
};
}  // namespace detail

// "<i>:\t<%.12e>\n" per timestep (LatticeBoltzmannUtils.hpp:208-219: std::scientific, precision 12)
inline bool writeAverageVelocities(const std::string &filename, const std::vector<float> &av_vels) {
  detail::TextSink out(filename);
  if (!out.ok()) return false;
  for (size_t i = 0; i < av_vels.size(); ++i) {
    char *p = out.reserve(64);
    const int n = std::snprintf(p, 64, "%zu:\t%.12e\n", i, static_cast<double>(av_vels[i]));
    out.commit(64, static_cast<size_t>(n));
  }
  return out.close();
}

// per-cell fields of the final state (LatticeBoltzmannUtils.hpp:240-267)
struct CellFields {
  float u_x, u_y, u, pressure;
};
inline CellFields cellFields(const float *s, bool solid, float density) {
  const float c_sq = 1.f / 3.f;  // square of the speed of sound
  if (solid) return {0.f, 0.f, 0.f, density * c_sq};
  float local_density = 0.f;
  for (unsigned k = 0; k < NumSpeeds; ++k) local_density += s[k];
  const float u_x = (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / local_density;
  const float u_y = (s[2] + s[5] + s[6] - (s[4] + s[7] + s[8])) / local_density;
  return {u_x, u_y, sqrtf((u_x * u_x) + (u_y * u_y)), local_density * c_sq};
}

namespace detail {
// Formats rows [0, ny) of final_state.dat with `line(ii, jj)` -> CellFields on all host cores: the rows are cut into
// chunks, each wave of chunks is formatted in parallel into per-thread buffers and written in row order.  The
// reference streams one line at a time through an ofstream with std::endl (a flush per line,
// LatticeBoltzmannUtils.hpp:274); at 1024x1024 that costs more than the whole GPU solve.
template <typename FieldsAt>
bool writeFinalStateRows(const std::string &filename, const Params &params, const Obstacles &obstacles, FieldsAt fields_at,
                         unsigned threads = 0) {
  std::FILE *fp = std::fopen(filename.c_str(), "w");
  if (!fp) return false;
  if (threads == 0) threads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  const size_t rows_per_chunk = std::max<size_t>(1, (size_t(1) << 16) / std::max<size_t>(params.nx, 1));
  const size_t chunks = (params.ny + rows_per_chunk - 1) / rows_per_chunk;
  std::vector<std::string> bufs(std::min<size_t>(threads, std::max<size_t>(chunks, 1)));
  bool ok = true;
  auto format_chunk = [&](size_t chunk, std::string &out) {
    out.clear();
    const size_t j0 = chunk * rows_per_chunk, j1 = std::min(params.ny, j0 + rows_per_chunk);
    out.reserve((j1 - j0) * params.nx * 96);
    char line[192];
    for (size_t jj = j0; jj < j1; ++jj) {
      for (size_t ii = 0; ii < params.nx; ++ii) {
        const bool solid = obstacles.at(ii, jj);
        const CellFields f = fields_at(ii, jj, solid);
        const int n = std::snprintf(line, sizeof line, "%zu %zu %.12e %.12e %.12e %.12e %d\n", ii, jj, static_cast<double>(f.u_x),
                                    static_cast<double>(f.u_y), static_cast<double>(f.u), static_cast<double>(f.pressure),
                                    solid ? 1 : 0);
        out.append(line, static_cast<size_t>(n));
      }
    }
  };
  for (size_t base = 0; base < chunks && ok; base += bufs.size()) {
    const size_t n = std::min(bufs.size(), chunks - base);
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n; ++t) pool.emplace_back(format_chunk, base + t, std::ref(bufs[t]));
    format_chunk(base, bufs[0]);
    for (auto &th : pool) th.join();
    for (size_t t = 0; t < n && ok; ++t) ok = std::fwrite(bufs[t].data(), 1, bufs[t].size(), fp) == bufs[t].size();
  }
  if (std::fclose(fp) != 0) ok = false;
  return ok;
}
}  // namespace detail

// "<ii> <jj> <u_x> <u_y> <u> <pressure> <obstacle>\n", jj outer, ii inner (LatticeBoltzmannUtils.hpp:221-281)
inline bool writeResults(const std::string &filename, const Params &params, const Obstacles &obstacles, const Cells &cells,
                         unsigned threads = 0) {
  return detail::writeFinalStateRows(
      filename, params, obstacles,
      [&](size_t ii, size_t jj, bool solid) { return cellFields(&cells.data[(ii + jj * params.nx) * NumSpeeds], solid, params.density); },
      threads);
}

// Same file from fields already evaluated on the device (lbm_download_fields): float[ny][nx][4] = u_x, u_y, u, pressure
inline bool writeResultsFromFields(const std::string &filename, const Params &params, const Obstacles &obstacles,
                                   const float *fields, unsigned threads = 0) {
  return detail::writeFinalStateRows(
      filename, params, obstacles,
      [&](size_t ii, size_t jj, bool) {
        const float *f = fields + (ii + jj * params.nx) * 4;
        return CellFields{f[0], f[1], f[2], f[3]};
      },
      threads);
}

}  // namespace lbm

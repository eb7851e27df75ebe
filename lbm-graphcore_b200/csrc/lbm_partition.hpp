// lbm_partition.hpp -- how a grid is cut across GPUs, and what that cut costs (C++17, no CUDA).
//
// One definition of the partition arithmetic for the library (lbm_capi.cu), the driver (lbm_runner --stats) and the
// lbm_partition tool.  The reference cuts its grid for IPUs, tiles and workers in main/include/StructuredGridUtils.hpp
// (partitionForIpus :472-561) and reports on the quality of the cut in main/TileMappingStats.cpp:54-101 (wasted tiles /
// workers, load balance, max speed-up, block sizes) and main/VisualiseTileMapping.cpp:66-205 (a picture of it).  The units
// differ on a B200 -- a GPU takes the place of an IPU, a 128 x 8-cell tile (= one CTA of the sweep kernel) that of an IPU
// tile's block, a slot for a resident CTA (148 SMs x 4) that of a worker -- the questions are the same.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace lbm {
namespace partition {

constexpr int kTileCols = 128, kTileRows = 8;      // lbm_kernels.cuh kTileW x kTileH
constexpr int kSmCount = 148, kCtasPerSm = 4;      // B200; __launch_bounds__(256, 4) of the sweep kernel
constexpr int kBytesPerCellStep = 72;              // 9 fp32 read + 9 fp32 written

// part i of n cells cut into `parts`: [begin, begin + count).  The same floor arithmetic everywhere.
inline void range_of(int n, int parts, int i, int &begin, int &count) {
  begin = (int)((long long)i * n / parts);
  count = (int)((long long)(i + 1) * n / parts) - begin;
}

// The reference's choice of a block grid for 1/2/4/8/16 devices (partitionForIpus, StructuredGridUtils.hpp:479-527):
// cut along the direction that divides more evenly; 4 -> 2 x 2, 8 -> 4 x 2 or 2 x 4, 16 -> 4 x 4.  px = blocks per row
// (the reference's numCols), py = blocks per column (numRows).  Other counts: row slabs.
inline void grid_for_gpus(int nx, int ny, int gpus, int &px, int &py) {
  const float row_imbalance = (float)(ny % gpus) / (float)ny, col_imbalance = (float)(nx % gpus) / (float)nx;
  px = 1;
  py = gpus;
  if (gpus == 2) { if (row_imbalance < col_imbalance) { py = 2; px = 1; } else { py = 1; px = 2; } }
  else if (gpus == 4) { py = 2; px = 2; }
  else if (gpus == 8) { if (row_imbalance < col_imbalance) { py = 4; px = 2; } else { py = 2; px = 4; } }
  else if (gpus == 16) { py = 4; px = 4; }
  px = std::min(px, nx);
  py = std::min(py, ny);
}

struct Block {
  int bx = 0, by = 0, col0 = 0, nxl = 0, row0 = 0, nyl = 0;
  long long cells() const { return (long long)nxl * nyl; }
  int tile_cols() const { return (nxl + kTileCols - 1) / kTileCols; }
  int tile_rows() const { return (nyl + 2 + kTileRows - 1) / kTileRows; }  // storage rows incl. the two ghost rows
  long long tiles() const { return (long long)tile_cols() * tile_rows(); }
  // bytes this block stores into its neighbours per timestep: three planes of its top and bottom row, and with ghost
  // columns three planes of its west and east column plus the four corner populations
  long long halo_bytes(bool ghost_columns) const {
    return 2ll * 3 * nxl * 4 + (ghost_columns ? 2ll * 3 * nyl * 4 + 4 * 4 : 0);
  }
};

inline std::vector<Block> blocks(int nx, int ny, int px, int py) {
  std::vector<Block> out;
  for (int by = 0; by < py; ++by)
    for (int bx = 0; bx < px; ++bx) {
      Block b;
      b.bx = bx;
      b.by = by;
      range_of(nx, px, bx, b.col0, b.nxl);
      range_of(ny, py, by, b.row0, b.nyl);
      out.push_back(b);
    }
  return out;
}

// The columns of main/TileMappingStats.cpp's CSV, restated for GPUs
struct Stats {
  int gpus = 0, px = 0, py = 0;
  long long largest = 0, smallest = 0;    // cells of the busiest / idlest block
  double average = 0;                     // cells per block
  double load_balance_pct = 0;            // average / largest (TileMappingStats.cpp:86)
  double max_speedup = 0;                 // all cells / largest block (:93)
  long long max_tiles = 0;                // CTAs per timestep on the busiest GPU
  double waves = 0;                       // of 148 x 4 resident CTAs
  double tail_efficiency_pct = 0;         // waves / ceil(waves): how full the last wave is
  long long wasted_slots = 0;             // resident-CTA slots idle in the last wave, summed over GPUs (cf. wastedWorkers)
  double tile_padding_pct = 0;            // cells of partially filled tiles and ghost rows, relative to owned cells
  long long max_halo_bytes = 0;           // per timestep, busiest GPU
  double halo_to_lattice_pct = 0;         // halo bytes / (72 B x cells) on that GPU
};

inline Stats stats(int nx, int ny, int px, int py, bool ghost_columns) {
  Stats st;
  st.px = px;
  st.py = py;
  st.gpus = px * py;
  const std::vector<Block> bl = blocks(nx, ny, px, py);
  st.smallest = bl[0].cells();
  long long padded = 0, slots = (long long)kSmCount * kCtasPerSm;
  for (const Block &b : bl) {
    st.largest = std::max(st.largest, b.cells());
    st.smallest = std::min(st.smallest, b.cells());
    st.max_tiles = std::max(st.max_tiles, b.tiles());
    st.max_halo_bytes = std::max(st.max_halo_bytes, b.halo_bytes(ghost_columns));
    padded += b.tiles() * kTileCols * kTileRows;
    const long long last = b.tiles() % slots;
    st.wasted_slots += last ? slots - last : 0;
  }
  const double all = (double)nx * ny;
  st.average = all / st.gpus;
  st.load_balance_pct = st.average / (double)st.largest * 100.0;
  st.max_speedup = all / (double)st.largest;
  st.waves = (double)st.max_tiles / (double)slots;
  st.tail_efficiency_pct = st.waves / (double)((st.max_tiles + slots - 1) / slots) * 100.0;
  st.tile_padding_pct = ((double)padded - all) / all * 100.0;
  st.halo_to_lattice_pct = (double)st.max_halo_bytes / ((double)kBytesPerCellStep * (double)st.largest) * 100.0;
  return st;
}

inline const char *csv_header() {
  return "numGpus,width,height,px,py,wastedSlots,loadBalance,maxSpeedup,aveBlockSize,largestBlockSize,smallestBlockSize,"
         "maxTilesPerGpu,wavesPerStep,tailEfficiency,tilePadding,haloBytesPerStep,haloToLattice";
}
inline std::string csv_row(int nx, int ny, const Stats &s) {
  char buf[512];
  std::snprintf(buf, sizeof buf, "%d,%d,%d,%d,%d,%lld,%.4g,%.6g,%.8g,%lld,%lld,%lld,%.4g,%.4g,%.4g,%lld,%.4g", s.gpus, nx, ny, s.px, s.py,
                s.wasted_slots, s.load_balance_pct, s.max_speedup, s.average, s.largest, s.smallest, s.max_tiles, s.waves,
                s.tail_efficiency_pct, s.tile_padding_pct, s.max_halo_bytes, s.halo_to_lattice_pct);
  return buf;
}

}  // namespace partition
}  // namespace lbm

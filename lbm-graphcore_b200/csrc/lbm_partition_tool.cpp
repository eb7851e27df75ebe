// lbm_partition -- how would a grid of a given size be cut across B200 GPUs, and how good is the cut?
//
// The B200 counterpart of the reference's two partition tools: `tile_mapping_stats` (main/TileMappingStats.cpp: one CSV
// row per grid size with wasted tiles / workers, load balance, max speed-up and block sizes, either for one size or for
// random samples of a size range) and `viz_tile_map` (main/VisualiseTileMapping.cpp: a picture of the mapping).  No GPU is
// needed: the arithmetic is lbm_partition.hpp, which the solver itself uses.
//
//   lbm_partition stats --width W --height H [--num-gpus N] [--grid PXxPY] [--obstacles FILE]
//   lbm_partition stats --sample --num-samples K --min-width A --max-width B --min-height C --max-height D [--num-gpus N]
//   lbm_partition map   --width W --height H [--num-gpus N] [--grid PXxPY] [--obstacles FILE] -o OUT.png
//
// --grid PXxPY asks for PX x PY blocks with ghost columns (lbm_create_grid); without it the grid is cut into N row slabs
// (lbm_create); `--grid auto` applies the reference's rule for 2/4/8/16 devices (partitionForIpus).
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <optional>
#include <string>
#include <vector>

#include "lbm_host.hpp"
#include "lbm_partition.hpp"

namespace {
namespace part = lbm::partition;

struct Options {
  std::string command, out, obstacles, grid;
  int gpus = 1, width = 0, height = 0;
  bool sample = false;
  int samples = 100, min_w = 128, max_w = 16384, min_h = 128, max_h = 16384;
  unsigned seed = 1;
};

void usage(const char *exe) {
  std::cerr << exe << " - Show how a grid of the given size would be mapped to B200 GPUs, and how well\n"
            << "  stats | map            what to produce: a CSV row per grid size, or a picture\n"
            << "      --num-gpus arg     number of GPUs to target (alias: --num-ipus) (default: 1)\n"
            << "      --grid arg         PXxPY blocks with ghost columns, or 'auto' (the reference's rule); default: row slabs\n"
            << "      --width arg        width of input\n"
            << "      --height arg       height of input\n"
            << "      --obstacles arg    obstacles file: adds the free-cell load balance (stats) / shades solids (map)\n"
            << "      --sample           stats for random sizes instead of one size\n"
            << "      --num-samples arg  (default: 100)\n"
            << "      --min-width, --max-width, --min-height, --max-height arg   the sampled range\n"
            << "      --seed arg         seed of the sampler (default: 1)\n"
            << "  -o, --output arg       filename of the picture (PNG)" << std::endl;
}

bool parse(int argc, char **argv, Options &o) {
  if (argc < 2) return false;
  o.command = argv[1];
  if (o.command != "stats" && o.command != "map") return false;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i], v;
    const auto eq = a.find('=');
    bool inl = false;
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); inl = true; }
    auto value = [&]() -> const char * { if (inl) return v.c_str(); return i + 1 < argc ? argv[++i] : nullptr; };
    auto number = [&](int &dst) { const char *s = value(); if (!s) return false; char *e = nullptr; dst = (int)std::strtol(s, &e, 10); return e != s && *e == 0; };
    const char *s = nullptr;
    if (a == "--sample") o.sample = true;
    else if (a == "--num-gpus" || a == "--num-ipus") { if (!number(o.gpus)) return false; }
    else if (a == "--width") { if (!number(o.width)) return false; }
    else if (a == "--height") { if (!number(o.height)) return false; }
    else if (a == "--num-samples") { if (!number(o.samples)) return false; }
    else if (a == "--min-width") { if (!number(o.min_w)) return false; }
    else if (a == "--max-width") { if (!number(o.max_w)) return false; }
    else if (a == "--min-height") { if (!number(o.min_h)) return false; }
    else if (a == "--max-height") { if (!number(o.max_h)) return false; }
    else if (a == "--seed") { int sd = 0; if (!number(sd)) return false; o.seed = (unsigned)sd; }
    else if (a == "--grid") { if (!(s = value())) return false; o.grid = s; }
    else if (a == "--obstacles") { if (!(s = value())) return false; o.obstacles = s; }
    else if (a == "-o" || a == "--output") { if (!(s = value())) return false; o.out = s; }
    else return false;
  }
  if (o.gpus < 1 || o.gpus > 4096) return false;
  if (o.command == "map") return o.width > 0 && o.height > 0 && !o.out.empty();
  if (o.sample) return o.samples > 0 && o.min_w > 0 && o.max_w >= o.min_w && o.min_h > 0 && o.max_h >= o.min_h;
  return o.width > 0 && o.height > 0;
}

// the block grid the options ask for; false when it cannot hold the grid
bool choose(const Options &o, int nx, int ny, int &px, int &py, bool &ghost_columns) {
  ghost_columns = !o.grid.empty();
  if (o.grid.empty()) { px = 1; py = o.gpus; }
  else if (o.grid == "auto") part::grid_for_gpus(nx, ny, o.gpus, px, py);
  else {
    char *end = nullptr;
    px = (int)std::strtol(o.grid.c_str(), &end, 10);
    if (!end || (*end != 'x' && *end != 'X')) return false;
    py = (int)std::strtol(end + 1, nullptr, 10);
  }
  return px >= 1 && py >= 1 && px <= nx && py <= ny;
}

// ---- a minimal PNG writer (8-bit RGB, stored deflate blocks) ------------------------------------------
uint32_t crc32_of(const uint8_t *p, size_t n, uint32_t crc = 0) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}
void put32(std::vector<uint8_t> &v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
void chunk(std::vector<uint8_t> &png, const char *type, const std::vector<uint8_t> &data) {
  put32(png, (uint32_t)data.size());
  std::vector<uint8_t> body(type, type + 4);
  body.insert(body.end(), data.begin(), data.end());
  png.insert(png.end(), body.begin(), body.end());
  put32(png, crc32_of(body.data(), body.size()));
}
bool write_png(const std::string &path, int w, int h, const std::vector<uint8_t> &rgb) {
  std::vector<uint8_t> raw;  // scanlines, each with filter byte 0
  raw.reserve((size_t)h * (1 + 3 * (size_t)w));
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);
    raw.insert(raw.end(), rgb.begin() + (size_t)y * 3 * w, rgb.begin() + (size_t)(y + 1) * 3 * w);
  }
  std::vector<uint8_t> z = {0x78, 0x01};
  uint32_t a = 1, b = 0;
  for (uint8_t c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
  for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
    const size_t n = std::min<size_t>(65535, raw.size() - off);
    z.push_back(off + n >= raw.size() ? 1 : 0);
    z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
    z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    if (raw.empty()) break;
  }
  put32(z, (b << 16) | a);
  std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
  put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
  ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
  chunk(png, "IHDR", ihdr);
  chunk(png, "IDAT", z);
  chunk(png, "IEND", {});
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char *>(png.data()), (std::streamsize)png.size());
  return f.good();
}

int run_map(const Options &o) {
  int px, py;
  bool ghost;
  if (!choose(o, o.width, o.height, px, py, ghost)) {
    std::cerr << "Couldn't fit the problem on the " << o.gpus << " GPUs." << std::endl;
    return EXIT_FAILURE;
  }
  std::cout << "Visualising a grid that is " << o.height << " rows by " << o.width << " cols in size, cut into " << px << " x " << py
            << (ghost ? " blocks." : " row slabs.") << std::endl;
  std::optional<lbm::Obstacles> obst;
  if (!o.obstacles.empty()) {
    obst = lbm::Obstacles::fromFile((size_t)o.width, (size_t)o.height, o.obstacles);
    if (!obst.has_value()) { std::cerr << "Could not parse obstacles file" << std::endl; return EXIT_FAILURE; }
  }
  // one pixel per `scale` x `scale` cells, at most 2048 pixels a side; row 0 of the grid at the bottom of the picture
  const int scale = std::max(1, (std::max(o.width, o.height) + 2047) / 2048);
  const int w = (o.width + scale - 1) / scale, h = (o.height + scale - 1) / scale;
  static const uint8_t colours[8][3] = {{0x1F, 0x3A, 0x5F}, {0x8C, 0x2F, 0x39}, {0xE0, 0x9F, 0x3E}, {0x2A, 0x9D, 0x8F},
                                        {0x45, 0x7B, 0x9D}, {0xE7, 0x6F, 0x51}, {0x8A, 0xB1, 0x7D}, {0x6D, 0x59, 0x7A}};
  std::vector<uint8_t> rgb((size_t)w * h * 3);
  const std::vector<part::Block> bl = part::blocks(o.width, o.height, px, py);
  for (const part::Block &b : bl) {
    const uint8_t *col = colours[(b.bx + 3 * b.by) % 8];
    for (int y = b.row0; y < b.row0 + b.nyl; y += scale)
      for (int x = b.col0; x < b.col0 + b.nxl; x += scale) {
        // tile lines (128 x 8 cells: one CTA of the sweep) lighten the block colour; block borders are white; solids black
        const int lx = x - b.col0, ly = y - b.row0 + 1;  // +1: storage row 0 is the ghost row
        const bool border = lx < scale || ly - 1 < scale || lx >= b.nxl - scale || ly - 1 >= b.nyl - scale;
        const bool tile_line = scale <= 4 && ((lx % part::kTileCols) < scale || (scale == 1 && ly % part::kTileRows == 0));
        bool solid = false;
        if (obst) {
          int n = 0, tot = 0;
          for (int yy = y; yy < std::min(y + scale, o.height); ++yy)
            for (int xx = x; xx < std::min(x + scale, o.width); ++xx) { n += obst->at((size_t)xx, (size_t)yy); ++tot; }
          solid = 2 * n >= tot;
        }
        uint8_t *p = &rgb[((size_t)(h - 1 - y / scale) * w + x / scale) * 3];
        for (int k = 0; k < 3; ++k)
          p[k] = solid ? 0 : (border ? 255 : (tile_line ? (uint8_t)std::min(255, col[k] + 48) : col[k]));
      }
  }
  if (!write_png(o.out, w, h, rgb)) { std::cerr << "Could not write " << o.out << std::endl; return EXIT_FAILURE; }
  std::cout << "Wrote " << o.out << " (" << w << " x " << h << " pixels, " << scale << " cell(s) per pixel)" << std::endl;
  return EXIT_SUCCESS;
}

bool one_row(const Options &o, int nx, int ny, const lbm::Obstacles *obst) {
  int px, py;
  bool ghost;
  if (!choose(o, nx, ny, px, py, ghost)) return false;
  const part::Stats st = part::stats(nx, ny, px, py, ghost);
  std::cout << part::csv_row(nx, ny, st);
  if (obst) {  // free cells are what costs arithmetic: balance of those as well
    long long most = 0, all = 0;
    for (const part::Block &b : part::blocks(nx, ny, px, py)) {
      long long n = 0;
      for (int y = b.row0; y < b.row0 + b.nyl; ++y)
        for (int x = b.col0; x < b.col0 + b.nxl; ++x) n += !obst->at((size_t)x, (size_t)y);
      most = std::max(most, n);
      all += n;
    }
    std::cout << "," << all << "," << (most ? (double)all / st.gpus / (double)most * 100.0 : 100.0);
  }
  std::cout << std::endl;
  return true;
}

}  // namespace

int main(int argc, char **argv) {
  Options o;
  if (!parse(argc, argv, o)) {
    usage(argv[0]);
    return EXIT_FAILURE;
  }
  if (o.command == "map") return run_map(o);
  std::optional<lbm::Obstacles> obst;
  if (!o.obstacles.empty() && !o.sample) {
    obst = lbm::Obstacles::fromFile((size_t)o.width, (size_t)o.height, o.obstacles);
    if (!obst.has_value()) { std::cerr << "Could not parse obstacles file" << std::endl; return EXIT_FAILURE; }
  }
  std::cout << part::csv_header() << (obst ? ",freeCells,freeCellLoadBalance" : "") << std::endl;
  if (!o.sample) {
    if (!one_row(o, o.width, o.height, obst ? &*obst : nullptr)) {
      std::cerr << "Couldn't fit the problem on the " << o.gpus << " GPUs." << std::endl;
      return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
  }
  std::srand(o.seed);  // like the reference's sampler (rand() % range, TileMappingStats.cpp:104-112)
  for (int done = 0, tries = 0; done < o.samples && tries < 100 * o.samples; ++tries) {
    const int height = std::rand() % (o.max_h - o.min_h + 1) + o.min_h, width = std::rand() % (o.max_w - o.min_w + 1) + o.min_w;
    if (one_row(o, width, height, nullptr)) ++done;
  }
  return EXIT_SUCCESS;
}

// lbm_runner.cpp -- LbmRunner-style driver for the B200 solver.  A plain g++ translation unit: it knows
// nothing about CUDA and talks to the device only through the C ABI of include/lbm_b200.h.
//
// Keeps the sequence, flags and console lines of the reference's main/LbmRunner.cpp:
//   parse flags (:16-33) -> Params::fromFile (:38) -> Obstacles::fromFile (:44) -> Cells::initialise
//   (:67-68) -> "create engine" (:73-91) -> "Loading graph to device" (:94) -> run(0) (:98) ->
//   run(1) "Running LBM" (:102) -> run(2) (:106) -> write av_vels.dat / final_state.dat (:110-113) ->
//   "==done==", "Total compute time was", "Reynolds number:" (:124-130) -> 5 timed re-runs using the
//   device-side timer (:133-144).
//
//   lbm_runner --params FILE --obstacles FILE [-n|--num-gpus N] [-d|--debug]
//              [--iters N] [--outdir DIR] [--no-output] [--download-cells] [--reruns N]
//              [--synthetic NXxNY --solid-prob P --seed S --iters N]
//              [--exe FILE [--prepare]] [-o|--save FILE] [--grid PXxPY] [--stats]
//   (--num-ipus and --device of the reference are accepted, the latter ignored.  --exe is the reference's
//   "filename of compiled graph", main/LbmRunner.cpp:20,75-84: here a prepared-solver file written by
//   `--prepare` -- the analogue of main/LbmPoplibs.cpp:375-389, compile + serialise without running -- or by -o.)
#include <chrono>
#include <fstream>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <string>

#include "../../include/lbm_b200.h"
#include "lbm_host.hpp"
#include "lbm_partition.hpp"

namespace {

// utils::timedStep, main/include/GraphcoreUtils.hpp:130-138: description padded to 60, " took ", seconds
template <typename F>
double timedStep(const std::string &description, F f) {
  std::cerr << std::setw(60) << description;
  const auto tic = std::chrono::high_resolution_clock::now();
  f();
  const auto toc = std::chrono::high_resolution_clock::now();
  const double diff = std::chrono::duration_cast<std::chrono::duration<double>>(toc - tic).count();
  std::cerr << " took " << std::right << std::setw(12) << std::setprecision(5) << diff << "s" << std::endl;
  return diff;
}

struct Options {
  std::string params, obstacles, outdir = ".", exe, save;
  int num_gpus = 1;
  bool debug = false, no_output = false, download_cells = false, prepare = false, stats = false;
  long iters = -1;
  int reruns = 5;
  bool synthetic = false;
  long syn_nx = 0, syn_ny = 0;
  float solid_prob = 0.01f;
  unsigned long long seed = 1234;
  std::string grid;  // "" = row slabs, "PXxPY" = 2-D block partition
};

void usage(const char *exe) {
  std::cerr << exe << " - Runs a Lattice Boltzmann simulation on B200 GPUs\n"
            << "  -d, --debug            print timing/roofline details\n"
            << "      --device arg       accepted for compatibility (ignored)\n"
            << "  -n, --num-gpus arg     number of GPUs to use (alias: --num-ipus) (default: 1)\n"
            << "      --exe arg          prepared solver file: loaded instead of create + copy-to-device when it exists\n"
            << "                         (written by --prepare or -o); without it the solver is prepared on the fly\n"
            << "      --prepare          create the solver, upload the problem, write it to --exe and stop (cf. lbm_poplibs)\n"
            << "  -o, --save arg         after the run, write the solver state (resumable with --exe) to this file\n"
            << "      --params arg       filename of parameters file\n"
            << "      --obstacles arg    filename of obstacles file\n"
            << "      --iters arg        run this many timesteps instead of maxIters\n"
            << "      --outdir arg       where av_vels.dat and final_state.dat go (default: .)\n"
            << "      --no-output        skip final_state.dat (av_vels.dat is still written)\n"
            << "      --download-cells   copy the whole lattice back and derive final_state.dat on the host (reference way)\n"
            << "      --reruns arg       timed re-runs after the main run (default: 5)\n"
            << "      --grid PXxPY       2-D block partition over PX x PY GPUs instead of row slabs (-n = PX*PY)\n"
            << "      --stats            print the partition statistics (cf. tile_mapping_stats) and a map of the slabs\n"
            << "      --synthetic NXxNY  generate the synthetic channel instead of reading files\n"
            << "      --solid-prob arg   random solid fraction of the synthetic channel (default: 0.01)\n"
            << "      --seed arg         seed of the synthetic channel (default: 1234)" << std::endl;
}

// Option table with the spellings cxxopts 2.2 (the reference's parser, main/LbmRunner.cpp:14-33) accepts:
//   --name value | --name=value | -s value | grouped shorts (-dn 2: every flag but the last must be boolean) |
//   booleans as bare flags or --flag=true/false/t/f/1/0.  GNU-style -n2 is accepted on top.
struct OptSpec {
  char short_name;        // 0 = none
  const char *long_name;  // without the dashes
  const char *alias;      // second long spelling or nullptr
  bool takes_value;
  int id;
};
enum { O_DEBUG, O_DEVICE, O_NUM, O_EXE, O_PREPARE, O_SAVE, O_PARAMS, O_OBST, O_ITERS, O_OUTDIR, O_NOOUT, O_DLCELLS,
       O_RERUNS, O_GRID, O_STATS, O_SYN, O_SOLID, O_SEED };
const OptSpec kOptions[] = {
    {'d', "debug", nullptr, false, O_DEBUG},         {0, "device", nullptr, true, O_DEVICE},
    {'n', "num-gpus", "num-ipus", true, O_NUM},      {0, "exe", nullptr, true, O_EXE},
    {0, "prepare", nullptr, false, O_PREPARE},       {'o', "save", nullptr, true, O_SAVE},
    {0, "params", nullptr, true, O_PARAMS},          {0, "obstacles", nullptr, true, O_OBST},
    {0, "iters", nullptr, true, O_ITERS},            {0, "outdir", nullptr, true, O_OUTDIR},
    {0, "no-output", nullptr, false, O_NOOUT},       {0, "download-cells", nullptr, false, O_DLCELLS},
    {0, "reruns", nullptr, true, O_RERUNS},          {0, "grid", nullptr, true, O_GRID},
    {0, "stats", nullptr, false, O_STATS},           {0, "synthetic", nullptr, true, O_SYN},
    {0, "solid-prob", nullptr, true, O_SOLID},       {0, "seed", nullptr, true, O_SEED},
};

bool parse_bool(const std::string &v, bool &out) {  // cxxopts truthy/falsy patterns
  if (v == "t" || v == "T" || v == "true" || v == "True" || v == "1") { out = true; return true; }
  if (v == "f" || v == "F" || v == "false" || v == "False" || v == "0") { out = false; return true; }
  return false;
}

bool apply(Options &o, const OptSpec &spec, const std::string &v) {
  bool flag = true;
  if (!spec.takes_value && !v.empty() && !parse_bool(v, flag)) return false;
  char *end = nullptr;
  switch (spec.id) {
    case O_DEBUG: o.debug = flag; return true;
    case O_PREPARE: o.prepare = flag; return true;
    case O_NOOUT: o.no_output = flag; return true;
    case O_DLCELLS: o.download_cells = flag; return true;
    case O_STATS: o.stats = flag; return true;
    case O_DEVICE: return true;
    case O_EXE: o.exe = v; return !v.empty();
    case O_SAVE: o.save = v; return !v.empty();
    case O_PARAMS: o.params = v; return true;
    case O_OBST: o.obstacles = v; return true;
    case O_OUTDIR: o.outdir = v; return true;
    case O_GRID: o.grid = v; return true;
    case O_NUM: o.num_gpus = static_cast<int>(std::strtol(v.c_str(), &end, 10)); return end != v.c_str() && *end == 0;
    case O_ITERS: o.iters = std::strtol(v.c_str(), &end, 10); return end != v.c_str() && *end == 0;
    case O_RERUNS: o.reruns = static_cast<int>(std::strtol(v.c_str(), &end, 10)); return end != v.c_str() && *end == 0;
    case O_SOLID: o.solid_prob = std::strtof(v.c_str(), &end); return end != v.c_str();
    case O_SEED: o.seed = std::strtoull(v.c_str(), &end, 10); return end != v.c_str();
    case O_SYN:
      o.syn_nx = std::strtol(v.c_str(), &end, 10);
      if (!end || (*end != 'x' && *end != 'X')) return false;
      o.syn_ny = std::strtol(end + 1, nullptr, 10);
      o.synthetic = o.syn_nx > 0 && o.syn_ny > 0;
      return o.synthetic;
  }
  return false;
}

bool parse(int argc, char **argv, Options &o) {
  auto find_long = [](const std::string &name) -> const OptSpec * {
    for (const OptSpec &s : kOptions)
      if (name == s.long_name || (s.alias && name == s.alias)) return &s;
    return nullptr;
  };
  auto find_short = [](char c) -> const OptSpec * {
    for (const OptSpec &s : kOptions)
      if (s.short_name && s.short_name == c) return &s;
    return nullptr;
  };
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--", 0) == 0) {
      const auto eq = a.find('=');
      const OptSpec *spec = find_long(a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2));
      if (!spec) return false;
      if (eq != std::string::npos) {
        if (!apply(o, *spec, a.substr(eq + 1))) return false;
      } else if (spec->takes_value) {
        if (i + 1 >= argc || !apply(o, *spec, argv[++i])) return false;
      } else if (!apply(o, *spec, "")) return false;
    } else if (a.size() >= 2 && a[0] == '-') {
      for (size_t k = 1; k < a.size(); ++k) {
        const OptSpec *spec = find_short(a[k]);
        if (!spec) return false;
        if (!spec->takes_value) {
          if (!apply(o, *spec, "")) return false;
        } else if (k + 1 == a.size()) {  // last of the group: the value is the next argument
          if (i + 1 >= argc || !apply(o, *spec, argv[++i])) return false;
        } else {  // -n2
          if (!apply(o, *spec, a.substr(k + 1))) return false;
          break;
        }
      }
    } else return false;
  }
  if (o.prepare && o.exe.empty()) return false;
  // like the reference (:27-30) both input files are mandatory -- unless the grid is synthetic
  return o.synthetic ? o.iters > 0 : (!o.params.empty() && !o.obstacles.empty());
}

int die(const char *what) {
  std::cerr << what << ": " << lbm_last_error() << std::endl;
  return EXIT_FAILURE;
}

}  // namespace

int main(int argc, char *argv[]) {
  Options opt;
  if (!parse(argc, argv, opt)) {
    usage(argv[0]);
    return EXIT_FAILURE;
  }
  if (opt.debug) std::cout << "Capturing profile information during this run." << std::endl;

  std::optional<lbm::Params> params;
  if (opt.synthetic) {
    lbm::Params p;
    p.nx = static_cast<size_t>(opt.syn_nx);
    p.ny = static_cast<size_t>(opt.syn_ny);
    p.maxIters = static_cast<size_t>(opt.iters);
    p.reynolds_dim = 10;
    p.density = 0.1f;
    p.accel = 0.005f;
    p.omega = 1.85f;
    params = p;
  } else {
    params = lbm::Params::fromFile(opt.params);
    if (!params.has_value()) {
      std::cerr << "Could not parse parameters file. Aborting" << std::endl;
      return EXIT_FAILURE;
    }
  }
  const size_t iters = opt.iters >= 0 ? static_cast<size_t>(opt.iters) : params->maxIters;
  if (iters > params->maxIters) params->maxIters = iters;

  std::optional<lbm::Obstacles> obstacles;
  if (!opt.synthetic) {
    obstacles = lbm::Obstacles::fromFile(params->nx, params->ny, opt.obstacles);
    if (!obstacles.has_value()) {
      std::cerr << "Could not parse obstacles file" << std::endl;
      return EXIT_FAILURE;
    }
  }

  int ndev = 0;
  if (lbm_device_count(&ndev) != LBM_OK || ndev < 1) {
    std::cerr << "No CUDA device available (" << lbm_last_error() << "); this solver has no CPU path" << std::endl;
    return EXIT_FAILURE;
  }
  // --grid PXxPY | auto: a 2-D block partition instead of row slabs ("auto": the reference's rule for 2/4/8/16 devices)
  int grid_px = 0, grid_py = 0;
  if (!opt.grid.empty()) {
    if (opt.grid == "auto") {
      lbm::partition::grid_for_gpus(static_cast<int>(params->nx), static_cast<int>(params->ny), opt.num_gpus, grid_px, grid_py);
    } else {
      char *end = nullptr;
      grid_px = static_cast<int>(std::strtol(opt.grid.c_str(), &end, 10));
      if (end && (*end == 'x' || *end == 'X')) grid_py = static_cast<int>(std::strtol(end + 1, nullptr, 10));
    }
    if (grid_px < 1 || grid_py < 1) {
      std::cerr << "--grid wants PXxPY or auto (got " << opt.grid << ")" << std::endl;
      return EXIT_FAILURE;
    }
    if (opt.num_gpus != 1 && opt.num_gpus != grid_px * grid_py) {
      std::cerr << "--grid " << grid_px << "x" << grid_py << " needs -n " << grid_px * grid_py << " (got " << opt.num_gpus << ")" << std::endl;
      return EXIT_FAILURE;
    }
    opt.num_gpus = grid_px * grid_py;
  }
  std::cout << "Running on " << opt.num_gpus << " of " << ndev << " visible GPU(s)" << std::endl;

  // --exe FILE: the reference's "compiled graph" (main/LbmRunner.cpp:20,75-84).  When the file exists the solver is
  // loaded from it (problem + state, resumable) and the create / copy-to-device steps are skipped.
  bool have_exe = false;
  if (!opt.exe.empty() && !opt.prepare) {
    std::ifstream probe(opt.exe, std::ios::binary);
    have_exe = probe.good();
    if (!have_exe) std::cerr << "No prepared solver at " << opt.exe << ": preparing one on the fly" << std::endl;
  }

  double total_compute_time = 0.0;
  const bool keep_cells = !opt.synthetic;  // the synthetic channel lives on the device only
  const bool want_final_state = keep_cells && !opt.no_output;
  auto cells = lbm::Cells(keep_cells ? params->nx : 0, keep_cells ? params->ny : 0);
  if (keep_cells && !have_exe) cells.initialise(*params);

  lbm_handle *engine = nullptr;
  int rc = LBM_OK;
  int steps_before = 0;
  if (have_exe) {
    timedStep("Loading prepared solver to device", [&]() {
      rc = grid_px > 0 ? lbm_load_grid(&engine, opt.exe.c_str(), grid_px, grid_py, nullptr) : lbm_load(&engine, opt.exe.c_str(), opt.num_gpus);
    });
    if (rc != LBM_OK) return die("lbm_load");
    int nx = 0, ny = 0, max_iters = 0;
    float density = 0, accel = 0, omega = 0;
    lbm_problem(engine, &nx, &ny, &max_iters, &density, &accel, &omega);
    if (static_cast<size_t>(nx) != params->nx || static_cast<size_t>(ny) != params->ny || density != params->density ||
        accel != params->accel || omega != params->omega) {
      std::cerr << "The prepared solver in " << opt.exe << " (" << nx << "x" << ny << ") was built for a different problem than "
                << (opt.synthetic ? "--synthetic" : opt.params) << " describes. Aborting" << std::endl;
      lbm_destroy(engine);
      return EXIT_FAILURE;
    }
    lbm_steps_done(engine, &steps_before);
    if (steps_before + iters > static_cast<size_t>(max_iters)) {
      std::cerr << "The prepared solver has " << steps_before << " of its " << max_iters << " timesteps done; cannot run "
                << iters << " more" << std::endl;
      lbm_destroy(engine);
      return EXIT_FAILURE;
    }
    params->maxIters = static_cast<size_t>(max_iters);
  } else {
    timedStep("Creating solver and allocating device lattices", [&]() {
      if (grid_px > 0)  // 2-D blocks with ghost columns (the reference's partitionForIpus layout)
        rc = lbm_create_grid(&engine, static_cast<int>(params->nx), static_cast<int>(params->ny), static_cast<int>(params->maxIters),
                             params->density, params->accel, params->omega, grid_px, grid_py, nullptr);
      else
        rc = lbm_create(&engine, static_cast<int>(params->nx), static_cast<int>(params->ny),
                        static_cast<int>(params->maxIters), params->density, params->accel, params->omega, opt.num_gpus);
    });
    if (rc != LBM_OK) return die("lbm_create");
    timedStep("Running copy to device step", [&]() {
      rc = opt.synthetic ? lbm_init_synthetic(engine, opt.seed, opt.solid_prob)
                         : lbm_upload(engine, cells.data.data(), obstacles->getData());
    });
    if (rc != LBM_OK) return die("copy to device");
  }
  if (opt.prepare) {
    // the analogue of lbm_poplibs: build the solver for this problem, serialise it, do not run (main/LbmPoplibs.cpp:375-389)
    timedStep("Serialising prepared solver", [&]() { rc = lbm_save(engine, opt.exe.c_str()); });
    if (rc != LBM_OK) return die("lbm_save");
    std::cout << "==done==" << std::endl;
    lbm_destroy(engine);
    return EXIT_SUCCESS;
  }
  auto av_vels = std::vector<float>(params->maxIters, 0.0f);
  const size_t steps_total = static_cast<size_t>(steps_before) + iters;

  total_compute_time += timedStep("Running LBM", [&]() { rc = lbm_run(engine, static_cast<int>(iters)); });
  if (rc != LBM_OK) return die("lbm_run");
  double device_ms = 0.0;
  lbm_elapsed_ms(engine, &device_ms);

  // run(2).  The reference streams the whole AoS lattice back and derives u_x, u_y, u and pressure on the host
  // (main/LbmRunner.cpp:106-112); here the device evaluates those four columns (16 instead of 36 bytes per cell
  // cross PCIe) unless --download-cells asks for the reference's behaviour.
  std::vector<float> fields;
  if (want_final_state && !opt.download_cells) fields.resize(params->nx * params->ny * 4);
  timedStep("Running copy to host step", [&]() {
    rc = lbm_download(engine, (want_final_state && opt.download_cells) ? cells.data.data() : nullptr, av_vels.data());
    if (rc == LBM_OK && !fields.empty()) rc = lbm_download_fields(engine, fields.data());
  });
  if (rc != LBM_OK) return die("copy to host");
  av_vels.resize(steps_total);  // a resumed solver also returns the steps done before it was saved

  bool written = true;
  timedStep("Writing output files ", [&]() {
    written = lbm::writeAverageVelocities(opt.outdir + "/av_vels.dat", av_vels);
    if (!fields.empty())
      written = lbm::writeResultsFromFields(opt.outdir + "/final_state.dat", *params, *obstacles, fields.data()) && written;
    else if (want_final_state)
      written = lbm::writeResults(opt.outdir + "/final_state.dat", *params, *obstacles, cells) && written;
  });
  if (!written) {
    std::cerr << "Could not write the output files to " << opt.outdir << ". Aborting" << std::endl;
    lbm_destroy(engine);
    return EXIT_FAILURE;
  }
  if (!opt.save.empty()) {
    timedStep("Serialising solver state", [&]() { rc = lbm_save(engine, opt.save.c_str()); });
    if (rc != LBM_OK) return die("lbm_save");
  }

  const double cell_updates = static_cast<double>(params->nx) * static_cast<double>(params->ny) * static_cast<double>(iters);
  if (opt.stats) {
    // partition quality, the numbers of the reference's tile_mapping_stats restated for GPUs (lbm_partition.hpp), plus what
    // only the device knows: the non-obstacle cells of every block
    int ppx = 1, ppy = 1, nblocks = 0;
    lbm_partition(engine, &ppx, &ppy);
    lbm_slab_count(engine, &nblocks);
    const auto st = lbm::partition::stats(static_cast<int>(params->nx), static_cast<int>(params->ny), ppx, ppy, grid_px > 0);
    std::cout << lbm::partition::csv_header() << std::endl
              << lbm::partition::csv_row(static_cast<int>(params->nx), static_cast<int>(params->ny), st) << std::endl;
    long long most_free = 0, all_free = 0;
    for (int i = 0; i < nblocks; ++i) {
      int dev = 0, r0 = 0, nr = 0, c0 = 0, nc = 0;
      long long free_cells = 0;
      lbm_block_info(engine, i, &dev, &r0, &nr, &c0, &nc, &free_cells);
      std::cout << "Block " << i << ": GPU " << dev << ", rows [" << r0 << ", " << r0 + nr << "), cols [" << c0 << ", " << c0 + nc << "), "
                << free_cells << " free cells" << std::endl;
      most_free = std::max(most_free, free_cells);
      all_free += free_cells;
    }
    std::cout << "Free-cell load balance: " << std::fixed << std::setprecision(2)
              << (most_free ? 100.0 * static_cast<double>(all_free) / nblocks / static_cast<double>(most_free) : 100.0) << " %" << std::endl;
    std::cout.unsetf(std::ios::floatfield);
  }
  if (opt.debug) {
    const double mlups = cell_updates / (device_ms * 1e-3) / 1e6;
    std::cout << "Kernel variant:      \t" << lbm_kernel_variant(engine) << std::endl;
    int nslabs = 0;
    lbm_slab_count(engine, &nslabs);
    for (int i = 0; i < nslabs; ++i) {  // partition diagnostics (cf. main/TileMappingStats.cpp)
      int dev = 0, r0 = 0, nr = 0;
      long long free_cells = 0;
      lbm_slab_info(engine, i, &dev, &r0, &nr, &free_cells);
      std::cout << "Slab " << i << ":              \tGPU " << dev << ", rows [" << r0 << ", " << r0 + nr << "), " << free_cells
                << " free cells (" << std::fixed << std::setprecision(2)
                << 100.0 * static_cast<double>(free_cells) / (static_cast<double>(nr) * static_cast<double>(params->nx)) << " %)" << std::endl;
      std::cout.unsetf(std::ios::floatfield);
    }
    std::cout << "Device time:         \t" << std::fixed << std::setprecision(3) << device_ms << " ms" << std::endl;
    std::cout << "Throughput:          \t" << std::setprecision(1) << mlups << " MLUPS, "
              << mlups * 72.0 / 1e3 << " GB/s at 72 B per cell update" << std::endl;
    std::cout.unsetf(std::ios::floatfield);
  }

  std::cout << "==done==" << std::endl;
  std::cout << "Total compute time was \t" << std::right << std::setw(12) << std::setprecision(5) << total_compute_time
            << "s" << std::endl;
  std::cout << "Reynolds number:  \t" << std::right << std::setw(12) << std::setprecision(12) << std::scientific
            << lbm::reynoldsNumber(*params, steps_total ? av_vels[steps_total - 1] : 0.f) << std::endl;

  if (opt.reruns > 0 && !have_exe) {
    std::cout << "Now doing " << opt.reruns << " runs and averaging GPU-reported timing:" << std::endl;
    double total_ms = 0.;
    for (int run = 0; run < opt.reruns; ++run) {
      // the reference re-runs program 1 on whatever state is on the device; we restart from the same
      // initial state so every run does identical work
      rc = opt.synthetic ? lbm_init_synthetic(engine, opt.seed, opt.solid_prob)
                         : (cells.initialise(*params), lbm_upload(engine, cells.data.data(), obstacles->getData()));
      if (rc == LBM_OK) rc = lbm_run(engine, static_cast<int>(iters));
      if (rc != LBM_OK) return die("timed re-run");
      double ms = 0.;
      lbm_elapsed_ms(engine, &ms);
      total_ms += ms;
    }
    const double avg_s = total_ms / opt.reruns / 1e3;
    std::cout << "Average GPU timing for program is: " << std::fixed << std::setprecision(5) << std::setw(12) << avg_s
              << "s" << std::endl;
    std::cout << "MLUPS: " << std::setprecision(1) << cell_updates / avg_s / 1e6 << std::endl;
  }

  lbm_destroy(engine);
  return EXIT_SUCCESS;
}

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
//   (--num-ipus, --device and --exe of the reference are accepted; the latter two are ignored: there
//   is no separately compiled graph to deserialise.)
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <string>

#include "../../include/lbm_b200.h"
#include "lbm_host.hpp"

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
  std::string params, obstacles, outdir = ".";
  int num_gpus = 1;
  bool debug = false, no_output = false, download_cells = false;
  long iters = -1;
  int reruns = 5;
  bool synthetic = false;
  long syn_nx = 0, syn_ny = 0;
  float solid_prob = 0.01f;
  unsigned long long seed = 1234;
};

void usage(const char *exe) {
  std::cerr << exe << " - Runs a Lattice Boltzmann simulation on B200 GPUs\n"
            << "  -d, --debug            print timing/roofline details\n"
            << "      --device arg       accepted for compatibility (ignored)\n"
            << "  -n, --num-gpus arg     number of GPUs to use (alias: --num-ipus) (default: 1)\n"
            << "      --exe arg          accepted for compatibility (ignored)\n"
            << "      --params arg       filename of parameters file\n"
            << "      --obstacles arg    filename of obstacles file\n"
            << "      --iters arg        run this many timesteps instead of maxIters\n"
            << "      --outdir arg       where av_vels.dat and final_state.dat go (default: .)\n"
            << "      --no-output        skip final_state.dat (av_vels.dat is still written)\n"
            << "      --download-cells   copy the whole lattice back and derive final_state.dat on the host (reference way)\n"
            << "      --reruns arg       timed re-runs after the main run (default: 5)\n"
            << "      --synthetic NXxNY  generate the synthetic channel instead of reading files\n"
            << "      --solid-prob arg   random solid fraction of the synthetic channel (default: 0.01)\n"
            << "      --seed arg         seed of the synthetic channel (default: 1234)" << std::endl;
}

bool parse(int argc, char **argv, Options &o) {
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], v;
    const auto eq = a.find('=');
    bool has_inline = false;
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) {
      v = a.substr(eq + 1);
      a = a.substr(0, eq);
      has_inline = true;
    }
    auto value = [&]() -> const char * {
      if (has_inline) return v.c_str();
      if (i + 1 >= argc) return nullptr;
      return argv[++i];
    };
    const char *s = nullptr;
    if (a == "-d" || a == "--debug") o.debug = true;
    else if (a == "--no-output") o.no_output = true;
    else if (a == "--download-cells") o.download_cells = true;
    else if (a == "--params") { if (!(s = value())) return false; o.params = s; }
    else if (a == "--obstacles") { if (!(s = value())) return false; o.obstacles = s; }
    else if (a == "-n" || a == "--num-gpus" || a == "--num-ipus") { if (!(s = value())) return false; o.num_gpus = std::atoi(s); }
    else if (a == "--device" || a == "--exe") { if (!value()) return false; }
    else if (a == "--iters") { if (!(s = value())) return false; o.iters = std::atol(s); }
    else if (a == "--outdir") { if (!(s = value())) return false; o.outdir = s; }
    else if (a == "--reruns") { if (!(s = value())) return false; o.reruns = std::atoi(s); }
    else if (a == "--solid-prob") { if (!(s = value())) return false; o.solid_prob = std::strtof(s, nullptr); }
    else if (a == "--seed") { if (!(s = value())) return false; o.seed = std::strtoull(s, nullptr, 10); }
    else if (a == "--synthetic") {
      if (!(s = value())) return false;
      char *end = nullptr;
      o.syn_nx = std::strtol(s, &end, 10);
      if (!end || (*end != 'x' && *end != 'X')) return false;
      o.syn_ny = std::strtol(end + 1, nullptr, 10);
      o.synthetic = o.syn_nx > 0 && o.syn_ny > 0;
      if (!o.synthetic) return false;
    } else return false;
  }
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
  std::cout << "Running on " << opt.num_gpus << " of " << ndev << " visible GPU(s)" << std::endl;

  double total_compute_time = 0.0;
  const bool keep_cells = !opt.synthetic;  // the synthetic channel lives on the device only
  const bool want_final_state = keep_cells && !opt.no_output;
  auto cells = lbm::Cells(keep_cells ? params->nx : 0, keep_cells ? params->ny : 0);
  if (keep_cells) cells.initialise(*params);
  auto av_vels = std::vector<float>(params->maxIters, 0.0f);

  lbm_handle *engine = nullptr;
  int rc = LBM_OK;
  timedStep("Creating solver and allocating device lattices", [&]() {
    rc = lbm_create(&engine, static_cast<int>(params->nx), static_cast<int>(params->ny),
                    static_cast<int>(params->maxIters), params->density, params->accel, params->omega, opt.num_gpus);
  });
  if (rc != LBM_OK) return die("lbm_create");

  timedStep("Running copy to device step", [&]() {
    rc = opt.synthetic ? lbm_init_synthetic(engine, opt.seed, opt.solid_prob)
                       : lbm_upload(engine, cells.data.data(), obstacles->getData());
  });
  if (rc != LBM_OK) return die("copy to device");

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
  av_vels.resize(iters);

  timedStep("Writing output files ", [&]() {
    lbm::writeAverageVelocities(opt.outdir + "/av_vels.dat", av_vels);
    if (!fields.empty()) lbm::writeResultsFromFields(opt.outdir + "/final_state.dat", *params, *obstacles, fields.data());
    else if (want_final_state) lbm::writeResults(opt.outdir + "/final_state.dat", *params, *obstacles, cells);
  });

  const double cell_updates = static_cast<double>(params->nx) * static_cast<double>(params->ny) * static_cast<double>(iters);
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
            << lbm::reynoldsNumber(*params, iters ? av_vels[iters - 1] : 0.f) << std::endl;

  if (opt.reruns > 0) {
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

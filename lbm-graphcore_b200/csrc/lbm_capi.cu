// lbm_capi.cu -- the C ABI of include/lbm_b200.h on top of the kernels in lbm_kernels.cuh.
//
// Mirrors the reference's Engine-as-operator surface (main/LbmRunner.cpp:75-144): create = load the
// program for a fixed problem, upload = run(0), run = run(1), download = run(2), elapsed = readTimer.
// There is no CPU path in this file: every compute entry point needs a CUDA device.
#include "../../include/lbm_b200.h"
#include "lbm_kernels.cuh"
#include "lbm_partition.hpp"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(LBM_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

enum { KERNEL_SCALAR = lbm::SWEEP_SCALAR, KERNEL_VEC4 = lbm::SWEEP_VEC4, KERNEL_HYBRID = lbm::SWEEP_HYBRID };
// LBM_B200_DEBUG_SYNC=1: synchronise after every kernel launch and name the kernel that failed (read once: the
// check sits inside the per-timestep launch loop)
const bool g_debug_sync = getenv("LBM_B200_DEBUG_SYNC") != nullptr;
#define KERNEL_CHECK(stream, name)                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                                 \
    if (e_ == cudaSuccess && g_debug_sync) e_ = cudaStreamSynchronize(stream);                           \
    if (e_ != cudaSuccess) return fail(LBM_E_CUDA, "kernel %s failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)

constexpr size_t kFlagBytes = 256;        // exported flag area at the end of a slab's allocation (south, north, six sides)
// AoS staging buffer for upload/download, used as two halves (LBM_B200_STAGE_MB overrides the size: tuning runs)
const size_t kStageBytes = (size_t)(getenv("LBM_B200_STAGE_MB") ? std::max(8, atoi(getenv("LBM_B200_STAGE_MB"))) : 64) << 20;
constexpr int kClusterChunkSteps = 8192;  // timesteps per launch of the cluster-resident kernel
constexpr int kPersistentMaxCtas = 1024;  // upper bound on the persistent kernel's grid
constexpr size_t kPersistentPartials = 2 * (size_t)kPersistentMaxCtas * lbm::kTileH;

struct Slab {
  int device = 0;
  int row0 = 0;  // first owned global row
  int nyl = 0;   // owned rows
  int col0 = 0;  // first owned global column (0 unless the grid is cut into 2-D blocks)
  int nxl = 0;   // owned columns (nx for a row slab)
  int ntx_sweep = 0;  // tile columns that hold owned cells = ceil(nxl / 128)
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  // upload/download pipeline: the staging buffer is used as two halves; PCIe copies run on copy_stream while the
  // transposition kernel of the other half runs on `stream`
  cudaStream_t copy_stream = nullptr;
  int stage_chunk = 0;  // chunks enqueued since the last drain (which half of the staging buffer is next)
  cudaEvent_t filled[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
  // one exported allocation: [lattice 0][lattice 1][flags]
  char *base = nullptr;
  float *buf[2] = {nullptr, nullptr};
  int *flags = nullptr;  // [0] from_south, [1] from_north
  // private allocations
  uint32_t *mask = nullptr;
  double *partials = nullptr;      // [kWindowSteps][grid] CTA partials of the last few timesteps (+ persistent kernel's warp partials)
  double *row_partials = nullptr;  // [kWindowSteps][tiles_y]
  double *tot_u = nullptr;
  int window_first = 0;            // first timestep whose partials wait in the window
  int window_count = 0;            // how many
  unsigned *ticket = nullptr;      // [1..2] edge tickets (multi-slab)
  int *step = nullptr;         // [1] error flag set by a kernel that gave up waiting ([0] unused)
  unsigned long long *barrier = nullptr;  // persistent kernel: grid-barrier arrivals since the last upload
  unsigned long long barrier_arrivals = 0;  // host copy of what *barrier holds when the queued launches have finished
  int persistent_grid = 0;     // > 0: the whole time loop runs as one cooperative launch of this many CTAs
  // cluster-resident kernel (lattices in distributed shared memory): the smallest shipped inputs
  int cluster_ctas = 0;        // > 0: the whole time loop runs as one launch of this many CTAs ...
  int cluster_size = 0;        // ... in clusters of this many
  int cluster_cells = 0, cluster_cpt = 0;  // cells per CTA / per thread
  bool cluster_refused = false;  // a launch failed: do not try again
  uint2 *cluster_halo = nullptr;   // tagged packets of the edge rows that cross a cluster boundary
  size_t cluster_halo_bytes = 0;
  double *cluster_tot = nullptr;   // [kClusterChunkSteps][cluster_ctas] CTA partials of one launch
  unsigned long long *free_dev = nullptr;
  float *stage = nullptr;
  // neighbours (own pointers when there is a single slab)
  float *south_buf[2] = {nullptr, nullptr};
  float *north_buf[2] = {nullptr, nullptr};
  int *flag_to_south = nullptr;
  int *flag_to_north = nullptr;
  int south_ghost_row = 0;
  // 2-D blocks: the six further neighbours (lbm::SIDE_W .. SIDE_NE), own pointers where the partition wraps onto itself
  float *side_buf[6][2] = {};
  int *flag_to_side[6] = {};
  // launch configuration
  int kernel = 0;  // KERNEL_SCALAR / KERNEL_VEC4 / KERNEL_HYBRID
  int tiles_y = 1; // tile rows this slab sweeps = ceil((nyl + 2) / 8)
  int tot_ty_lo = 0, tot_ty_hi = -1;    // tile rows whose cells count towards tot_u (-1: all); free cells: storage rows ...
  int count_lo = 0, count_hi = 1 << 30; // ... [count_lo, count_hi).  Only the bands of lbm_solve_host restrict these.
  int grid = 1;    // one CTA per tile
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long free_cells = 0;
};

// replay unit of the time loop on one stream: kGraphSteps timesteps of every slab that runs on it
struct StreamGraph {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};
  long long nodes = 0;  // kernel launches one replay stands for
};
constexpr int kGraphSteps = 2 * lbm::kWindowSteps;  // even: a replay leaves the lattice parity unchanged
constexpr long long kGraphMaxCells = 8ll << 20;     // per slab: above this a step takes >> a launch, nothing to gain

}  // namespace

struct lbm_handle {
  int nx = 0, ny = 0, max_iters = 0;
  float density = 0, accel = 0, omega = 0;
  int ntx = 0;        // tiles per storage row
  int tiles_y = 0;    // tile rows allocated per slab (same on every slab so IPC offsets agree)
  int rows_alloc = 0; // storage rows per slab incl. the two ghost rows
  size_t lattice_floats = 0;
  size_t lattice_bytes = 0;
  size_t flag_bytes = kFlagBytes;
  std::vector<Slab> slabs;
  bool multi = false;      // more than one slab in the ring (or a 2-D block partition): flags and halo pushes are on
  bool xghost = false;     // 2-D block partition: blocks keep ghost columns (lbm_create_grid)
  int px = 1, py = 1;      // blocks per row / per column of the partition (1 x slabs for row slabs)
  bool rank_mode = false;  // one process per GPU
  // a rank of a block partition passes / receives only its block: host rows are nxl cells wide and start at its first column
  bool rank_block() const { return rank_mode && xghost; }
  int rank = 0, world = 1;
  bool connected = false;
  std::vector<void *> ipc_opened;
  int accel_row_global = 0;  // the driven row: ny - 2 (main/LastChance.cpp:198); a band of lbm_solve_host holds it elsewhere or not at all
  int cur = 0;  // which lattice holds the current state
  int steps_done = 0;
  bool uploaded = false;
  bool first_run_pending = false;  // the one-off guarded accelerate_flow has not been applied yet
  bool halo_pending = false;       // the ghost rows of the current lattice (and the first epoch) have not been published yet
  double last_ms = 0.0;
  long long launches = 0;
  // time loop as a CUDA graph (one per distinct stream and per lattice parity), built on first use and replayed
  bool no_graphs = false;          // band solvers of lbm_solve_host: their driven row changes from band to band
  bool graph_capture = false;      // launches are being captured: kernels take the step number from device memory
  int graph_state[2] = {0, 0};     // per parity: 0 not built, 1 ready, -1 unavailable
  std::vector<StreamGraph> graphs;
};

namespace {

size_t blob_bytes() { return sizeof(cudaIpcMemHandle_t) + 4 * sizeof(int); }

// the instantiations of sweep_cluster_kernel: (nx, cells per CTA, cells per thread)
const void *cluster_kernel(int nx, int cells, int cpt) {
#define LBM_CK(NX, CELLS, CPT) \
  if (nx == NX && cells == CELLS && cpt == CPT) return (const void *)lbm::sweep_cluster_kernel<NX, CELLS, CPT>;
  LBM_CK(128, 1024, 4) LBM_CK(128, 1024, 2) LBM_CK(128, 512, 4) LBM_CK(128, 512, 2) LBM_CK(128, 512, 1) LBM_CK(128, 256, 1)
  LBM_CK(256, 1024, 4) LBM_CK(256, 1024, 2) LBM_CK(256, 512, 4) LBM_CK(256, 512, 2) LBM_CK(256, 512, 1)
  LBM_CK(512, 1024, 2) LBM_CK(512, 2048, 4)
#undef LBM_CK
  return nullptr;
}
void cluster_launch_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int nx, int cells, int cpt, int ctas, int cs,
                           cudaStream_t stream) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)ctas, 1, 1);
  cfg.blockDim = dim3((unsigned)lbm::cluster_threads(cells, cpt), 1, 1);
  cfg.dynamicSmemBytes = lbm::cluster_smem_bytes(nx, cells);
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// Choose the sweep kernel and its launch shape for one slab: one CTA of 8 warps per 128 x 8 tile.
int configure_launch(lbm_handle *h, Slab &s) {
  s.kernel = (s.nxl % lbm::kTileW == 0) ? KERNEL_VEC4 : (s.nxl > lbm::kTileW ? KERNEL_HYBRID : KERNEL_SCALAR);
  // development knob (parity tests of the scalar path on vectorisable sizes)
  if (const char *e = getenv("LBM_B200_KERNEL"))
    if (!strcmp(e, "scalar")) s.kernel = KERNEL_SCALAR;
  s.tiles_y = (s.nyl + 2 + lbm::kTileH - 1) / lbm::kTileH;
  s.grid = s.ntx_sweep * s.tiles_y;
  // sweep_persistent_kernel runs the whole time loop as ONE cooperative launch with a grid barrier per step.  It was the
  // default for lattices that fit in L2 until the per-step loop was replayed from a CUDA graph, which is as fast or faster at
  // every size (profiles/r02j_mid_grids.txt: 768^2 7.2 vs 8.6 us per step, 1024^2 12.8 vs 13.3, 2048^2 48.6 vs 51.6) and needs
  // no co-residency.  LBM_B200_PERSISTENT=1 still selects it (tests, comparisons); =0 also switches the cluster kernel off.
  s.persistent_grid = 0;
  const int want = getenv("LBM_B200_PERSISTENT") ? atoi(getenv("LBM_B200_PERSISTENT")) : -1;
  if (!h->multi && want == 1) {
    CUDA_TRY(cudaSetDevice(s.device));
    int sms = 0, coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device));
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s.device));
    if (coop && sms > 0) {
      // LBM_PERSIST_MINB (2) CTAs per SM at most -- the barrier costs one atomic per CTA; tile rows are dealt over
      // all their warps
      const int units = s.grid * lbm::kTileH;
      int ctas = std::min(std::min(LBM_PERSIST_MINB * sms, kPersistentMaxCtas), units);
      if (const char *e = getenv("LBM_B200_PERSIST_CTAS")) ctas = std::max(1, std::min(atoi(e), ctas));
      s.persistent_grid = ctas;
    }
  }
  // The smallest inputs (nx = 128, 256 or 512, up to ~300 K cells) keep both lattices in the distributed shared memory of
  // thread-block clusters (sweep_cluster_kernel): CTAs of 512 cells (1024 or, for nx = 512, 2048 if that many CTAs cannot be
  // co-resident or do not hold two full rows; 256 on request) -- one cell per thread (two at 1024, four at 2048), the largest cluster size <= 16 that divides the CTA count and fits.  LBM_B200_CLUSTER=0 disables; LBM_B200_CLUSTER_CELLS /
  // _CPT / _SIZE pin the shape (tuning runs).
  s.cluster_ctas = 0;
  const int want_cluster = getenv("LBM_B200_CLUSTER") ? atoi(getenv("LBM_B200_CLUSTER")) : 1;
  if (!h->multi && want_cluster != 0 && want == -1 && !getenv("LBM_B200_KERNEL") && !s.cluster_refused &&
      (h->nx == 128 || h->nx == 256 || h->nx == 512)) {
    CUDA_TRY(cudaSetDevice(s.device));
    const int env_cells = getenv("LBM_B200_CLUSTER_CELLS") ? atoi(getenv("LBM_B200_CLUSTER_CELLS")) : 0;
    const int env_cpt = getenv("LBM_B200_CLUSTER_CPT") ? atoi(getenv("LBM_B200_CLUSTER_CPT")) : 0;
    const int env_cs = getenv("LBM_B200_CLUSTER_SIZE") ? atoi(getenv("LBM_B200_CLUSTER_SIZE")) : 0;
    // smallest CTAs first: the update is issue-bound per SM, so the more SMs share a lattice the shorter a step
    // (profiles/r02b_small_grids_cluster_v3.txt: 128^2 1.31 / 0.86 us per step with 1024 / 512 cells per CTA)
    for (int cells : {512, 1024, 2048, 256}) {  // (256-cell CTAs measured no faster than 512: twice the SMs for the same step time)
      if (s.cluster_ctas > 0) break;
      if (env_cells && env_cells != cells) continue;
      const int rows = cells / h->nx, cpt = env_cpt ? env_cpt : (cells == 2048 ? 4 : cells == 1024 ? 2 : 1);
      if (rows < 2 || s.nyl % rows != 0) continue;
      const int G = s.nyl / rows;
      const void *fn = cluster_kernel(h->nx, cells, cpt);
      if (!fn) continue;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lbm::cluster_smem_bytes(h->nx, cells)) != cudaSuccess ||
          cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      for (int cs = lbm::kClusterMaxSize; cs >= 1; --cs) {  // the largest cluster that divides the grid and fits
        if (G % cs != 0 || (env_cs && env_cs != cs)) continue;
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr[1];
        cluster_launch_config(cfg, attr, h->nx, cells, cpt, G, cs, s.stream);
        int active = 0;
        if (cudaOccupancyMaxActiveClusters(&active, fn, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
        if (active * cs < G) continue;  // the clusters wait for each other: all of them must be resident
        s.cluster_ctas = G;
        s.cluster_size = cs;
        s.cluster_cells = cells;
        s.cluster_cpt = cpt;
        break;
      }
    }
    cudaGetLastError();
    if (s.cluster_ctas > 0) {
      // (re)allocate for this shape: halo packets per CTA, one partial per CTA and step of a launch
      cudaFree(s.cluster_halo); cudaFree(s.cluster_tot);
      s.cluster_halo = nullptr; s.cluster_tot = nullptr;
      s.cluster_halo_bytes = (size_t)2 * s.cluster_ctas * 2 * 3 * h->nx * sizeof(uint2);
      CUDA_TRY(cudaMalloc((void **)&s.cluster_halo, s.cluster_halo_bytes));
      CUDA_TRY(cudaMemset(s.cluster_halo, 0, s.cluster_halo_bytes));
      CUDA_TRY(cudaMalloc((void **)&s.cluster_tot, (size_t)kClusterChunkSteps * s.cluster_ctas * sizeof(double)));
      s.persistent_grid = 0;
    }
  }
  return LBM_OK;
}

int alloc_slab(lbm_handle *h, Slab &s) {
  CUDA_TRY(cudaSetDevice(s.device));
  const size_t total = 2 * h->lattice_bytes + h->flag_bytes;
  cudaError_t e = cudaMalloc((void **)&s.base, total);
  if (e != cudaSuccess)
    return fail(LBM_E_NOMEM, "cudaMalloc of %zu bytes for a %d-row slab failed: %s", total, s.nyl,
                cudaGetErrorString(e));
  s.buf[0] = reinterpret_cast<float *>(s.base);
  s.buf[1] = reinterpret_cast<float *>(s.base + h->lattice_bytes);
  s.flags = reinterpret_cast<int *>(s.base + 2 * h->lattice_bytes);
  CUDA_TRY(cudaMemset(s.flags, 0, h->flag_bytes));
  const size_t mask_bytes = (size_t)h->tiles_y * h->ntx * lbm::kTileMaskWords * sizeof(uint32_t);
  CUDA_TRY(cudaMalloc((void **)&s.mask, mask_bytes));
  CUDA_TRY(cudaMemset(s.mask, 0, mask_bytes));
  // per-step kernels: kWindowSteps slots of one partial per CTA (= tile); persistent kernel: two sets of one partial per warp
  CUDA_TRY(cudaMalloc((void **)&s.partials,
                      std::max(kPersistentPartials, (size_t)lbm::kWindowSteps * h->tiles_y * h->ntx) * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&s.barrier, sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(s.barrier, 0, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc((void **)&s.row_partials, (size_t)lbm::kWindowSteps * h->tiles_y * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&s.tot_u, (size_t)std::max(h->max_iters, 1) * sizeof(double)));
  CUDA_TRY(cudaMemset(s.tot_u, 0, (size_t)std::max(h->max_iters, 1) * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&s.ticket, 16 * sizeof(unsigned)));
  CUDA_TRY(cudaMemset(s.ticket, 0, 16 * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc((void **)&s.step, 2 * sizeof(int)));
  CUDA_TRY(cudaMemset(s.step, 0, 2 * sizeof(int)));
  CUDA_TRY(cudaMalloc((void **)&s.free_dev, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc((void **)&s.stage, kStageBytes));
  CUDA_TRY(cudaEventCreate(&s.ev0));
  CUDA_TRY(cudaEventCreate(&s.ev1));
  CUDA_TRY(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    CUDA_TRY(cudaEventCreateWithFlags(&s.filled[b], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.drained[b], cudaEventDisableTiming));
  }
  return LBM_OK;
}

void free_slab(Slab &s) {
  cudaSetDevice(s.device);
  if (s.ev0) cudaEventDestroy(s.ev0);
  if (s.ev1) cudaEventDestroy(s.ev1);
  for (int b = 0; b < 2; ++b) {
    if (s.filled[b]) cudaEventDestroy(s.filled[b]);
    if (s.drained[b]) cudaEventDestroy(s.drained[b]);
  }
  if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
  cudaFree(s.cluster_tot);
  cudaFree(s.cluster_halo);
  cudaFree(s.stage);
  cudaFree(s.free_dev);
  cudaFree(s.step);
  cudaFree(s.ticket);
  cudaFree(s.tot_u);
  cudaFree(s.barrier);
  cudaFree(s.row_partials);
  cudaFree(s.partials);
  cudaFree(s.mask);
  cudaFree(s.base);
  if (s.owns_stream && s.stream) cudaStreamDestroy(s.stream);
}

int check_problem(int nx, int ny, int max_iters, int slabs) {
  if (nx < 1 || ny < 1) return fail(LBM_E_ARG, "grid must be at least 1x1 (got %dx%d)", nx, ny);
  if (max_iters < 0) return fail(LBM_E_ARG, "max_iters must be >= 0 (got %d)", max_iters);
  if (slabs < 1) return fail(LBM_E_ARG, "need at least one GPU/slab (got %d)", slabs);
  if (ny < slabs) return fail(LBM_E_ARG, "cannot cut %d rows into %d slabs", ny, slabs);
  const long long tiles = (long long)((nx + lbm::kTileW - 1) / lbm::kTileW) * ((ny / slabs + 3 + lbm::kTileH) / lbm::kTileH);
  if (tiles >= (1ll << 31)) return fail(LBM_E_ARG, "slab of %dx%d cells exceeds 2^31 tiles", nx, ny / slabs + 1);
  return LBM_OK;
}

// `ring` row slabs (px == 1), or px x ring blocks with ghost columns (xghost)
void set_geometry(lbm_handle *h, int nx, int ny, int max_iters, float density, float accel, float omega,
                  int ring, int px = 1, bool xghost = false) {
  h->nx = nx; h->ny = ny; h->max_iters = max_iters;
  h->density = density; h->accel = accel; h->omega = omega;
  h->px = px; h->py = ring; h->xghost = xghost;
  const int widest = (nx + px - 1) / px;
  h->ntx = (widest + lbm::kTileW - 1) / lbm::kTileW + (xghost ? 1 : 0);  // + the halo tile column; same on every block
  h->rows_alloc = (ny + ring - 1) / ring + 2;  // the tallest slab + two ghost rows; same on every slab
  h->tiles_y = (h->rows_alloc + lbm::kTileH - 1) / lbm::kTileH;
  h->lattice_floats = (size_t)h->tiles_y * h->ntx * lbm::kTileFloats;
  h->lattice_bytes = h->lattice_floats * sizeof(float);
  h->multi = ring > 1 || xghost;
  h->accel_row_global = ny - 2;
}

lbm::SweepArgs sweep_args(const lbm_handle *h, const Slab &s) {
  lbm::SweepArgs a{};
  const int c = h->cur;
  a.src = s.buf[c];
  a.dst = s.buf[c ^ 1];
  a.north_dst = s.north_buf[c ^ 1];
  a.south_dst = s.south_buf[c ^ 1];
  a.mask = s.mask;
  a.lattice_floats = h->lattice_floats;
  a.ntx = h->ntx;
  a.ntx_sweep = s.ntx_sweep;
  a.tiles_y = s.tiles_y;
  a.nx = s.nxl;
  a.nyl = s.nyl;
  a.south_ghost_row = s.south_ghost_row;
  const int arow = h->accel_row_global;  // main/LastChance.cpp:198 (never matches when ny < 2, like the reference)
  a.accel_row = (arow >= s.row0 && arow < s.row0 + s.nyl) ? (arow - s.row0 + 1) : -1;
  a.omega = h->omega;
  a.one_minus_omega = 1 - h->omega;                 // main/LastChance.cpp:387
  a.w1 = h->density * h->accel / 9.f;               // main/LastChance.cpp:158
  a.w2 = h->density * h->accel / 36.f;              // :159
  a.t = h->steps_done;  // launch_sweep overrides it for every step of a run
  a.t_base = h->graph_capture ? s.step : nullptr;
  a.partials = s.partials;
  a.tot_u = s.tot_u;
  a.edge_ticket = s.ticket + 1;
  a.flag_from_south = s.flags + 0;
  a.flag_from_north = s.flags + 1;
  a.flag_to_south = s.flag_to_south;
  a.flag_to_north = s.flag_to_north;
  a.error = s.step + 1;
  for (int i = 0; i < 6; ++i) {
    a.side_dst[i] = s.side_buf[i][c ^ 1];
    a.flag_to_side[i] = s.flag_to_side[i];
  }
  a.flag_from_side = s.flags + 2;
  a.sdiag_ghost_row = s.south_ghost_row;
  return a;
}

template <typename K>
cudaError_t launch_pdl(K kernel, dim3 grid, cudaStream_t stream, bool pdl, const lbm::SweepArgs &a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(32, lbm::kTileH, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

// One timestep of one slab.  `overlap` = this launch directly follows another sweep of the same slab in the
// stream, so it may be scheduled while that one drains (programmatic dependent launch; LBM_B200_PDL=0 disables).
// Finish the sums of the timesteps whose CTA partials wait in the slab's window: tot_u[window_first ...].
int flush_reductions(lbm_handle *h, Slab &s) {
  if (s.window_count == 0) return LBM_OK;
  lbm::reduce_rows_kernel<<<dim3((unsigned)s.tiles_y, (unsigned)s.window_count), 256, 0, s.stream>>>(s.partials, s.row_partials,
                                                                                                 s.ntx_sweep, s.tiles_y);
  lbm::reduce_total_kernel<<<(unsigned)s.window_count, 256, 0, s.stream>>>(s.row_partials, s.tot_u, s.tiles_y, s.window_first,
                                                                           h->graph_capture ? s.step : nullptr, s.tot_ty_lo,
                                                                           s.tot_ty_hi < 0 ? s.tiles_y : s.tot_ty_hi);
  h->launches += 2;
  KERNEL_CHECK(s.stream, "reduce_rows/reduce_total");
  s.window_first += s.window_count;
  s.window_count = 0;
  return LBM_OK;
}

int launch_sweep(lbm_handle *h, Slab &s, int t, bool overlap) {
  static const bool pdl_enabled = !(getenv("LBM_B200_PDL") && atoi(getenv("LBM_B200_PDL")) == 0);
  if (s.window_count == lbm::kWindowSteps) {
    int rc = flush_reductions(h, s);
    if (rc) return rc;
    overlap = false;  // the reduce kernels now sit between the two sweeps
  }
  if (s.window_count == 0) s.window_first = t;
  lbm::SweepArgs a = sweep_args(h, s);
  a.t = t;
  if (h->xghost) {
    // whose completion this launch announces (ghost columns / corners, see sweep_kernel): the slab launched just before it in
    // this stream -- the previous slab on the stream at this step, or the stream's last slab at the previous step.  The order
    // is the fixed one of enqueue_steps, so this is static (and the same inside a captured graph, with relative steps).
    const Slab *pred = nullptr, *last = nullptr;
    for (const Slab &o : h->slabs) {
      if (o.stream != s.stream) continue;
      if (&o < &s) pred = &o;
      last = &o;
    }
    const Slab *who = pred ? pred : last;
    for (int i = 0; i < 6; ++i) a.pub_flag[i] = who->flag_to_side[i];
    a.pub_epoch = pred ? t + 2 : t + 1;  // (pred ran step t: its epoch t + 2; the last slab ran step t - 1: t + 1)
  }
  a.partials = s.partials + (size_t)s.window_count * s.grid;
  ++s.window_count;
  const bool pdl = overlap && pdl_enabled;
  cudaError_t e;
  const int multi = h->xghost ? 2 : (h->multi ? 1 : 0);
  // blocks: a 2-D grid (tile column, tile row); slabs: tiles in one dimension
  const dim3 grid1((unsigned)s.grid, 1, 1), grid2((unsigned)s.ntx_sweep, (unsigned)s.tiles_y, 1);
#define LBM_SWEEP(MODE)                                                                              \
  (multi == 2   ? launch_pdl(lbm::sweep_kernel<MODE, 2>, grid2, s.stream, pdl, a)                     \
   : multi == 1 ? launch_pdl(lbm::sweep_kernel<MODE, 1>, grid1, s.stream, pdl, a)                     \
                : launch_pdl(lbm::sweep_kernel<MODE, 0>, grid1, s.stream, pdl, a))
  if (s.kernel == KERNEL_VEC4) e = LBM_SWEEP(lbm::SWEEP_VEC4);
  else if (s.kernel == KERNEL_HYBRID) e = LBM_SWEEP(lbm::SWEEP_HYBRID);
  else e = LBM_SWEEP(lbm::SWEEP_SCALAR);
#undef LBM_SWEEP
  if (e != cudaSuccess) return fail(LBM_E_CUDA, "sweep launch failed: %s", cudaGetErrorString(e));
  ++h->launches;
  KERNEL_CHECK(s.stream, "sweep");
  return LBM_OK;
}

// The whole time loop of a single small slab as one cooperative launch.
int launch_persistent(lbm_handle *h, Slab &s, int iters) {
  lbm::SweepArgs a = sweep_args(h, s);
  float *buf_in = s.buf[h->cur], *buf_out = s.buf[h->cur ^ 1];
  int t0 = h->steps_done, steps = iters;
  unsigned long long *bar = s.barrier, bar0 = s.barrier_arrivals;
  void *args[] = {&a, &buf_in, &buf_out, &t0, &steps, &bar, &bar0};
  const void *fn = s.kernel == KERNEL_VEC4     ? (const void *)lbm::sweep_persistent_kernel<lbm::SWEEP_VEC4>
                   : s.kernel == KERNEL_HYBRID ? (const void *)lbm::sweep_persistent_kernel<lbm::SWEEP_HYBRID>
                                               : (const void *)lbm::sweep_persistent_kernel<lbm::SWEEP_SCALAR>;
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)s.persistent_grid), dim3(32, lbm::kTileH, 1), args, 0, s.stream);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
    // the CTAs cannot all be resident here (a shared or partitioned GPU): use one launch per timestep instead
    cudaGetLastError();
    s.persistent_grid = 0;
    return LBM_OK;
  }
  if (e != cudaSuccess) return fail(LBM_E_CUDA, "cooperative launch of the persistent sweep failed: %s", cudaGetErrorString(e));
  s.barrier_arrivals += (unsigned long long)iters * (unsigned long long)s.persistent_grid;
  ++h->launches;
  KERNEL_CHECK(s.stream, "sweep_persistent");
  return LBM_OK;
}

// The time loop of a single tiny slab as launches of thread-block clusters that keep the lattices in shared memory
// (kClusterChunkSteps timesteps per launch: the per-step CTA partials of one launch wait in s.cluster_tot).
int launch_cluster(lbm_handle *h, Slab &s, int iters) {
  lbm::ClusterArgs a{};
  a.lattice = s.buf[h->cur];
  a.mask = s.mask;
  a.halo = s.cluster_halo;
  a.cta_tot = s.cluster_tot;
  a.error = s.step + 1;
  a.ntx = h->ntx;
  a.ny = h->ny;
  a.accel_row = h->accel_row_global;
  a.omega = h->omega;
  a.one_minus_omega = 1 - h->omega;
  a.w1 = h->density * h->accel / 9.f;
  a.w2 = h->density * h->accel / 36.f;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_launch_config(cfg, attr, h->nx, s.cluster_cells, s.cluster_cpt, s.cluster_ctas, s.cluster_size, s.stream);
  const void *fn = cluster_kernel(h->nx, s.cluster_cells, s.cluster_cpt);
  for (int done = 0; done < iters; done += kClusterChunkSteps) {
    a.t0 = h->steps_done + done;
    a.steps = std::min(kClusterChunkSteps, iters - done);
    void *args[] = {&a};
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) {
      cudaGetLastError();
      if (done > 0) return fail(LBM_E_CUDA, "cluster launch failed mid-run: %s", cudaGetErrorString(e));
      s.cluster_ctas = 0;  // not launchable here after all: fall back to the other paths for good
      s.cluster_refused = true;
      return LBM_OK;
    }
    lbm::reduce_ctas_kernel<<<(a.steps + 255) / 256, 256, 0, s.stream>>>(s.cluster_tot, s.tot_u, s.cluster_ctas, a.t0, a.steps);
    h->launches += 2;
    KERNEL_CHECK(s.stream, "sweep_cluster/reduce_ctas");
  }
  return LBM_OK;
}

// ---- staged host <-> device pipelines -------------------------------------------------------------
// Rows travel in chunks through the two halves of the slab's staging buffer.  `chunk` counts chunks since the last
// drain so that consecutive pipelines (cells, then obstacles) keep alternating halves correctly.
const size_t kHalfStage = kStageBytes / 2;

// host -> staging half (copy_stream) -> `consume(stage_half, first_row, rows)` launches a kernel on s.stream
// (host rows are `host_pitch` bytes apart: the rows of a 2-D block are pieces of the caller's full-width rows)
cudaError_t copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows, cudaMemcpyKind kind,
                      cudaStream_t stream) {
  if (dpitch == row_bytes && spitch == row_bytes) return cudaMemcpyAsync(dst, src, (size_t)rows * row_bytes, kind, stream);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, (size_t)rows, kind, stream);
}

template <typename Consume>
int stage_in(Slab &s, const char *host, size_t host_pitch, size_t row_bytes, int total_rows, int &chunk, Consume consume) {
  if (row_bytes > kHalfStage) return fail(LBM_E_ARG, "a single row of %zu bytes does not fit the staging buffer", row_bytes);
  const int rows_per_chunk = (int)std::max<size_t>(1, kHalfStage / row_bytes);
  for (int r = 0; r < total_rows; r += rows_per_chunk, ++chunk) {
    const int rows = std::min(rows_per_chunk, total_rows - r), b = chunk & 1;
    char *half = reinterpret_cast<char *>(s.stage) + (size_t)b * kHalfStage;
    if (chunk >= 2) CUDA_TRY(cudaStreamWaitEvent(s.copy_stream, s.drained[b], 0));  // its previous user has read it
    CUDA_TRY(copy_rows(half, row_bytes, host + (size_t)r * host_pitch, host_pitch, row_bytes, rows, cudaMemcpyHostToDevice, s.copy_stream));
    CUDA_TRY(cudaEventRecord(s.filled[b], s.copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(s.stream, s.filled[b], 0));
    int rc = consume(half, r, rows);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s.drained[b], s.stream));
  }
  return LBM_OK;
}
// `produce(stage_half, first_row, rows)` launches a kernel on s.stream -> staging half -> host (copy_stream)
template <typename Produce>
int stage_out(Slab &s, char *host, size_t host_pitch, size_t row_bytes, int total_rows, int &chunk, Produce produce) {
  if (row_bytes > kHalfStage) return fail(LBM_E_ARG, "a single row of %zu bytes does not fit the staging buffer", row_bytes);
  const int rows_per_chunk = (int)std::max<size_t>(1, kHalfStage / row_bytes);
  for (int r = 0; r < total_rows; r += rows_per_chunk, ++chunk) {
    const int rows = std::min(rows_per_chunk, total_rows - r), b = chunk & 1;
    char *half = reinterpret_cast<char *>(s.stage) + (size_t)b * kHalfStage;
    if (chunk >= 2) CUDA_TRY(cudaStreamWaitEvent(s.stream, s.drained[b], 0));  // its previous contents are on the host
    int rc = produce(half, r, rows);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s.filled[b], s.stream));
    CUDA_TRY(cudaStreamWaitEvent(s.copy_stream, s.filled[b], 0));
    CUDA_TRY(copy_rows(host + (size_t)r * host_pitch, host_pitch, half, row_bytes, row_bytes, rows, cudaMemcpyDeviceToHost, s.copy_stream));
    CUDA_TRY(cudaEventRecord(s.drained[b], s.copy_stream));
  }
  return LBM_OK;
}
int drain_stages(Slab &s) {
  CUDA_TRY(cudaStreamSynchronize(s.copy_stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  return LBM_OK;
}

// wire the ring inside one process: slab i's south neighbour is i-1, north is i+1 (periodic)
int wire_local_ring(lbm_handle *h) {
  const int n = (int)h->slabs.size();
  for (int i = 0; i < n; ++i) {
    Slab &s = h->slabs[i];
    Slab &so = h->slabs[(i + n - 1) % n];
    Slab &no = h->slabs[(i + 1) % n];
    for (int b = 0; b < 2; ++b) {
      s.south_buf[b] = so.buf[b];
      s.north_buf[b] = no.buf[b];
    }
    s.flag_to_south = so.flags + 1;  // its flag_from_north
    s.flag_to_north = no.flags + 0;  // its flag_from_south
    s.south_ghost_row = so.nyl + 1;
  }
  // peer access between distinct devices of neighbouring slabs
  for (int i = 0; i < n; ++i) {
    for (int d : {h->slabs[(i + n - 1) % n].device, h->slabs[(i + 1) % n].device}) {
      const int me = h->slabs[i].device;
      if (d == me) continue;
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, me, d));
      if (!can) return fail(LBM_E_PEER, "GPU %d cannot access GPU %d as a peer", me, d);
      CUDA_TRY(cudaSetDevice(me));
      cudaError_t e = cudaDeviceEnablePeerAccess(d, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess)
        return fail(LBM_E_PEER, "cudaDeviceEnablePeerAccess(%d -> %d): %s", me, d, cudaGetErrorString(e));
    }
  }
  h->connected = true;
  return LBM_OK;
}

int reset_counters(lbm_handle *h) {
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaMemsetAsync(s.flags, 0, h->flag_bytes, s.stream));
    CUDA_TRY(cudaMemsetAsync(s.ticket, 0, 16 * sizeof(unsigned), s.stream));
    s.window_first = 0;
    s.window_count = 0;
    CUDA_TRY(cudaMemsetAsync(s.step, 0, 2 * sizeof(int), s.stream));
    CUDA_TRY(cudaMemsetAsync(s.barrier, 0, sizeof(unsigned long long), s.stream));
    s.barrier_arrivals = 0;
    CUDA_TRY(cudaMemsetAsync(s.tot_u, 0, (size_t)std::max(h->max_iters, 1) * sizeof(double), s.stream));
    CUDA_TRY(cudaMemsetAsync(s.free_dev, 0, sizeof(unsigned long long), s.stream));
    if (s.cluster_halo) CUDA_TRY(cudaMemsetAsync(s.cluster_halo, 0, s.cluster_halo_bytes, s.stream));  // packet tags restart at 0
  }
  h->cur = 0;
  h->steps_done = 0;
  return LBM_OK;
}

int finish_upload(lbm_handle *h) {
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    unsigned long long f = 0;
    CUDA_TRY(cudaMemcpyAsync(&f, s.free_dev, sizeof f, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    s.free_cells = (long long)f;
  }
  h->uploaded = true;
  h->first_run_pending = true;
  h->halo_pending = true;
  return LBM_OK;
}

int create_common(lbm_handle **out, lbm_handle *h) {
  for (Slab &s : h->slabs) {
    int rc = alloc_slab(h, s);
    if (rc) { lbm_destroy(h); return rc; }
    rc = configure_launch(h, s);
    if (rc) { lbm_destroy(h); return rc; }
  }
  *out = h;
  return LBM_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

const char *lbm_last_error(void) { return g_last_error.c_str(); }
int lbm_abi_version(void) { return LBM_B200_ABI_VERSION; }

int lbm_device_count(int *count) {
  if (!count) return fail(LBM_E_ARG, "count is null");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(LBM_E_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
  }
  return LBM_OK;
}

int lbm_create_on_devices(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel,
                          float omega, int num_slabs, const int *devices) {
  if (!out || !devices) return fail(LBM_E_ARG, "null argument");
  *out = nullptr;
  int rc = check_problem(nx, ny, max_iters, num_slabs);
  if (rc) return rc;
  int ndev = 0;
  rc = lbm_device_count(&ndev);
  if (rc) return rc;
  if (ndev < 1) return fail(LBM_E_CUDA, "no CUDA device present; this library has no CPU path");
  for (int i = 0; i < num_slabs; ++i)
    if (devices[i] < 0 || devices[i] >= ndev)
      return fail(LBM_E_ARG, "slab %d asks for GPU %d but only %d are visible", i, devices[i], ndev);
  lbm_handle *h = new (std::nothrow) lbm_handle();
  if (!h) return fail(LBM_E_NOMEM, "out of host memory");
  set_geometry(h, nx, ny, max_iters, density, accel, omega, num_slabs);
  h->slabs.resize(num_slabs);
  for (int i = 0; i < num_slabs; ++i) {
    Slab &s = h->slabs[i];
    s.device = devices[i];
    lbm::partition::range_of(ny, num_slabs, i, s.row0, s.nyl);
    s.nxl = nx;
    s.ntx_sweep = h->ntx;
    // slabs sharing a device share one stream, so their kernels run back to back (no co-scheduling
    // assumption between kernels that wait on each other)
    for (int j = 0; j < i; ++j)
      if (h->slabs[j].device == s.device) { s.stream = h->slabs[j].stream; break; }
    if (!s.stream) {
      cudaError_t e = cudaSetDevice(s.device);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) {
        lbm_destroy(h);
        return fail(LBM_E_CUDA, "stream creation on GPU %d: %s", s.device, cudaGetErrorString(e));
      }
      s.owns_stream = true;
    }
  }
  rc = create_common(out, h);
  if (rc) return rc;
  rc = wire_local_ring(h);
  if (rc) { lbm_destroy(h); *out = nullptr; return rc; }
  return LBM_OK;
}

int lbm_create(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel, float omega,
               int num_gpus) {
  if (num_gpus < 1 || num_gpus > 64) return fail(LBM_E_ARG, "num_gpus must be in 1..64 (got %d)", num_gpus);
  int ndev = 0;
  int rc = lbm_device_count(&ndev);
  if (rc) return rc;
  if (num_gpus > ndev) return fail(LBM_E_ARG, "asked for %d GPUs but only %d are visible", num_gpus, ndev);
  std::vector<int> dev(num_gpus);
  for (int i = 0; i < num_gpus; ++i) dev[i] = i;
  return lbm_create_on_devices(out, nx, ny, max_iters, density, accel, omega, num_gpus, dev.data());
}

}  // extern "C"
namespace {
// wire a px x py block partition inside one process: block (bx, by) sits at index by * px + bx; both directions are
// periodic, so with px == 1 (or py == 1) a block is its own east / west (north / south) neighbour
int wire_local_grid(lbm_handle *h) {
  const int px = h->px, py = h->py;
  auto at = [&](int bx, int by) -> Slab & { return h->slabs[(size_t)((by + py) % py) * px + (bx + px) % px]; };
  for (int by = 0; by < py; ++by)
    for (int bx = 0; bx < px; ++bx) {
      Slab &s = at(bx, by);
      Slab &so = at(bx, by - 1), &no = at(bx, by + 1);
      for (int b = 0; b < 2; ++b) {
        s.south_buf[b] = so.buf[b];
        s.north_buf[b] = no.buf[b];
      }
      s.flag_to_south = so.flags + 1;  // its "from north"
      s.flag_to_north = no.flags + 0;  // its "from south"
      s.south_ghost_row = so.nyl + 1;
      // side i of this block -> (neighbour, the side of the neighbour this block is on)
      const struct { int dx, dy, back; } sides[6] = {{-1, 0, lbm::SIDE_E},  {1, 0, lbm::SIDE_W},   {-1, -1, lbm::SIDE_NE},
                                                     {1, -1, lbm::SIDE_NW}, {-1, 1, lbm::SIDE_SE}, {1, 1, lbm::SIDE_SW}};
      for (int i = 0; i < 6; ++i) {
        Slab &nb = at(bx + sides[i].dx, by + sides[i].dy);
        for (int b = 0; b < 2; ++b) s.side_buf[i][b] = nb.buf[b];
        s.flag_to_side[i] = nb.flags + 2 + sides[i].back;
      }
    }
  // peer access between every pair of distinct devices that are neighbours (8-neighbourhood)
  for (int by = 0; by < py; ++by)
    for (int bx = 0; bx < px; ++bx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int me = at(bx, by).device, d = at(bx + dx, by + dy).device;
          if (d == me) continue;
          int can = 0;
          CUDA_TRY(cudaDeviceCanAccessPeer(&can, me, d));
          if (!can) return fail(LBM_E_PEER, "GPU %d cannot access GPU %d as a peer", me, d);
          CUDA_TRY(cudaSetDevice(me));
          cudaError_t e = cudaDeviceEnablePeerAccess(d, 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else if (e != cudaSuccess)
            return fail(LBM_E_PEER, "cudaDeviceEnablePeerAccess(%d -> %d): %s", me, d, cudaGetErrorString(e));
        }
  h->connected = true;
  return LBM_OK;
}
}  // namespace
extern "C" {

int lbm_create_grid(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel, float omega, int px,
                    int py, const int *devices) {
  if (!out) return fail(LBM_E_ARG, "null argument");
  *out = nullptr;
  if (px < 1 || py < 1 || px > 64 || py > 64) return fail(LBM_E_ARG, "the block grid must be 1..64 x 1..64 (got %d x %d)", px, py);
  int rc = check_problem(nx, ny, max_iters, py);
  if (rc) return rc;
  if (nx < px) return fail(LBM_E_ARG, "cannot cut %d columns into %d blocks", nx, px);
  if ((ny / py + 3 + lbm::kTileH) / lbm::kTileH > 65535)
    return fail(LBM_E_ARG, "blocks of %d rows exceed the 65535 tile rows of a launch grid: use more block rows", ny / py + 1);
  int ndev = 0;
  rc = lbm_device_count(&ndev);
  if (rc) return rc;
  if (ndev < 1) return fail(LBM_E_CUDA, "no CUDA device present; this library has no CPU path");
  const int n = px * py;
  std::vector<int> dev(n);
  for (int i = 0; i < n; ++i) {
    dev[i] = devices ? devices[i] : i;
    if (dev[i] < 0 || dev[i] >= ndev)
      return fail(LBM_E_ARG, "block %d asks for GPU %d but only %d are visible", i, dev[i], ndev);
  }
  lbm_handle *h = new (std::nothrow) lbm_handle();
  if (!h) return fail(LBM_E_NOMEM, "out of host memory");
  set_geometry(h, nx, ny, max_iters, density, accel, omega, py, px, true);
  h->slabs.resize(n);
  for (int by = 0; by < py; ++by)
    for (int bx = 0; bx < px; ++bx) {
      const int i = by * px + bx;
      Slab &s = h->slabs[i];
      s.device = dev[i];
      lbm::partition::range_of(ny, py, by, s.row0, s.nyl);
      lbm::partition::range_of(nx, px, bx, s.col0, s.nxl);
      s.ntx_sweep = (s.nxl + lbm::kTileW - 1) / lbm::kTileW;
      for (int j = 0; j < i; ++j)
        if (h->slabs[j].device == s.device) { s.stream = h->slabs[j].stream; break; }
      if (!s.stream) {
        cudaError_t e = cudaSetDevice(s.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
          lbm_destroy(h);
          return fail(LBM_E_CUDA, "stream creation on GPU %d: %s", s.device, cudaGetErrorString(e));
        }
        s.owns_stream = true;
      }
    }
  rc = create_common(out, h);
  if (rc) return rc;
  rc = wire_local_grid(h);
  if (rc) { lbm_destroy(h); *out = nullptr; return rc; }
  return LBM_OK;
}

int lbm_create_rank(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel,
                    float omega, int rank, int world, int device) {
  if (!out) return fail(LBM_E_ARG, "null argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(LBM_E_ARG, "bad rank %d of %d", rank, world);
  int rc = check_problem(nx, ny, max_iters, world);
  if (rc) return rc;
  int ndev = 0;
  rc = lbm_device_count(&ndev);
  if (rc) return rc;
  if (device < 0 || device >= ndev) return fail(LBM_E_ARG, "GPU %d not visible (%d devices)", device, ndev);
  lbm_handle *h = new (std::nothrow) lbm_handle();
  if (!h) return fail(LBM_E_NOMEM, "out of host memory");
  set_geometry(h, nx, ny, max_iters, density, accel, omega, world);
  h->rank_mode = true;
  h->rank = rank;
  h->world = world;
  h->slabs.resize(1);
  Slab &s = h->slabs[0];
  s.device = device;
  lbm::partition::range_of(ny, world, rank, s.row0, s.nyl);
  s.nxl = nx;
  s.ntx_sweep = h->ntx;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return fail(LBM_E_CUDA, "stream creation on GPU %d: %s", device, cudaGetErrorString(e));
  }
  s.owns_stream = true;
  rc = create_common(out, h);
  if (rc) return rc;
  if (world == 1) return wire_local_ring(h);
  return LBM_OK;
}

}  // extern "C"
namespace {
// one process per GPU, block partition: map the (up to eight distinct) neighbouring ranks' allocations and point this
// block's halo stores and flags at them.  A neighbour that is this rank itself (px or py of 1 or 2 make that common) needs
// no mapping.
int connect_block_neighbours(lbm_handle *h, const char *all) {
  Slab &s = h->slabs[0];
  const int px = h->px, py = h->py, bx = h->rank % px, by = h->rank / px;
  auto rank_of = [&](int dx, int dy) { return ((by + dy + py) % py) * px + (bx + dx + px) % px; };
  std::vector<std::pair<int, char *>> mapped;  // rank -> base of its allocation as seen from here
  std::vector<int> nyl_of(h->world, 0);
  auto base_of = [&](int r, char **base) -> int {
    if (r == h->rank) { *base = s.base; nyl_of[r] = s.nyl; return LBM_OK; }
    for (auto &m : mapped)
      if (m.first == r) { *base = m.second; return LBM_OK; }
    const char *b = all + (size_t)r * blob_bytes();
    cudaIpcMemHandle_t mh;
    int meta[4];
    memcpy(&mh, b, sizeof mh);
    memcpy(meta, b + sizeof mh, sizeof meta);
    if (meta[0] != r || meta[2] != h->tiles_y || meta[3] != h->ntx)
      return fail(LBM_E_ARG, "IPC blob %d does not describe rank %d of the same problem", r, r);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(LBM_E_PEER, "cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
    h->ipc_opened.push_back(ptr);
    mapped.emplace_back(r, static_cast<char *>(ptr));
    nyl_of[r] = meta[1];
    *base = static_cast<char *>(ptr);
    return LBM_OK;
  };
  auto lattice = [&](char *base, int b) { return reinterpret_cast<float *>(base + (size_t)b * h->lattice_bytes); };
  auto flags = [&](char *base) { return reinterpret_cast<int *>(base + 2 * h->lattice_bytes); };
  char *so = nullptr, *no = nullptr;
  int rc = base_of(rank_of(0, -1), &so);
  if (rc == LBM_OK) rc = base_of(rank_of(0, 1), &no);
  if (rc) return rc;
  for (int b = 0; b < 2; ++b) {
    s.south_buf[b] = lattice(so, b);
    s.north_buf[b] = lattice(no, b);
  }
  s.flag_to_south = flags(so) + 1;
  s.flag_to_north = flags(no) + 0;
  s.south_ghost_row = nyl_of[rank_of(0, -1)] + 1;
  const struct { int dx, dy, back; } sides[6] = {{-1, 0, lbm::SIDE_E},  {1, 0, lbm::SIDE_W},   {-1, -1, lbm::SIDE_NE},
                                                 {1, -1, lbm::SIDE_NW}, {-1, 1, lbm::SIDE_SE}, {1, 1, lbm::SIDE_SW}};
  for (int i = 0; i < 6; ++i) {
    char *nb = nullptr;
    rc = base_of(rank_of(sides[i].dx, sides[i].dy), &nb);
    if (rc) return rc;
    for (int b = 0; b < 2; ++b) s.side_buf[i][b] = lattice(nb, b);
    s.flag_to_side[i] = flags(nb) + 2 + sides[i].back;
  }
  h->connected = true;
  return LBM_OK;
}
}  // namespace
extern "C" {

int lbm_create_rank_grid(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel, float omega, int rank,
                         int px, int py, int device) {
  if (!out) return fail(LBM_E_ARG, "null argument");
  *out = nullptr;
  if (px < 1 || py < 1 || px > 64 || py > 64) return fail(LBM_E_ARG, "the block grid must be 1..64 x 1..64 (got %d x %d)", px, py);
  const int world = px * py;
  if (rank < 0 || rank >= world) return fail(LBM_E_ARG, "bad rank %d of %d", rank, world);
  int rc = check_problem(nx, ny, max_iters, py);
  if (rc) return rc;
  if (nx < px) return fail(LBM_E_ARG, "cannot cut %d columns into %d blocks", nx, px);
  if ((ny / py + 3 + lbm::kTileH) / lbm::kTileH > 65535)
    return fail(LBM_E_ARG, "blocks of %d rows exceed the 65535 tile rows of a launch grid: use more block rows", ny / py + 1);
  int ndev = 0;
  rc = lbm_device_count(&ndev);
  if (rc) return rc;
  if (device < 0 || device >= ndev) return fail(LBM_E_ARG, "GPU %d not visible (%d devices)", device, ndev);
  lbm_handle *h = new (std::nothrow) lbm_handle();
  if (!h) return fail(LBM_E_NOMEM, "out of host memory");
  set_geometry(h, nx, ny, max_iters, density, accel, omega, py, px, true);
  h->rank_mode = true;
  h->rank = rank;
  h->world = world;
  h->slabs.resize(1);
  Slab &s = h->slabs[0];
  s.device = device;
  lbm::partition::range_of(ny, py, rank / px, s.row0, s.nyl);
  lbm::partition::range_of(nx, px, rank % px, s.col0, s.nxl);
  s.ntx_sweep = (s.nxl + lbm::kTileW - 1) / lbm::kTileW;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return fail(LBM_E_CUDA, "stream creation on GPU %d: %s", device, cudaGetErrorString(e));
  }
  s.owns_stream = true;
  rc = create_common(out, h);
  if (rc) return rc;
  if (world == 1) return wire_local_grid(h);  // a 1 x 1 "grid" is its own eight neighbours
  return LBM_OK;
}

size_t lbm_ipc_blob_bytes(void) { return blob_bytes(); }

int lbm_ipc_export(lbm_handle *h, void *blob) {
  if (!h || !blob) return fail(LBM_E_ARG, "null argument");
  if (!h->rank_mode) return fail(LBM_E_STATE, "lbm_ipc_export needs a handle from lbm_create_rank");
  Slab &s = h->slabs[0];
  CUDA_TRY(cudaSetDevice(s.device));
  cudaIpcMemHandle_t mh;
  CUDA_TRY(cudaIpcGetMemHandle(&mh, s.base));
  char *p = static_cast<char *>(blob);
  memcpy(p, &mh, sizeof mh);
  const int meta[4] = {h->rank, s.nyl, h->tiles_y, h->ntx};
  memcpy(p + sizeof mh, meta, sizeof meta);
  return LBM_OK;
}

int lbm_ipc_connect(lbm_handle *h, const void *blobs) {
  if (!h || !blobs) return fail(LBM_E_ARG, "null argument");
  if (!h->rank_mode) return fail(LBM_E_STATE, "lbm_ipc_connect needs a handle from lbm_create_rank");
  if (h->connected) return fail(LBM_E_STATE, "already connected");
  Slab &s = h->slabs[0];
  CUDA_TRY(cudaSetDevice(s.device));
  if (h->xghost) return connect_block_neighbours(h, static_cast<const char *>(blobs));
  const int w = h->world, south = (h->rank + w - 1) % w, north = (h->rank + 1) % w;
  const char *all = static_cast<const char *>(blobs);
  char *mapped[2] = {nullptr, nullptr};
  int peer_nyl[2] = {0, 0};
  const int peers[2] = {south, north};
  for (int i = 0; i < 2; ++i) {
    if (i == 1 && north == south) { mapped[1] = mapped[0]; peer_nyl[1] = peer_nyl[0]; break; }
    const char *b = all + (size_t)peers[i] * blob_bytes();
    cudaIpcMemHandle_t mh;
    int meta[4];
    memcpy(&mh, b, sizeof mh);
    memcpy(meta, b + sizeof mh, sizeof meta);
    if (meta[0] != peers[i] || meta[2] != h->tiles_y || meta[3] != h->ntx)
      return fail(LBM_E_ARG, "IPC blob %d does not describe rank %d of the same problem", peers[i], peers[i]);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(LBM_E_PEER, "cudaIpcOpenMemHandle for rank %d: %s", peers[i], cudaGetErrorString(e));
    h->ipc_opened.push_back(ptr);
    mapped[i] = static_cast<char *>(ptr);
    peer_nyl[i] = meta[1];
  }
  for (int b = 0; b < 2; ++b) {
    s.south_buf[b] = reinterpret_cast<float *>(mapped[0] + (size_t)b * h->lattice_bytes);
    s.north_buf[b] = reinterpret_cast<float *>(mapped[1] + (size_t)b * h->lattice_bytes);
  }
  s.flag_to_south = reinterpret_cast<int *>(mapped[0] + 2 * h->lattice_bytes) + 1;
  s.flag_to_north = reinterpret_cast<int *>(mapped[1] + 2 * h->lattice_bytes) + 0;
  s.south_ghost_row = peer_nyl[0] + 1;
  h->connected = true;
  return LBM_OK;
}

// Teardown of the one-process-per-GPU mode is two-phase (see the protocol in include/lbm_b200.h): every rank first
// drops its mappings of the neighbours' slabs, the ranks meet at a barrier, and only then does anyone free the
// allocation its neighbours had mapped (freeing memory that another process still has open through
// cudaIpcOpenMemHandle is undefined behaviour).
int lbm_ipc_disconnect(lbm_handle *h) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (!h->rank_mode) return fail(LBM_E_STATE, "lbm_ipc_disconnect needs a handle from lbm_create_rank");
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    if (s.stream) CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (s.copy_stream) CUDA_TRY(cudaStreamSynchronize(s.copy_stream));
  }
  for (void *p : h->ipc_opened) CUDA_TRY(cudaIpcCloseMemHandle(p));
  h->ipc_opened.clear();
  if (h->world > 1) {
    Slab &s = h->slabs[0];
    for (int b = 0; b < 2; ++b) s.south_buf[b] = s.north_buf[b] = nullptr;
    s.flag_to_south = s.flag_to_north = nullptr;
    for (int i = 0; i < 6; ++i) {
      s.side_buf[i][0] = s.side_buf[i][1] = nullptr;
      s.flag_to_side[i] = nullptr;
    }
    h->connected = false;  // compute calls now fail with LBM_E_STATE instead of touching unmapped peer memory
  }
  return LBM_OK;
}

void lbm_destroy(lbm_handle *h) {
  if (!h) return;
  for (Slab &s : h->slabs) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
  }
  if (!h->slabs.empty()) cudaSetDevice(h->slabs[0].device);
  for (StreamGraph &g : h->graphs)
    for (cudaGraphExec_t e : g.exec)
      if (e) { cudaSetDevice(g.device); cudaGraphExecDestroy(e); }
  if (!h->slabs.empty()) cudaSetDevice(h->slabs[0].device);
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);  // (none left after lbm_ipc_disconnect)
  for (Slab &s : h->slabs) free_slab(s);
  cudaGetLastError();
  delete h;
}

// ---- run(0) --------------------------------------------------------------------------------------
}  // extern "C"
namespace {

// Global rows [g0, g0 + n) of the problem -> the slabs / blocks that own them.  `cells` / `obstacles` point at row g0 of
// full-width rows (nx cells apart), from which every block takes its own columns.
int upload_rows(lbm_handle *h, int g0, int n, const float *cells, const uint8_t *obstacles) {
  const int nx = h->rank_block() ? h->slabs[0].nxl : h->nx;  // cells per host row
  for (Slab &s : h->slabs) {
    const int lo = std::max(g0, s.row0), hi = std::min(g0 + n, s.row0 + s.nyl);
    if (lo >= hi) continue;
    CUDA_TRY(cudaSetDevice(s.device));
    const int col = h->rank_block() ? 0 : s.col0;
    const float *c = cells + ((size_t)(lo - g0) * nx + col) * lbm::kSpeeds;
    const uint8_t *o = obstacles + (size_t)(lo - g0) * nx + col;
    const int first_storage_row = 1 + (lo - s.row0), nxl = s.nxl;
    // cells: AoS rows -> tiles; obstacles: bytes -> bits.  Chunks of whole rows, PCIe copy of one half of the
    // staging buffer overlapped with the transposition of the other
    int rc = stage_in(s, reinterpret_cast<const char *>(c), (size_t)nx * lbm::kSpeeds * sizeof(float),
                      (size_t)nxl * lbm::kSpeeds * sizeof(float), hi - lo, s.stage_chunk, [&](char *stage, int r, int rows) {
                        lbm::aos_to_tiles_kernel<<<rows * s.ntx_sweep, lbm::kTileW, 0, s.stream>>>(
                            reinterpret_cast<const float *>(stage), s.buf[0], h->ntx, nxl, rows, first_storage_row + r, s.ntx_sweep);
                        ++h->launches;
                        KERNEL_CHECK(s.stream, "aos_to_tiles");
                        return (int)LBM_OK;
                      });
    if (rc) return rc;
    const int words = (nxl + 31) / 32;
    rc = stage_in(s, reinterpret_cast<const char *>(o), (size_t)nx, (size_t)nxl, hi - lo, s.stage_chunk, [&](char *stage, int r, int rows) {
      const long long threads = (long long)rows * words * 32;
      lbm::pack_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s.stream>>>(
          reinterpret_cast<const uint8_t *>(stage), s.mask, h->ntx, nxl, rows, first_storage_row + r, s.free_dev, s.count_lo, s.count_hi);
      ++h->launches;
      KERNEL_CHECK(s.stream, "pack_mask");
      return (int)LBM_OK;
    });
    if (rc) return rc;
  }
  return LBM_OK;
}
// all slabs' pipelines were only enqueued (they overlap across GPUs); wait for them
int drain_all(lbm_handle *h) {
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    int rc = drain_stages(s);
    if (rc) return rc;
    s.stage_chunk = 0;
  }
  return LBM_OK;
}
// Global rows [g0, g0 + n) of the current state -> host AoS / obstacle bytes (either may be null), full-width rows
int download_rows(lbm_handle *h, int g0, int n, float *cells, uint8_t *obstacles) {
  const int nx = h->rank_block() ? h->slabs[0].nxl : h->nx;  // cells per host row
  for (Slab &s : h->slabs) {
    const int lo = std::max(g0, s.row0), hi = std::min(g0 + n, s.row0 + s.nyl);
    if (lo >= hi) continue;
    CUDA_TRY(cudaSetDevice(s.device));
    const int first_storage_row = 1 + (lo - s.row0), nxl = s.nxl;
    if (cells) {
      const int col = h->rank_block() ? 0 : s.col0;
      int rc = stage_out(s, reinterpret_cast<char *>(cells + ((size_t)(lo - g0) * nx + col) * lbm::kSpeeds),
                         (size_t)nx * lbm::kSpeeds * sizeof(float), (size_t)nxl * lbm::kSpeeds * sizeof(float), hi - lo, s.stage_chunk,
                         [&](char *stage, int r, int rows) {
                           lbm::tiles_to_aos_kernel<<<rows * s.ntx_sweep, lbm::kTileW, 0, s.stream>>>(
                               s.buf[h->cur], reinterpret_cast<float *>(stage), h->ntx, nxl, rows, first_storage_row + r, s.ntx_sweep);
                           ++h->launches;
                           KERNEL_CHECK(s.stream, "tiles_to_aos");
                           return (int)LBM_OK;
                         });
      if (rc) return rc;
    }
    if (obstacles) {
      const int col = h->rank_block() ? 0 : s.col0;
      int rc = stage_out(s, reinterpret_cast<char *>(obstacles + (size_t)(lo - g0) * nx + col), (size_t)nx, (size_t)nxl, hi - lo,
                         s.stage_chunk, [&](char *stage, int r, int rows) {
                           const long long cellsn = (long long)rows * nxl;
                           lbm::unpack_mask_kernel<<<(unsigned)((cellsn + 255) / 256), 256, 0, s.stream>>>(
                               s.mask, reinterpret_cast<uint8_t *>(stage), h->ntx, nxl, rows, first_storage_row + r);
                           ++h->launches;
                           KERNEL_CHECK(s.stream, "unpack_mask");
                           return (int)LBM_OK;
                         });
      if (rc) return rc;
    }
  }
  return LBM_OK;
}

}  // namespace
extern "C" {

int lbm_upload(lbm_handle *h, const float *cells_aos, const uint8_t *obstacles) {
  if (!h || !cells_aos || !obstacles) return fail(LBM_E_ARG, "null argument");
  if (!h->connected) return fail(LBM_E_STATE, "ranks are not connected yet (lbm_ipc_connect)");
  int rc = reset_counters(h);
  if (rc) return rc;
  // in rank mode the caller passes only the owned rows
  if (h->rank_mode) rc = upload_rows(h, h->slabs[0].row0, h->slabs[0].nyl, cells_aos, obstacles);
  else rc = upload_rows(h, 0, h->ny, cells_aos, obstacles);
  if (rc) return rc;
  rc = drain_all(h);
  if (rc) return rc;
  return finish_upload(h);
}

int lbm_init_synthetic(lbm_handle *h, uint64_t seed, float solid_prob) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (!(solid_prob >= 0.f && solid_prob <= 1.f)) return fail(LBM_E_ARG, "solid_prob must be in [0,1]");
  if (!h->connected) return fail(LBM_E_STATE, "ranks are not connected yet (lbm_ipc_connect)");
  int rc = reset_counters(h);
  if (rc) return rc;
  const uint32_t threshold24 = (uint32_t)((double)solid_prob * 16777216.0);
  // Cells::initialise, main/include/LatticeBoltzmannUtils.hpp:137-140
  const float w0 = h->density * 4.f / 9.f, w1 = h->density / 9.f, w2 = h->density / 36.f;
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    lbm::fill_lattice_kernel<<<148 * 8, 256, 0, s.stream>>>(s.buf[0], h->lattice_floats, w0, w1, w2);
    const int words = (s.nxl + 31) / 32;
    const long long threads = (long long)s.nyl * words * 32;
    lbm::synthetic_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s.stream>>>(
        s.mask, h->ntx, h->nx, h->ny, s.row0, s.nyl, seed, threshold24, s.free_dev, s.col0, s.nxl);
    h->launches += 2;
    KERNEL_CHECK(s.stream, "fill_lattice/synthetic_mask");
  }
  return finish_upload(h);
}

// ---- run(1) --------------------------------------------------------------------------------------
}  // extern "C"
namespace {

// slabs that share a stream (several slabs on one device) alternate in it, so a sweep is not preceded by a
// sweep of the same slab: no dependent-launch overlap there (the flag waits need the other slab to finish)
bool shares_a_stream(const lbm_handle *h) {
  for (size_t i = 0; i < h->slabs.size(); ++i)
    for (size_t j = 0; j < i; ++j)
      if (h->slabs[i].stream == h->slabs[j].stream) return true;
  return false;
}

// `n` timesteps starting at step number t0 (relative to *t_base while capturing), one launch per slab per timestep,
// slabs interleaved so no GPU's queue runs ahead of its neighbours
int enqueue_steps(lbm_handle *h, int t0, int n, bool shared_stream) {
  for (int t = 0; t < n; ++t) {
    for (Slab &s : h->slabs) {
      if (h->slabs.size() > 1) CUDA_TRY(cudaSetDevice(s.device));
      int rc = launch_sweep(h, s, t0 + t, t > 0 && !shared_stream);
      if (rc) return rc;
    }
    h->cur ^= 1;
  }
  for (Slab &s : h->slabs) {
    if (h->slabs.size() > 1) CUDA_TRY(cudaSetDevice(s.device));
    int rc = flush_reductions(h, s);
    if (rc) return rc;
  }
  return LBM_OK;
}

// Capture kGraphSteps timesteps of the loop above once per distinct stream and keep the instantiated graphs in the
// handle ("prepare once, run many": a re-run or a long run replays them instead of re-enqueueing every launch; cf. the
// reference's whole time loop being one pre-compiled device program, main/LbmAoS.cpp:353-356).  The captured kernels
// read the step number from s.step[0] (set before the first replay, advanced by the last node of every replay), so
// one graph serves any position in the run; it bakes in the lattice parity, hence one graph per parity.
// On any failure graphs are switched off for this parity and the caller falls back to plain launches.
int build_graphs(lbm_handle *h, int parity) {
  if (h->graphs.empty()) {
    for (Slab &s : h->slabs) {
      bool seen = false;
      for (StreamGraph &g : h->graphs) seen |= (g.stream == s.stream);
      if (seen) continue;
      StreamGraph g;
      g.device = s.device;
      g.stream = s.stream;
      h->graphs.push_back(g);
    }
  }
  const int cur0 = h->cur;
  const long long launches0 = h->launches;
  std::vector<int> first0;
  for (Slab &s : h->slabs) first0.push_back(s.window_first);
  bool ok = true;
  size_t begun = 0;
  for (StreamGraph &g : h->graphs) {
    if (cudaSetDevice(g.device) != cudaSuccess || cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      ok = false;
      break;
    }
    ++begun;
  }
  h->graph_capture = true;
  if (ok) ok = enqueue_steps(h, 0, kGraphSteps, shares_a_stream(h)) == LBM_OK;
  if (ok) {
    for (Slab &s : h->slabs) {
      if (cudaSetDevice(s.device) != cudaSuccess) { ok = false; break; }
      lbm::advance_step_kernel<<<1, 1, 0, s.stream>>>(s.step, kGraphSteps);
      ++h->launches;
      ok = ok && cudaGetLastError() == cudaSuccess;
    }
  }
  h->graph_capture = false;
  for (size_t i = 0; i < begun; ++i) {
    StreamGraph &g = h->graphs[i];
    cudaGraph_t graph = nullptr;
    cudaSetDevice(g.device);
    if (cudaStreamEndCapture(g.stream, &graph) != cudaSuccess || !graph) ok = false;
    if (ok && cudaGraphInstantiate(&g.exec[parity], graph, 0) != cudaSuccess) {
      g.exec[parity] = nullptr;
      ok = false;
    }
    if (graph) cudaGraphDestroy(graph);
  }
  cudaGetLastError();
  // nothing ran: put the bookkeeping back
  const long long per_replay = h->launches - launches0;
  h->launches = launches0;
  h->cur = cur0;
  for (size_t i = 0; i < h->slabs.size(); ++i) {
    h->slabs[i].window_first = first0[i];
    h->slabs[i].window_count = 0;
  }
  if (!ok) {
    for (StreamGraph &g : h->graphs)
      if (g.exec[parity]) { cudaGraphExecDestroy(g.exec[parity]); g.exec[parity] = nullptr; }
    h->graph_state[parity] = -1;
    return LBM_OK;
  }
  for (StreamGraph &g : h->graphs) g.nodes = per_replay / (long long)h->graphs.size();
  h->graph_state[parity] = 1;
  return LBM_OK;
}

bool graphs_wanted(const lbm_handle *h, int iters) {
  const int env = getenv("LBM_B200_GRAPH") ? atoi(getenv("LBM_B200_GRAPH")) : -1;  // 0 off, 1 force, default auto
  if (env == 0 || g_debug_sync || iters < kGraphSteps || h->no_graphs) return false;
  if (env == 1) return true;
  for (const Slab &s : h->slabs)
    if ((long long)s.nyl * s.nxl > kGraphMaxCells) return false;
  return true;
}

}  // namespace
extern "C" {

int lbm_run(lbm_handle *h, int iters) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (!h->uploaded) return fail(LBM_E_STATE, "lbm_run before lbm_upload");
  if (!h->connected) return fail(LBM_E_STATE, "ranks are not connected (lbm_ipc_connect; or already disconnected)");
  if (iters < 0 || h->steps_done + iters > h->max_iters)
    return fail(LBM_E_ARG, "cannot run %d more steps: %d of max_iters=%d already done", iters, h->steps_done,
                h->max_iters);
  if (g_debug_sync && h->xghost)
    for (const Slab &s : h->slabs)
      if (s.stream != h->slabs[0].stream)
        return fail(LBM_E_STATE, "LBM_B200_DEBUG_SYNC cannot be used with a block partition on several GPUs: a block launch waits for "
                                 "its neighbours' launch of the same timestep, which a synchronisation after every launch prevents");
  const bool clustered = h->slabs.size() == 1 && h->slabs[0].cluster_ctas > 0 && iters > 0;
  const bool persistent = !clustered && h->slabs.size() == 1 && h->slabs[0].persistent_grid > 0 && iters > 0;
  // build (once) the replayable time loop before the timed region starts
  const bool use_graphs = !persistent && !clustered && graphs_wanted(h, iters);
  if (use_graphs && h->graph_state[h->cur] == 0) {
    int rc = build_graphs(h, h->cur);
    if (rc) return rc;
  }
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
  }
  int iters_left_for_loop = iters;
  if (h->first_run_pending) {
    // accelerate_flow once before step 0 (main/LbmAoS.cpp:352; main/LastChance.cpp:158-183) ...
    const float w1 = h->density * h->accel / 9.f, w2 = h->density * h->accel / 36.f;
    const int arow = h->accel_row_global;
    for (Slab &s : h->slabs) {
      if (arow < s.row0 || arow >= s.row0 + s.nyl) continue;
      CUDA_TRY(cudaSetDevice(s.device));
      lbm::accelerate_flow_kernel<<<(s.nxl + 255) / 256, 256, 0, s.stream>>>(s.buf[h->cur], h->ntx, s.nxl, arow - s.row0 + 1,
                                                                             s.mask, w1, w2);
      ++h->launches;
      KERNEL_CHECK(s.stream, "accelerate_flow");
    }
    h->first_run_pending = false;
  }
  if (clustered) {
    // a single tiny slab: the lattices move into the shared memory of thread-block clusters for the whole call.  The
    // kernel works on the owned rows only (y wraps by index), so the HBM ghost rows are stale afterwards.
    Slab &s = h->slabs[0];
    int rc = launch_cluster(h, s, iters);
    if (rc) return rc;
    if (s.cluster_ctas > 0) {
      iters_left_for_loop = 0;
      h->halo_pending = true;
    } else {
      rc = configure_launch(h, s);  // the launch was refused: pick the persistent / per-step path instead
      if (rc) return rc;
    }
  }
  const bool persistent_now = iters_left_for_loop > 0 && h->slabs.size() == 1 && h->slabs[0].persistent_grid > 0;
  if (h->halo_pending && iters_left_for_loop > 0) {
    // ... then the initial halo of a fresh (or re-loaded) lattice: every slab fills its neighbours' ghost rows and
    // publishes the epoch the next timestep waits for
    for (Slab &s : h->slabs) {
      CUDA_TRY(cudaSetDevice(s.device));
      lbm::init_halo_kernel<<<(s.nxl + 255) / 256, 256, 0, s.stream>>>(s.buf[h->cur], s.north_buf[h->cur], s.south_buf[h->cur],
                                                                       h->ntx, s.nxl, s.nyl, s.south_ghost_row);
      ++h->launches;
      lbm::SidePointers side{};
      if (h->xghost) {  // 2-D blocks: ghost columns and corners as well
        for (int i = 0; i < 6; ++i) {
          side.dst[i] = s.side_buf[i][h->cur];
          side.flag[i] = s.flag_to_side[i];
        }
        lbm::init_halo_sides_kernel<<<(s.nyl + 255) / 256, 256, 0, s.stream>>>(s.buf[h->cur], side, h->ntx, s.nxl, s.nyl, s.south_ghost_row);
        ++h->launches;
      }
      if (h->multi) {
        lbm::publish_epoch_kernel<<<1, 1, 0, s.stream>>>(s.flag_to_south, s.flag_to_north, h->steps_done + 1);
        ++h->launches;
        if (h->xghost) {
          lbm::publish_epoch_sides_kernel<<<1, 1, 0, s.stream>>>(side, h->steps_done + 1);
          ++h->launches;
        }
      }
      KERNEL_CHECK(s.stream, "init_halo/publish_epoch");
    }
    h->halo_pending = false;
  }
  if (persistent_now) {
    // a single small slab: the whole time loop is one cooperative launch
    Slab &s = h->slabs[0];
    int rc = launch_persistent(h, s, iters);
    if (rc) return rc;
    if (s.persistent_grid > 0) {  // (reset to 0 when the cooperative launch was refused: fall through to the loop)
      h->cur ^= (iters & 1);
      iters_left_for_loop = 0;
    }
  }
  int t = 0;
  if (use_graphs && h->graph_state[h->cur] == 1 && iters_left_for_loop >= kGraphSteps) {
    const int parity = h->cur;
    for (Slab &s : h->slabs) {
      CUDA_TRY(cudaSetDevice(s.device));
      lbm::set_step_kernel<<<1, 1, 0, s.stream>>>(s.step, h->steps_done);
      ++h->launches;
      KERNEL_CHECK(s.stream, "set_step");
    }
    for (; iters_left_for_loop - t >= kGraphSteps; t += kGraphSteps) {
      for (StreamGraph &g : h->graphs) {
        if (h->graphs.size() > 1) CUDA_TRY(cudaSetDevice(g.device));
        CUDA_TRY(cudaGraphLaunch(g.exec[parity], g.stream));
        h->launches += g.nodes;
      }
    }
    for (Slab &s : h->slabs) s.window_first = h->steps_done + t;
  }
  int rc = enqueue_steps(h, h->steps_done + t, iters_left_for_loop - t, shares_a_stream(h));
  if (rc) return rc;
  h->steps_done += iters;
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
  }
  double ms_max = 0.0;
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    cudaError_t e = cudaEventSynchronize(s.ev1);
    if (e != cudaSuccess) {
      int err[2] = {0, 0};
      cudaMemcpy(err, s.step, sizeof err, cudaMemcpyDeviceToHost);
      return fail(LBM_E_CUDA, "time loop failed on GPU %d: %s%s", s.device, cudaGetErrorString(e),
                  err[1] ? " (timed out waiting for a neighbour slab's halo)" : "");
    }
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    ms_max = std::max(ms_max, (double)ms);
  }
  h->last_ms = ms_max;
  return LBM_OK;
}

int lbm_elapsed_ms(lbm_handle *h, double *ms) {
  if (!h || !ms) return fail(LBM_E_ARG, "null argument");
  *ms = h->last_ms;
  return LBM_OK;
}

// ---- run(2) --------------------------------------------------------------------------------------
int lbm_download_tot_u(lbm_handle *h, double *tot_u) {
  if (!h || !tot_u) return fail(LBM_E_ARG, "null argument");
  if (!h->uploaded) return fail(LBM_E_STATE, "nothing uploaded");
  std::vector<double> part(h->steps_done);
  for (int t = 0; t < h->steps_done; ++t) tot_u[t] = 0.0;
  // slabs are added in slab order: the documented, deterministic reduction order
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (h->steps_done)
      CUDA_TRY(cudaMemcpy(part.data(), s.tot_u, (size_t)h->steps_done * sizeof(double), cudaMemcpyDeviceToHost));
    for (int t = 0; t < h->steps_done; ++t) tot_u[t] += part[t];
  }
  return LBM_OK;
}

int lbm_download(lbm_handle *h, float *cells_aos, float *av_vels) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (!h->uploaded) return fail(LBM_E_STATE, "lbm_download before lbm_upload");
  if (av_vels && h->rank_mode && h->world > 1)
    return fail(LBM_E_ARG, "with one process per GPU av_vels needs all ranks: use lbm_download_tot_u and "
                           "combine in rank order");
  if (cells_aos) {
    int rc = h->rank_mode ? download_rows(h, h->slabs[0].row0, h->slabs[0].nyl, cells_aos, nullptr)
                          : download_rows(h, 0, h->ny, cells_aos, nullptr);
    if (rc) return rc;
    rc = drain_all(h);
    if (rc) return rc;
  }
  if (av_vels) {
    std::vector<double> tot(std::max(h->steps_done, 1));
    int rc = lbm_download_tot_u(h, tot.data());
    if (rc) return rc;
    long long free_cells = 0;
    for (Slab &s : h->slabs) free_cells += s.free_cells;
    // main/LastChance.cpp:266: fp32 quotient of the total by the fp32 free-cell count
    for (int t = 0; t < h->steps_done; ++t) av_vels[t] = (float)tot[t] / (float)free_cells;
  }
  return LBM_OK;
}

int lbm_download_fields(lbm_handle *h, float *fields) {
  if (!h || !fields) return fail(LBM_E_ARG, "null argument");
  if (!h->uploaded) return fail(LBM_E_STATE, "lbm_download_fields before lbm_upload");
  const int nx = h->rank_block() ? h->slabs[0].nxl : h->nx;  // cells per host row
  for (Slab &s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    const size_t first = h->rank_mode ? 0 : (size_t)s.row0;
    const int nxl = s.nxl, col = h->rank_block() ? 0 : s.col0;
    int rc = stage_out(s, reinterpret_cast<char *>(fields + (first * nx + col) * 4), (size_t)nx * 4 * sizeof(float),
                       (size_t)nxl * 4 * sizeof(float), s.nyl, s.stage_chunk, [&](char *stage, int r, int rows) {
                         lbm::final_state_kernel<<<rows * s.ntx_sweep, lbm::kTileW, 0, s.stream>>>(
                             s.buf[h->cur], s.mask, reinterpret_cast<float4 *>(stage), h->ntx, nxl, rows, 1 + r, h->density, s.ntx_sweep);
                         ++h->launches;
                         KERNEL_CHECK(s.stream, "final_state");
                         return (int)LBM_OK;
                       });
    if (rc) return rc;
  }
  return drain_all(h);
}

// ---- prepare once / run many ------------------------------------------------------------------------
// The reference compiles its graph once, serialises it (main/LbmPoplibs.cpp:375-389) and lbm_runner deserialises it
// with --exe (main/LbmRunner.cpp:20,75-84).  A B200 solver has nothing to compile at run time; what is worth keeping
// is the prepared PROBLEM and its STATE.  The file is layout-independent (host interchange formats), so it can be
// loaded onto a different number of GPUs than it was saved from:
//   SaveHeader | obstacles uint8[ny][nx] | cells float[ny][nx][9] (current state) | tot_u double[steps_done]
namespace {
struct SaveHeader {
  char magic[8];
  uint32_t version, header_bytes;
  int32_t nx, ny, max_iters, steps_done;
  float density, accel, omega;
  int32_t first_run_pending;
  int64_t free_cells;
};
constexpr char kSaveMagic[8] = {'L', 'B', 'M', 'B', '2', '0', '0', 0};
constexpr size_t kSaveChunkBytes = 256u << 20;
struct FileCloser {
  FILE *f;
  ~FileCloser() { if (f) fclose(f); }
};
}  // namespace

}  // extern "C"
namespace {
int save_solver(lbm_handle *h, const char *path) {
  if (!h || !path) return fail(LBM_E_ARG, "null argument");
  if (!h->uploaded) return fail(LBM_E_STATE, "lbm_save before lbm_upload");
  if (h->rank_mode && h->world > 1)
    return fail(LBM_E_STATE, "lbm_save needs a handle that owns the whole grid (lbm_create / lbm_create_on_devices)");
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return fail(LBM_E_ARG, "cannot open %s for writing", path);
  SaveHeader hd{};
  memcpy(hd.magic, kSaveMagic, 8);
  hd.version = 1;
  hd.header_bytes = sizeof hd;
  hd.nx = h->nx; hd.ny = h->ny; hd.max_iters = h->max_iters; hd.steps_done = h->steps_done;
  hd.density = h->density; hd.accel = h->accel; hd.omega = h->omega;
  hd.first_run_pending = h->first_run_pending ? 1 : 0;
  hd.free_cells = 0;
  for (Slab &s : h->slabs) hd.free_cells += s.free_cells;
  if (fwrite(&hd, sizeof hd, 1, fc.f) != 1) return fail(LBM_E_ARG, "short write to %s", path);
  const size_t row_bytes = (size_t)h->nx * lbm::kSpeeds * sizeof(float);
  const int rows_per_chunk = (int)std::max<size_t>(1, kSaveChunkBytes / row_bytes);
  std::vector<char> buf((size_t)std::min(rows_per_chunk, h->ny) * row_bytes);
  for (int pass = 0; pass < 2; ++pass) {  // obstacles, then cells
    const size_t rb = pass == 0 ? (size_t)h->nx : row_bytes;
    for (int r = 0; r < h->ny; r += rows_per_chunk) {
      const int n = std::min(rows_per_chunk, h->ny - r);
      int rc = pass == 0 ? download_rows(h, r, n, nullptr, reinterpret_cast<uint8_t *>(buf.data()))
                         : download_rows(h, r, n, reinterpret_cast<float *>(buf.data()), nullptr);
      if (rc == LBM_OK) rc = drain_all(h);
      if (rc) return rc;
      if (fwrite(buf.data(), rb, (size_t)n, fc.f) != (size_t)n) return fail(LBM_E_ARG, "short write to %s", path);
    }
  }
  std::vector<double> tot(std::max(h->steps_done, 1));
  int rc = lbm_download_tot_u(h, tot.data());
  if (rc) return rc;
  if (h->steps_done && fwrite(tot.data(), sizeof(double), (size_t)h->steps_done, fc.f) != (size_t)h->steps_done)
    return fail(LBM_E_ARG, "short write to %s", path);
  FILE *f = fc.f;
  fc.f = nullptr;
  if (fclose(f) != 0) return fail(LBM_E_ARG, "closing %s failed", path);
  return LBM_OK;
}

// px == 0: row slabs on num_gpus GPUs (lbm_create); otherwise px x py blocks (lbm_create_grid)
int load_solver(lbm_handle **out, const char *path, int num_gpus, int px, int py, const int *devices) {
  if (!out || !path) return fail(LBM_E_ARG, "null argument");
  *out = nullptr;
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fail(LBM_E_ARG, "cannot open %s", path);
  SaveHeader hd{};
  if (fread(&hd, sizeof hd, 1, fc.f) != 1 || memcmp(hd.magic, kSaveMagic, 8) != 0)
    return fail(LBM_E_ARG, "%s is not a prepared-solver file", path);
  if (hd.version != 1 || hd.header_bytes != sizeof hd)
    return fail(LBM_E_ARG, "%s has version %u (this library reads version 1)", path, hd.version);
  if (hd.steps_done < 0 || hd.steps_done > hd.max_iters) return fail(LBM_E_ARG, "%s: corrupt step counter", path);
  if (hd.nx < 1 || hd.ny < 1) return fail(LBM_E_ARG, "%s: corrupt grid size", path);
  // host buffers first (they may throw; nothing to undo yet).  The obstacle bytes come first in the file but upload_rows wants
  // both: the obstacles are read whole (1 B per cell), the cells chunk by chunk
  const size_t row_bytes = (size_t)hd.nx * lbm::kSpeeds * sizeof(float);
  const int rows_per_chunk = (int)std::max<size_t>(1, kSaveChunkBytes / row_bytes);
  std::vector<uint8_t> obst((size_t)hd.nx * hd.ny);
  std::vector<char> buf((size_t)std::min(rows_per_chunk, hd.ny) * row_bytes);
  std::vector<double> tot((size_t)std::max(hd.steps_done, 1));
  lbm_handle *h = nullptr;
  int rc = px > 0 ? lbm_create_grid(&h, hd.nx, hd.ny, hd.max_iters, hd.density, hd.accel, hd.omega, px, py, devices)
                  : lbm_create(&h, hd.nx, hd.ny, hd.max_iters, hd.density, hd.accel, hd.omega, num_gpus);
  if (rc) return rc;
  auto bail = [&](int code) { lbm_destroy(h); return code; };
  rc = reset_counters(h);
  if (rc) return bail(rc);
  if (fread(obst.data(), 1, obst.size(), fc.f) != obst.size()) return bail(fail(LBM_E_ARG, "%s is truncated (obstacles)", path));
  for (int r = 0; r < h->ny; r += rows_per_chunk) {
    const int n = std::min(rows_per_chunk, h->ny - r);
    if (fread(buf.data(), row_bytes, (size_t)n, fc.f) != (size_t)n) return bail(fail(LBM_E_ARG, "%s is truncated (cells)", path));
    rc = upload_rows(h, r, n, reinterpret_cast<const float *>(buf.data()), obst.data() + (size_t)r * hd.nx);
    if (rc == LBM_OK) rc = drain_all(h);  // buf is reused by the next chunk
    if (rc) return bail(rc);
  }
  rc = finish_upload(h);
  if (rc) return bail(rc);
  long long free_cells = 0;
  for (Slab &s : h->slabs) free_cells += s.free_cells;
  if (free_cells != hd.free_cells) return bail(fail(LBM_E_ARG, "%s: free-cell count does not match its obstacles", path));
  // resume point: the saved lattice is the current state; per-step sums so far go to slab 0 (x + 0.0 is exact, so the
  // av_vels of the steps already done come back bit for bit whatever the new partition)
  h->steps_done = hd.steps_done;
  h->first_run_pending = hd.first_run_pending != 0;
  if (hd.steps_done) {
    if (fread(tot.data(), sizeof(double), (size_t)hd.steps_done, fc.f) != (size_t)hd.steps_done)
      return bail(fail(LBM_E_ARG, "%s is truncated (tot_u)", path));
    Slab &s0 = h->slabs[0];
    cudaError_t e = cudaSetDevice(s0.device);
    if (e == cudaSuccess) e = cudaMemcpy(s0.tot_u, tot.data(), (size_t)hd.steps_done * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(LBM_E_CUDA, "restoring tot_u: %s", cudaGetErrorString(e)));
  }
  for (Slab &s : h->slabs) s.window_first = hd.steps_done;
  *out = h;
  return LBM_OK;
}
}  // namespace
extern "C" {

// (the staging buffers of save / load are hundreds of MB: an allocation failure must come back as a code, not as an exception)
int lbm_save(lbm_handle *h, const char *path) {
  try {
    return save_solver(h, path);
  } catch (const std::exception &e) {
    return fail(LBM_E_NOMEM, "lbm_save: %s", e.what());
  }
}

int lbm_load(lbm_handle **out, const char *path, int num_gpus) {
  try {
    return load_solver(out, path, num_gpus, 0, 0, nullptr);
  } catch (const std::exception &e) {
    return fail(LBM_E_NOMEM, "lbm_load: %s", e.what());
  }
}

int lbm_load_grid(lbm_handle **out, const char *path, int px, int py, const int *devices) {
  if (px < 1 || py < 1) return fail(LBM_E_ARG, "the block grid must be at least 1 x 1 (got %d x %d)", px, py);
  try {
    return load_solver(out, path, 0, px, py, devices);
  } catch (const std::exception &e) {
    return fail(LBM_E_NOMEM, "lbm_load_grid: %s", e.what());
  }
}

int lbm_problem(lbm_handle *h, int *nx, int *ny, int *max_iters, float *density, float *accel, float *omega) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (nx) *nx = h->nx;
  if (ny) *ny = h->ny;
  if (max_iters) *max_iters = h->max_iters;
  if (density) *density = h->density;
  if (accel) *accel = h->accel;
  if (omega) *omega = h->omega;
  return LBM_OK;
}

// ---- one whole job from host buffers, pipelined -------------------------------------------------------
// lbm_upload -> lbm_run -> lbm_download is strictly sequential: PCIe carries 36 B per cell up, then nothing, then 36 B per
// cell down (2 x 9.66 GB at 16384^2, 0.36 s at 55 GB/s each way for 56 ms of arithmetic).  But information travels one row per
// timestep, so after K steps a band of rows depends only on itself plus K rows either side of the INITIAL state: the grid
// can be solved band by band -- upload a band with a margin of M >= K rows on both sides, run K steps on it (the band's own
// periodic wrap feeds garbage into the margins, which never reaches the band), download the band -- and the bands' uploads,
// sweeps and downloads overlap, which keeps both directions of the link busy at once.  Worker threads each drive their
// own band solver (an ordinary single-slab handle, created once per pipeline); one band uploads at a time and one downloads
// at a time, so each direction runs at its full rate.  The lattice is bit-identical to the sequential job; tot_u is summed
// per band over the band's own tile rows only (M + 1 is a multiple of the tile height) and the bands are added in band order.
}  // extern "C"

struct lbm_pipeline {
  int nx = 0, ny = 0, iters = 0, device = 0;
  // the rows this pipeline solves: all of them, or a WINDOW [row0, row0 + rows) of the global grid (one rank's slab).  A
  // window's input buffers start `margin` rows below row0 and end `margin` rows above its last row: the caller supplies the
  // initial state of those rows too (they belong to its neighbours), and no rank ever waits for another.
  bool window = false;
  int row0 = 0, rows = 0;
  float density = 0, accel = 0, omega = 0;
  bool banded = false;
  int margin = 0, band_rows = 0, nbands = 1, workers = 1;
  std::vector<lbm_handle *> full, last;  // per worker: a solver for full bands, and (if ny % band_rows) one for the shorter last band
  lbm_handle *whole = nullptr;           // not banded: the sequential job on one handle
  std::mutex up, down;                   // one band uploads / downloads at a time
  double last_ms = 0.0;
};

namespace {
constexpr int kSolveWorkers = 3;
constexpr long long kSolveBandCells = 16ll << 20;  // cells per band (without margins): ~0.6 GB of lattice each way

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct BandJob {
  const float *cells_in;
  const uint8_t *obstacles;
  float *cells_out;
  std::vector<std::vector<double>> tot;  // [band][step]
  std::vector<long long> free_cells;     // [band]
  double t_begin = 0;
};

int solve_band(lbm_pipeline *p, lbm_handle *h, BandJob &job, int b) {
  static const bool trace = getenv("LBM_B200_SOLVE_TRACE") != nullptr;
  const int nx = p->nx, ny = p->ny, M = p->margin;
  // band b = rows [r0, r0 + rows) of the pipeline's own rows (= of the grid unless it solves a window)
  const int r0 = b * p->band_rows, rows = std::min(p->band_rows, p->rows - r0), total = rows + 2 * M;
  Slab &s = h->slabs[0];
  // the band's rows -M .. rows + M - 1 are global rows row0 + r0 - M ... (periodic); the driven row may or may not be among them
  const int first = ((p->row0 + r0 - M) % ny + ny) % ny;
  const int arow = ((ny - 2 - first) % ny + ny) % ny;
  h->accel_row_global = (ny >= 2 && arow < total) ? arow : -1;
  s.tot_ty_lo = (M + 1) / lbm::kTileH;
  s.tot_ty_hi = s.tot_ty_lo + rows / lbm::kTileH;
  s.count_lo = M + 1;
  s.count_hi = M + 1 + rows;
  double t_a, t_b, t_c, t_d;
  int rc;
  {
    std::lock_guard<std::mutex> lock(p->up);
    t_a = now_ms();
    rc = reset_counters(h);
    if (p->window) {  // the caller's window is contiguous and starts M rows below the pipeline's first row
      rc = rc ? rc : upload_rows(h, 0, total, job.cells_in + (size_t)r0 * nx * lbm::kSpeeds, job.obstacles + (size_t)r0 * nx);
    } else {
      for (int done = 0; done < total && rc == LBM_OK;) {  // contiguous pieces of the periodic row range
        const int g = (first + done) % ny, n = std::min(total - done, ny - g);
        rc = upload_rows(h, done, n, job.cells_in + (size_t)g * nx * lbm::kSpeeds, job.obstacles + (size_t)g * nx);
        done += n;
      }
    }
    if (rc == LBM_OK) rc = drain_all(h);
    if (rc == LBM_OK) rc = finish_upload(h);
    t_b = now_ms();
  }
  if (rc == LBM_OK) rc = lbm_run(h, p->iters);
  if (rc) return rc;
  t_c = now_ms();
  job.free_cells[b] = s.free_cells;
  {
    std::lock_guard<std::mutex> lock(p->down);
    t_d = now_ms();
    rc = download_rows(h, M, rows, job.cells_out + (size_t)r0 * nx * lbm::kSpeeds, nullptr);
    if (rc == LBM_OK) rc = drain_all(h);
  }
  if (rc) return rc;
  job.tot[b].assign((size_t)p->iters, 0.0);
  if (p->iters) CUDA_TRY(cudaMemcpy(job.tot[b].data(), s.tot_u, (size_t)p->iters * sizeof(double), cudaMemcpyDeviceToHost));
  if (trace)
    fprintf(stderr, "band %2d: upload %.1f..%.1f  run ..%.1f (device %.1f)  download %.1f..%.1f ms\n", b, t_a - job.t_begin, t_b - job.t_begin,
            t_c - job.t_begin, h->last_ms, t_d - job.t_begin, now_ms() - job.t_begin);
  return LBM_OK;
}
}  // namespace
extern "C" {

void lbm_pipeline_destroy(lbm_pipeline *p) {
  if (!p) return;
  for (lbm_handle *h : p->full) lbm_destroy(h);
  for (lbm_handle *h : p->last) lbm_destroy(h);
  lbm_destroy(p->whole);
  delete p;
}

}  // extern "C"
namespace {

// How a job of `iters` timesteps on `rows` rows (the whole grid, or one rank's window of it) is cut into bands.
// false: a window that cannot be cut at all (the grid is shorter than one tile row of band plus its two margins)
bool plan_bands(int nx, int ny, bool window, int rows, int iters, bool &banded, int &band_rows, int &margin) {
  // the margin: >= iters rows, and M + 1 a multiple of the tile height so that a band is whole tile rows of its solver
  margin = ((std::max(iters, 1) + 1 + lbm::kTileH - 1) / lbm::kTileH) * lbm::kTileH - 1;
  const int min_band = (8 * margin + lbm::kTileH - 1) / lbm::kTileH * lbm::kTileH;  // margins at most a quarter of the upload
  long long want = std::max<long long>(lbm::kTileH, kSolveBandCells / nx / lbm::kTileH * lbm::kTileH);
  if (window) want = std::min<long long>(want, std::max(lbm::kTileH, rows / 4 / lbm::kTileH * lbm::kTileH));  // a slab: at least four bands
  band_rows = (int)want;
  if (const char *e = getenv("LBM_B200_BAND_ROWS")) band_rows = std::max(lbm::kTileH, atoi(e) / lbm::kTileH * lbm::kTileH);  // tests
  else if (window) band_rows = std::min(rows, std::max(band_rows, min_band));
  // a band solver drives ONE row: band + margins must not hold the grid's driven row twice, i.e. never be taller than the grid
  // (whole-grid jobs cannot get there: they band only when ny >= 4 bands and a band >= 8 margins)
  if (window && band_rows + 2 * margin > ny) {
    band_rows = (ny - 2 * margin) / lbm::kTileH * lbm::kTileH;
    if (band_rows < lbm::kTileH) return false;
  }
  // bands pay when the margins are a small part of a band and there are enough bands to overlap; a window has no other way
  banded = window || (ny % lbm::kTileH == 0 && band_rows >= min_band && ny >= 4 * band_rows && !getenv("LBM_B200_NO_BANDS"));
  return true;
}

// window == false: the whole grid (row0 = 0, rows = ny)
int pipeline_create(lbm_pipeline **out, int nx, int ny, bool window, int row0, int rows, int iters, float density, float accel,
                    float omega, int device) {
  if (!out) return fail(LBM_E_ARG, "null argument");
  *out = nullptr;
  int rc = check_problem(nx, ny, iters, 1);
  if (rc) return rc;
  if (window && (row0 < 0 || rows < 1 || row0 + rows > ny)) return fail(LBM_E_ARG, "window [%d, %d) is not inside 0..%d", row0, row0 + rows, ny);
  if (window && rows % lbm::kTileH != 0) return fail(LBM_E_ARG, "a window must be a multiple of %d rows (got %d)", lbm::kTileH, rows);
  lbm_pipeline *p = new (std::nothrow) lbm_pipeline();
  if (!p) return fail(LBM_E_NOMEM, "out of host memory");
  p->nx = nx; p->ny = ny; p->iters = iters; p->device = device;
  p->density = density; p->accel = accel; p->omega = omega;
  p->window = window; p->row0 = row0; p->rows = rows;
  if (!plan_bands(nx, ny, window, rows, iters, p->banded, p->band_rows, p->margin)) {
    const int m = p->margin;
    lbm_pipeline_destroy(p);
    return fail(LBM_E_ARG, "a %d-row grid is too short for a window pipeline of %d timesteps (%d margin rows either side of a band)", ny, iters, m);
  }
  if (!p->banded) {
    rc = lbm_create_on_devices(&p->whole, nx, ny, iters, density, accel, omega, 1, &device);
    if (rc) { lbm_pipeline_destroy(p); return rc; }
    *out = p;
    return LBM_OK;
  }
  p->nbands = (rows + p->band_rows - 1) / p->band_rows;
  p->workers = std::min(kSolveWorkers, p->nbands);
  if (const char *e = getenv("LBM_B200_SOLVE_WORKERS")) p->workers = std::max(1, std::min(atoi(e), p->nbands));
  const int last_rows = rows - (p->nbands - 1) * p->band_rows;
  auto band_solver = [&](int band, lbm_handle **slot) {
    int r = lbm_create_on_devices(slot, nx, band + 2 * p->margin, iters, density, accel, omega, 1, &device);
    if (r == LBM_OK) {  // bands take the plain per-step kernel (a cached graph would bake in one band's driven row)
      (*slot)->slabs[0].persistent_grid = 0;
      (*slot)->slabs[0].cluster_ctas = 0;
      (*slot)->no_graphs = true;
    }
    return r;
  };
  p->full.assign(p->workers, nullptr);
  p->last.assign(p->workers, nullptr);
  for (int w = 0; w < p->workers && rc == LBM_OK; ++w) {
    rc = band_solver(p->band_rows, &p->full[w]);
    // the last band goes to worker (nbands - 1) % workers
    if (rc == LBM_OK && last_rows != p->band_rows && w == (p->nbands - 1) % p->workers) rc = band_solver(last_rows, &p->last[w]);
  }
  if (rc) { lbm_pipeline_destroy(p); return rc; }
  *out = p;
  return LBM_OK;
}

// all bands; tot[t] = the per-step sums of the bands added in band order, free_cells = non-obstacle cells of the pipeline's rows
int pipeline_run(lbm_pipeline *p, const float *cells_in, const uint8_t *obstacles, float *cells_out, std::vector<double> &tot,
                 long long &free_cells) {
  BandJob job{};
  job.cells_in = cells_in; job.obstacles = obstacles; job.cells_out = cells_out;
  job.tot.resize(p->nbands);
  job.free_cells.assign(p->nbands, 0);
  job.t_begin = now_ms();
  std::vector<int> codes(p->workers, LBM_OK);
  std::vector<std::string> messages(p->workers);
  std::vector<std::thread> threads;
  threads.reserve(p->workers);
  std::atomic<bool> stop{false};  // a worker failed, or a thread could not be started: the others finish their band and leave
  bool spawned = true;
  for (int w = 0; w < p->workers && spawned; ++w) {
    try {
      threads.emplace_back([&, w]() {
        try {
          for (int b = w; b < p->nbands && !stop.load(); b += p->workers) {
            const int rows = std::min(p->band_rows, p->rows - b * p->band_rows);
            lbm_handle *h = rows == p->band_rows ? p->full[w] : p->last[w];
            const int r = solve_band(p, h, job, b);
            if (r != LBM_OK) { codes[w] = r; messages[w] = g_last_error; stop.store(true); }
          }
        } catch (const std::exception &e) {  // nothing may escape a worker thread (it would terminate the process)
          codes[w] = LBM_E_NOMEM;
          messages[w] = e.what();
          stop.store(true);
        }
      });
    } catch (const std::exception &) {  // no more threads: stop the ones that run, then report
      spawned = false;
      stop.store(true);
    }
  }
  for (std::thread &t : threads) t.join();  // always, before anything can unwind past the vector of joinable threads
  if (!spawned) return fail(LBM_E_NOMEM, "could not start the pipeline's worker threads");
  for (int w = 0; w < p->workers; ++w)
    if (codes[w] != LBM_OK) return fail(codes[w], "band solver: %s", messages[w].c_str());
  free_cells = 0;
  for (long long f : job.free_cells) free_cells += f;
  tot.assign((size_t)p->iters, 0.0);
  for (int t = 0; t < p->iters; ++t)
    for (int b = 0; b < p->nbands; ++b) tot[t] += job.tot[b][t];  // band order
  return LBM_OK;
}

}  // namespace
extern "C" {

int lbm_pipeline_create(lbm_pipeline **out, int nx, int ny, int iters, float density, float accel, float omega, int device) {
  return pipeline_create(out, nx, ny, false, 0, ny, iters, density, accel, omega, device);
}

int lbm_pipeline_create_window(lbm_pipeline **out, int nx, int ny, int row_begin, int row_count, int iters, float density,
                               float accel, float omega, int device) {
  return pipeline_create(out, nx, ny, true, row_begin, row_count, iters, density, accel, omega, device);
}

int lbm_pipeline_solve(lbm_pipeline *p, const float *cells_in, const uint8_t *obstacles, float *cells_out, float *av_vels) {
  if (!p || !cells_in || !obstacles || !cells_out) return fail(LBM_E_ARG, "null argument");
  if (p->window) return fail(LBM_E_STATE, "a window pipeline returns partial sums: use lbm_pipeline_solve_window");
  if (cells_in == cells_out) return fail(LBM_E_ARG, "cells_out must not alias cells_in");
  const double t_begin = now_ms();
  int rc = LBM_OK;
  if (!p->banded) {
    rc = lbm_upload(p->whole, cells_in, obstacles);
    if (rc == LBM_OK) rc = lbm_run(p->whole, p->iters);
    if (rc == LBM_OK) rc = lbm_download(p->whole, cells_out, av_vels);
    p->last_ms = now_ms() - t_begin;
    return rc;
  }
  try {
    std::vector<double> tot;
    long long free_cells = 0;
    rc = pipeline_run(p, cells_in, obstacles, cells_out, tot, free_cells);
    if (rc) return rc;
    if (av_vels)
      for (int t = 0; t < p->iters; ++t) av_vels[t] = (float)tot[t] / (float)free_cells;  // main/LastChance.cpp:266
  } catch (const std::exception &e) {  // thread or host-memory exhaustion: nothing may throw across the ABI
    return fail(LBM_E_NOMEM, "lbm_pipeline_solve: %s", e.what());
  }
  p->last_ms = now_ms() - t_begin;
  return LBM_OK;
}

int lbm_pipeline_solve_window(lbm_pipeline *p, const float *cells_window, const uint8_t *obstacles_window, float *cells_out,
                              double *tot_u, long long *free_cells) {
  if (!p || !cells_window || !obstacles_window || !cells_out) return fail(LBM_E_ARG, "null argument");
  if (!p->window) return fail(LBM_E_STATE, "not a window pipeline (lbm_pipeline_create_window)");
  const double t_begin = now_ms();
  try {
    std::vector<double> tot;
    long long free_count = 0;
    int rc = pipeline_run(p, cells_window, obstacles_window, cells_out, tot, free_count);
    if (rc) return rc;
    if (tot_u)
      for (int t = 0; t < p->iters; ++t) tot_u[t] = tot[t];
    if (free_cells) *free_cells = free_count;
  } catch (const std::exception &e) {
    return fail(LBM_E_NOMEM, "lbm_pipeline_solve_window: %s", e.what());
  }
  p->last_ms = now_ms() - t_begin;
  return LBM_OK;
}

int lbm_pipeline_plan(int nx, int ny, int window_rows, int iters, int *bands, int *band_rows, int *margin_rows) {
  const bool window = window_rows > 0;
  const int rows = window ? window_rows : ny;
  int rc = check_problem(nx, ny, iters, 1);
  if (rc) return rc;
  if (window && (rows > ny || rows % lbm::kTileH != 0)) return fail(LBM_E_ARG, "a window must be a multiple of %d rows inside the grid", lbm::kTileH);
  bool banded = false;
  int band = 0, margin = 0;
  if (!plan_bands(nx, ny, window, rows, iters, banded, band, margin))
    return fail(LBM_E_ARG, "a %d-row grid is too short for a window pipeline of %d timesteps (%d margin rows either side of a band)", ny, iters, margin);
  if (bands) *bands = banded ? (rows + band - 1) / band : 1;
  if (band_rows) *band_rows = banded ? band : ny;
  if (margin_rows) *margin_rows = banded ? margin : 0;
  return LBM_OK;
}

int lbm_pipeline_info(lbm_pipeline *p, int *bands, int *band_rows, int *margin_rows, double *last_ms) {
  if (!p) return fail(LBM_E_ARG, "null argument");
  if (bands) *bands = p->banded ? p->nbands : 1;
  if (band_rows) *band_rows = p->banded ? p->band_rows : p->ny;  // (window pipelines are always banded)
  if (margin_rows) *margin_rows = p->banded ? p->margin : 0;
  if (last_ms) *last_ms = p->last_ms;
  return LBM_OK;
}

int lbm_solve_host(int nx, int ny, int iters, float density, float accel, float omega, int device, const float *cells_in,
                   const uint8_t *obstacles, float *cells_out, float *av_vels, double *elapsed_ms) {
  if (!cells_in || !obstacles || !cells_out) return fail(LBM_E_ARG, "null argument");
  if (cells_in == cells_out) return fail(LBM_E_ARG, "cells_out must not alias cells_in");
  const double t_begin = now_ms();
  lbm_pipeline *p = nullptr;
  int rc = lbm_pipeline_create(&p, nx, ny, iters, density, accel, omega, device);
  if (rc) return rc;
  rc = lbm_pipeline_solve(p, cells_in, obstacles, cells_out, av_vels);
  lbm_pipeline_destroy(p);
  if (elapsed_ms) *elapsed_ms = now_ms() - t_begin;
  return rc;
}

// ---- introspection -------------------------------------------------------------------------------
int lbm_steps_done(lbm_handle *h, int *steps) {
  if (!h || !steps) return fail(LBM_E_ARG, "null argument");
  *steps = h->steps_done;
  return LBM_OK;
}
int lbm_free_cells(lbm_handle *h, long long *count) {
  if (!h || !count) return fail(LBM_E_ARG, "null argument");
  if (!h->uploaded) return fail(LBM_E_STATE, "nothing uploaded");
  *count = 0;
  for (Slab &s : h->slabs) *count += s.free_cells;
  return LBM_OK;
}
int lbm_launch_count(lbm_handle *h, long long *count) {
  if (!h || !count) return fail(LBM_E_ARG, "null argument");
  *count = h->launches;
  return LBM_OK;
}
const char *lbm_kernel_variant(lbm_handle *h) {
  if (!h || h->slabs.empty()) return "";
  if (h->slabs[0].cluster_ctas > 0) return "vec4-cluster";
  switch (h->slabs[0].kernel) {
    case KERNEL_VEC4: return h->slabs[0].persistent_grid > 0 ? "vec4-persistent" : "vec4";
    case KERNEL_HYBRID: return h->slabs[0].persistent_grid > 0 ? "hybrid-persistent" : "hybrid";
    default: return h->slabs[0].persistent_grid > 0 ? "scalar-persistent" : "scalar";
  }
}
int lbm_rows(lbm_handle *h, int *row_begin, int *row_count) {
  if (!h || !row_begin || !row_count) return fail(LBM_E_ARG, "null argument");
  if (h->rank_mode) {
    *row_begin = h->slabs[0].row0;
    *row_count = h->slabs[0].nyl;
  } else {
    *row_begin = 0;
    *row_count = h->ny;
  }
  return LBM_OK;
}

int lbm_slab_count(lbm_handle *h, int *count) {
  if (!h || !count) return fail(LBM_E_ARG, "null argument");
  *count = (int)h->slabs.size();
  return LBM_OK;
}
int lbm_slab_info(lbm_handle *h, int slab, int *device, int *row_begin, int *row_count, long long *free_cells) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (slab < 0 || slab >= (int)h->slabs.size()) return fail(LBM_E_ARG, "slab %d of %zu", slab, h->slabs.size());
  const Slab &s = h->slabs[slab];
  if (device) *device = s.device;
  if (row_begin) *row_begin = s.row0;
  if (row_count) *row_count = s.nyl;
  if (free_cells) *free_cells = h->uploaded ? s.free_cells : -1;
  return LBM_OK;
}

int lbm_block_info(lbm_handle *h, int block, int *device, int *row_begin, int *row_count, int *col_begin, int *col_count,
                   long long *free_cells) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (block < 0 || block >= (int)h->slabs.size()) return fail(LBM_E_ARG, "block %d of %zu", block, h->slabs.size());
  const Slab &s = h->slabs[block];
  if (device) *device = s.device;
  if (row_begin) *row_begin = s.row0;
  if (row_count) *row_count = s.nyl;
  if (col_begin) *col_begin = s.col0;
  if (col_count) *col_count = s.nxl;
  if (free_cells) *free_cells = h->uploaded ? s.free_cells : -1;
  return LBM_OK;
}
int lbm_partition(lbm_handle *h, int *px, int *py) {
  if (!h) return fail(LBM_E_ARG, "null argument");
  if (px) *px = h->px;
  if (py) *py = h->py;
  return LBM_OK;
}

int lbm_selftest_math(int device, unsigned long long pairs, unsigned long long *counts) {
  if (!counts) return fail(LBM_E_ARG, "null argument");
  int ndev = 0;
  int rc = lbm_device_count(&ndev);
  if (rc) return rc;
  if (device < 0 || device >= ndev) return fail(LBM_E_ARG, "GPU %d not visible (%d devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  unsigned long long *d = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d, 7 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(d, 0, 7 * sizeof(unsigned long long)));
  lbm::selftest_math_kernel<<<148 * 8, 256>>>(d, pairs, 0x5EEDull);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(counts, d, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(LBM_E_CUDA, "math self-test failed: %s", cudaGetErrorString(e));
  return LBM_OK;
}

}  // extern "C"
namespace {
// NUMA node and CPU list of the host bridge a GPU hangs off, from sysfs (-1 / "" when unknown)
int gpu_numa_node(int device, std::string *cpulist) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
  const std::string dir = std::string("/sys/bus/pci/devices/") + bus;
  int node = -1;
  if (FILE *f = fopen((dir + "/numa_node").c_str(), "r")) {
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
  }
  if (cpulist) {
    cpulist->clear();
    if (FILE *f = fopen((dir + "/local_cpulist").c_str(), "r")) {
      char buf[4096] = {0};
      if (fgets(buf, sizeof buf, f)) *cpulist = buf;
      fclose(f);
    }
  }
  return node;
}
}  // namespace
extern "C" {

// Pinned host memory placed on the NUMA node of `device`'s PCIe root (best effort: the memory policy of the calling
// thread is switched to "prefer that node" for the duration of the allocation; containers that forbid set_mempolicy
// get a plain pinned allocation).  On a two-socket 8-GPU board uploads from the far socket cross the inter-socket link.
void *lbm_host_alloc_near(size_t bytes, int device) {
  const int node = gpu_numa_node(device, nullptr);
  bool policy_set = false;
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024 + 1) == 0;
  }
  void *p = lbm_host_alloc(bytes);
  if (policy_set) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
  return p;
}

// Restrict the calling thread to the CPUs local to `device` (sysfs local_cpulist).  Returns the NUMA node, or a
// negative code when the topology is unknown or the affinity call is refused; purely an optimisation.
int lbm_bind_thread_near(int device) {
  std::string cpus;
  const int node = gpu_numa_node(device, &cpus);
  if (cpus.empty()) return fail(LBM_E_STATE, "no local_cpulist for GPU %d in sysfs", device);
  cpu_set_t set;
  CPU_ZERO(&set);
  int count = 0;
  const char *c = cpus.c_str();
  while (*c) {  // "0-31,64-95"
    char *end = nullptr;
    long a = strtol(c, &end, 10), b = a;
    if (end == c) break;
    if (*end == '-') b = strtol(end + 1, &end, 10);
    for (long k = a; k <= b && k < CPU_SETSIZE; ++k) { CPU_SET((int)k, &set); ++count; }
    c = (*end == ',') ? end + 1 : end;
    if (*end != ',') break;
  }
  if (!count) return fail(LBM_E_STATE, "empty local_cpulist for GPU %d", device);
  if (sched_setaffinity(0, sizeof set, &set) != 0) return fail(LBM_E_STATE, "sched_setaffinity refused for GPU %d's CPUs", device);
  return node >= 0 ? node : 0;
}

void *lbm_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    fail(LBM_E_NOMEM, "cudaMallocHost(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void lbm_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"

// pcie_probe.cu -- the host<->device copy ceiling that bench.py's e2e number lives under.
//
//   pcie_probe [bytes_per_gpu=2147483648] [reps=3]
//
// For N = 1, 2, 4, 8 (as many GPUs as are visible) N host threads, one per GPU, copy `bytes` of pinned host memory to
// their GPU (H2D), back (D2H) and both at once (two streams), all N at the same time, in 64 MB cudaMemcpyAsync pieces
// -- the transfer pattern of lbm_upload / lbm_download.  Each case runs twice: buffers allocated wherever the OS puts
// them ("default"), and allocated on the GPU's own NUMA node with the thread bound to the GPU-local CPUs ("near").
// Output: one line per case with the per-GPU min/max and the aggregate GB/s.  Also prints each GPU's bus id and NUMA node.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static int numa_node_of(int dev, std::string *cpus) {
  char bus[32] = {0};
  CK(cudaDeviceGetPCIBusId(bus, sizeof bus, dev));
  for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
  std::string dir = std::string("/sys/bus/pci/devices/") + bus;
  int node = -1;
  if (FILE *f = fopen((dir + "/numa_node").c_str(), "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
  if (cpus) {
    cpus->clear();
    if (FILE *f = fopen((dir + "/local_cpulist").c_str(), "r")) { char b[4096] = {0}; if (fgets(b, sizeof b, f)) *cpus = b; fclose(f); }
  }
  return node;
}
static bool bind_near(int dev) {
  std::string cpus;
  numa_node_of(dev, &cpus);
  cpu_set_t set; CPU_ZERO(&set); int n = 0;
  const char *c = cpus.c_str();
  while (*c) {
    char *end = nullptr; long a = strtol(c, &end, 10), b = a;
    if (end == c) break;
    if (*end == '-') b = strtol(end + 1, &end, 10);
    for (long k = a; k <= b && k < CPU_SETSIZE; ++k) { CPU_SET((int)k, &set); ++n; }
    if (*end != ',') break;
    c = end + 1;
  }
  return n && sched_setaffinity(0, sizeof set, &set) == 0;
}
static bool prefer_node(int node) {
  if (node < 0) return false;
  unsigned long mask[16] = {0};
  mask[node / 64] |= 1ul << (node % 64);
  return syscall(SYS_set_mempolicy, 1, mask, 1025) == 0;
}

struct Barrier {
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    const int g = gen.load();
    if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
    else while (gen.load() == g) std::this_thread::yield();
  }
};

int main(int argc, char **argv) {
  const size_t bytes = argc > 1 ? strtoull(argv[1], nullptr, 10) : (2ull << 30);
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  const size_t piece = 64ull << 20;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  printf("pcie_probe: %d GPU(s), %zu bytes per GPU per direction, %d reps, %zu-byte pieces, %ld online CPUs\n", ndev, bytes, reps, piece,
         sysconf(_SC_NPROCESSORS_ONLN));
  for (int d = 0; d < ndev; ++d) {
    std::string cpus; char bus[32];
    CK(cudaDeviceGetPCIBusId(bus, sizeof bus, d));
    const int node = numa_node_of(d, &cpus);
    printf("  GPU %d  bus %s  numa_node %d  local_cpulist %s", d, bus, node, cpus.empty() ? "?\n" : cpus.c_str());
  }
  for (int near = 0; near < 2; ++near) {
    for (int n = 1; n <= ndev; n *= 2) {
      std::vector<double> gbs[3];
      for (auto &v : gbs) v.assign(n, 0.0);
      std::vector<int> placed(n, 0);
      Barrier bar(n);
      std::vector<std::thread> th;
      for (int d = 0; d < n; ++d) th.emplace_back([&, d]() {
        CK(cudaSetDevice(d));
        bool ok = true;
        if (near) { ok = bind_near(d); ok = prefer_node(numa_node_of(d, nullptr)) && ok; }
        placed[d] = ok;
        char *h[2], *g[2];
        cudaStream_t st[2];
        for (int i = 0; i < 2; ++i) { CK(cudaMallocHost((void **)&h[i], bytes)); CK(cudaMalloc((void **)&g[i], bytes)); CK(cudaStreamCreate(&st[i])); }
        if (near) syscall(SYS_set_mempolicy, 0, nullptr, 0);
        for (int mode = 0; mode < 3; ++mode) {  // 0 H2D, 1 D2H, 2 both at once
          double best = 0;
          for (int r = 0; r <= reps; ++r) {
            bar.wait();
            const auto t0 = std::chrono::steady_clock::now();
            for (size_t o = 0; o < bytes; o += piece) {
              const size_t m = bytes - o < piece ? bytes - o : piece;
              if (mode != 1) CK(cudaMemcpyAsync(g[0] + o, h[0] + o, m, cudaMemcpyHostToDevice, st[0]));
              if (mode != 0) CK(cudaMemcpyAsync(h[1] + o, g[1] + o, m, cudaMemcpyDeviceToHost, st[1]));
            }
            CK(cudaStreamSynchronize(st[0]));
            CK(cudaStreamSynchronize(st[1]));
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const double v = (mode == 2 ? 2.0 : 1.0) * bytes / dt / 1e9;
            if (r > 0 && v > best) best = v;  // rep 0 warms up
            bar.wait();
          }
          gbs[mode][d] = best;
        }
        for (int i = 0; i < 2; ++i) { cudaFreeHost(h[i]); cudaFree(g[i]); cudaStreamDestroy(st[i]); }
      });
      for (auto &t : th) t.join();
      const char *names[3] = {"H2D", "D2H", "H2D+D2H"};
      for (int mode = 0; mode < 3; ++mode) {
        double lo = 1e30, hi = 0, sum = 0;
        for (double v : gbs[mode]) { lo = v < lo ? v : lo; hi = v > hi ? v : hi; sum += v; }
        int okc = 0; for (int v : placed) okc += v;
        printf("placement=%-7s gpus=%d  %-8s per-GPU min %6.1f max %6.1f  aggregate %7.1f GB/s%s\n", near ? "near" : "default", n, names[mode], lo, hi,
               sum, near ? (okc == n ? "" : "  (NUMA binding refused on some threads)") : "");
      }
      fflush(stdout);
    }
  }
  return 0;
}

// stream_probe.cu -- development tool: what HBM bandwidth can the sweep's ACCESS PATTERN reach with no arithmetic?
// 9 source planes -> 9 destination planes (structure of arrays, row pitch, planes 2,5,6 read one row south and
// 4,7,8 one row north like the pull-propagate), through the data paths the sweep kernel could use:
//   copy      1 -> 1 plane, float4 grid-stride (the MEASURED_PEAKS.json-style ceiling)
//   ldg       9 -> 9, LDG.128 / STG.128 from registers, persistent row-marching tiles
//   tma       9 narrow TMA boxes -> smem stage ring -> LDS.128 -> STG.128; producer = thread 0 after a CTA barrier
//   tmaws     same, warp-specialised: a producer warp + empty/full mbarriers, consumers never meet at a barrier
//   tmast     TMA load -> smem -> TMA store (no register path at all)
// Usage: stream_probe <mode> [nx ny W H stages blocks_per_sm iters]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kCol0 = 32;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(float *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const float *src, const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Args {
  const float *src; float *dst; size_t plane; int pitch, nx, ny, stages, rows_iter;
};
__device__ __forceinline__ int row_shift(int k) { return (k == 2 || k == 5 || k == 6) ? -1 : (k == 4 || k == 7 || k == 8) ? 1 : 0; }

__global__ void copy_kernel(const float4 *a, float4 *b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// 4 independent float4 per thread in flight, non-persistent grid (how library elementwise kernels are shaped)
__global__ void copy4_kernel(const float4 *a, float4 *b, size_t n) {
  const size_t i0 = ((size_t)blockIdx.x * blockDim.x * 4) + threadIdx.x;
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) if (i0 + j * blockDim.x < n) v[j] = a[i0 + j * blockDim.x];
#pragma unroll
  for (int j = 0; j < 4; ++j) if (i0 + j * blockDim.x < n) b[i0 + j * blockDim.x] = v[j];
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) ldg_kernel(Args a) {
  // blockDim (TX, TY); tile TX*4 wide, TY*rows_iter tall
  const int TX = blockDim.x, TY = blockDim.y, tile_w = TX * 4, tile_h = TY * a.rows_iter;
  const int tiles_x = a.nx / tile_w, tiles_y = (a.ny + tile_h - 1) / tile_h;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x = (tx * TX + threadIdx.x) * 4;
    for (int row = 1 + ty * tile_h + threadIdx.y; row < min(1 + (ty + 1) * tile_h, a.ny + 1); row += TY) {
      float4 v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = *reinterpret_cast<const float4 *>(a.src + k * a.plane + (size_t)(row + row_shift(k)) * a.pitch + kCol0 + x);
#pragma unroll
      for (int k = 0; k < 9; ++k) *reinterpret_cast<float4 *>(a.dst + k * a.plane + (size_t)row * a.pitch + kCol0 + x) = v[k];
    }
  }
}

// mode: 0 = producer thread 0 + CTA barrier per unit; 1 = warp-specialised; 2 = TMA store
template <int W, int H, int MODE>
__global__ void __launch_bounds__(32 * (H + 1), 1) tma_kernel(Args a, const __grid_constant__ CUtensorMap map_src,
                                                            const __grid_constant__ CUtensorMap map_dst) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  float *stage_base = reinterpret_cast<float *>(smem_raw);
  constexpr int kPlaneFloats = W * H, kStageFloats = 9 * W * H;
  constexpr int VPL = W / 32;  // floats per lane per row (4 or 8)
  const int S = a.stages, G = gridDim.x;
  const int tiles_x = a.nx / W, tiles_y = a.ny / H, nunits = tiles_x * tiles_y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mine = ((int)blockIdx.x < nunits) ? (nunits - 1 - (int)blockIdx.x) / G + 1 : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], H); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int i) {
    const int u = (int)blockIdx.x + i * G, ty = u / tiles_x, tx = u - ty * tiles_x;
    float *dst = stage_base + (size_t)(i % S) * kStageFloats;
    uint64_t *bar = &full_bar[i % S];
    mbar_expect_tx(bar, kStageFloats * 4);
#pragma unroll
    for (int k = 0; k < 9; ++k) tma_load_3d(dst + k * kPlaneFloats, &map_src, bar, kCol0 + tx * W, 1 + ty * H + row_shift(k), k);
  };
  if (MODE == 2) {
    // single elected thread moves everything: load ring -> store from the same stage
    if (threadIdx.x == 0) {
      for (int i = 0; i < S && i < n_mine; ++i) issue(i);
      for (int i = 0; i < n_mine; ++i) {
        const int u = (int)blockIdx.x + i * G, ty = u / tiles_x, tx = u - ty * tiles_x;
        mbar_wait(&full_bar[i % S], (i / S) & 1);
        const float *sp = stage_base + (size_t)(i % S) * kStageFloats;
#pragma unroll
        for (int k = 0; k < 9; ++k) tma_store_3d(sp + k * kPlaneFloats, &map_dst, kCol0 + tx * W, 1 + ty * H, k);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the stage that was stored S-1 units ago... simplest: wait until this one has been read, then refill it
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (i + S < n_mine) issue(i + S);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    return;
  }
  if (MODE == 1 && warp == H) {  // producer warp
    if (lane == 0) {
      for (int i = 0; i < n_mine; ++i) {
        if (i >= S) mbar_wait(&empty_bar[i % S], ((i / S) - 1) & 1);
        issue(i);
      }
    }
    return;
  }
  if (MODE == 0) {
    if (warp == H) return;  // unused warp in this mode
    if (threadIdx.x == 0) for (int i = 0; i < S && i < n_mine; ++i) issue(i);
  }
  for (int i = 0; i < n_mine; ++i) {
    const int u = (int)blockIdx.x + i * G, ty = u / tiles_x, tx = u - ty * tiles_x;
    const int slot = i % S;
    mbar_wait(&full_bar[slot], (i / S) & 1);
    const float *sp = stage_base + (size_t)slot * kStageFloats + warp * W + lane * 4;
    float4 v[9][VPL / 4];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int j = 0; j < VPL / 4; ++j) v[k][j] = *reinterpret_cast<const float4 *>(sp + k * kPlaneFloats + j * 128);
    if (MODE == 0) {
      asm volatile("bar.sync 1, %0;" ::"r"(32 * H) : "memory");
      if (threadIdx.x == 0 && i + S < n_mine) issue(i + S);
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
    float *out = a.dst + (size_t)(1 + ty * H + warp) * a.pitch + kCol0 + tx * W + lane * 4;
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int j = 0; j < VPL / 4; ++j) *reinterpret_cast<float4 *>(out + k * a.plane + j * 128) = v[k][j];
  }
}

// Blocked layouts, moved with 1-D bulk copies.  A unit is 9 planes x H rows x W cells = UNIT floats.
//   LAYOUT 0 "tile":     the whole unit is one contiguous blob; units follow each other in memory
//   LAYOUT 1 "rowchunk": per row, the 9 planes' W-cell chunks are contiguous (9*W floats); the H rows of a unit are
//                        row_stride apart (row_stride = tiles_x * 9 * W floats), consecutive units in x are adjacent
__device__ __forceinline__ void bulk_load(float *dst, const float *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int W, int H, int LAYOUT>
__global__ void __launch_bounds__(288, 1) blocked_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  float *stage_base = reinterpret_cast<float *>(smem_raw);
  constexpr int UNIT = 9 * W * H, ROWF = 9 * W;
  const int S = a.stages, G = gridDim.x;
  const int tiles_x = a.nx / W, tiles_y = a.ny / H, nunits = tiles_x * tiles_y;
  const size_t row_stride = (size_t)tiles_x * ROWF;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mine = ((int)blockIdx.x < nunits) ? (nunits - 1 - (int)blockIdx.x) / G + 1 : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < n_mine; ++i) {
        if (i >= S) mbar_wait(&empty_bar[i % S], ((i / S) - 1) & 1);
        const int u = (int)blockIdx.x + i * G, ty = u / tiles_x, tx = u - ty * tiles_x;
        float *dst = stage_base + (size_t)(i % S) * UNIT;
        uint64_t *bar = &full_bar[i % S];
        mbar_expect_tx(bar, UNIT * 4);
        if (LAYOUT == 0) {
          const float *src = a.src + (size_t)u * UNIT;
          for (int c = 0; c < 9; ++c) bulk_load(dst + c * (UNIT / 9), src + c * (UNIT / 9), UNIT / 9 * 4, bar);
        } else {
          const float *src = a.src + (size_t)ty * H * row_stride + (size_t)tx * ROWF;
          for (int r = 0; r < H; ++r) bulk_load(dst + r * ROWF, src + r * row_stride, ROWF * 4, bar);
        }
      }
    }
    return;
  }
  for (int i = 0; i < n_mine; ++i) {
    const int u = (int)blockIdx.x + i * G, ty = u / tiles_x, tx = u - ty * tiles_x;
    const int slot = i % S;
    mbar_wait(&full_bar[slot], (i / S) & 1);
    const float *sp = stage_base + (size_t)slot * UNIT;
    constexpr int PER = UNIT / 4 / 256;  // float4 per thread
    float4 v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = *reinterpret_cast<const float4 *>(sp + (j * 256 + threadIdx.x) * 4);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int f = (j * 256 + threadIdx.x) * 4;  // float index inside the unit
      float *out;
      if (LAYOUT == 0) out = a.dst + (size_t)u * UNIT + f;
      else out = a.dst + (size_t)ty * H * row_stride + (size_t)tx * ROWF + (size_t)(f / ROWF) * row_stride + (f % ROWF);
      *reinterpret_cast<float4 *>(out) = v[j];
    }
  }
}
template <int W, int H, int LAYOUT>
static void launch_blocked(const Args &a, int grid, cudaStream_t st) {
  const size_t smem = (size_t)a.stages * 9 * W * H * 4;
  static bool set = false;
  if (!set) { CK(cudaFuncSetAttribute(blocked_kernel<W, H, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
  blocked_kernel<W, H, LAYOUT><<<grid, 288, smem, st>>>(a);
}

// Non-persistent, tile layout: one CTA per unit blob.  VIA 0: one bulk copy -> smem -> LDS -> STG;
// VIA 1: LDG.128 -> STG.128 from registers; VIA 2: bulk copy in, bulk copy out (smem -> global)
template <int W, int H, int VIA>
__global__ void __launch_bounds__(32 * H) blob_kernel(Args a) {
  constexpr int UNIT = 9 * W * H, T = 32 * H, PER = UNIT / 4 / T;
  // every float4 of the blob must be moved: W=64 gave PER = 4.5 -> 4 and silently skipped 1/9 of the bytes, which is
  // what produced the 7.33 / 7.72 TB/s lines of profiles/r01c_stream_probe_4.txt (struck there)
  static_assert(PER * 4 * T == UNIT, "blob_kernel: the blob is not a whole number of float4 per thread");
  __shared__ __align__(128) float stage[VIA == 1 ? 4 : UNIT];
  __shared__ __align__(8) uint64_t bar;
  const size_t u = blockIdx.x;
  const float *src = a.src + u * UNIT;
  float *dst = a.dst + u * UNIT;
  if (VIA == 1) {
    float4 v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = *reinterpret_cast<const float4 *>(src + (j * T + threadIdx.x) * 4);
#pragma unroll
    for (int j = 0; j < PER; ++j) *reinterpret_cast<float4 *>(dst + (j * T + threadIdx.x) * 4) = v[j];
    return;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar, UNIT * 4);
    bulk_load(stage, src, UNIT * 4, &bar);
  }
  __syncthreads();
  if (VIA == 2) {
    if (threadIdx.x == 0) {
      mbar_wait(&bar, 0);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stage)), "r"(UNIT * 4) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  mbar_wait(&bar, 0);
  float4 v[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) v[j] = *reinterpret_cast<const float4 *>(stage + (j * T + threadIdx.x) * 4);
#pragma unroll
  for (int j = 0; j < PER; ++j) *reinterpret_cast<float4 *>(dst + (j * T + threadIdx.x) * 4) = v[j];
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(float *buf, int pitch, int rows, size_t plane, int W, int H) {
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, 9};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)plane * 4};
  const cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)H, 1}, es[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int W, int H, int MODE>
static void launch_tma(const Args &a, const CUtensorMap &ms, const CUtensorMap &md, int grid, cudaStream_t st) {
  const size_t smem = (size_t)a.stages * 9 * W * H * 4;
  static bool set = false;
  if (!set) { CK(cudaFuncSetAttribute(tma_kernel<W, H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
  tma_kernel<W, H, MODE><<<grid, 32 * (H + 1), smem, st>>>(a, ms, md);
}

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "copy";
  const int nx = argc > 2 ? atoi(argv[2]) : 16384, ny = argc > 3 ? atoi(argv[3]) : 16384;
  const int W = argc > 4 ? atoi(argv[4]) : 128, H = argc > 5 ? atoi(argv[5]) : 8;
  const int stages = argc > 6 ? atoi(argv[6]) : 4, bps = argc > 7 ? atoi(argv[7]) : 1, iters = argc > 8 ? atoi(argv[8]) : 20;
  const int pitch = (nx + kCol0 + 1 + 31) / 32 * 32, rows = ny + 2;
  const size_t plane = (size_t)rows * pitch, bytes = plane * 9 * 4;
  float *src, *dst;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes));
  CK(cudaMemset(src, 0x11, bytes)); CK(cudaMemset(dst, 0, bytes));
  Args a{src, dst, plane, pitch, nx, ny, stages, 8};
  Args b{dst, src, plane, pitch, nx, ny, stages, 8};
  CUtensorMap ms{}, md{};
  if (mode == "tma" || mode == "tmaws" || mode == "tmast") { ms = make_map(src, pitch, rows, plane, W, H); md = make_map(dst, pitch, rows, plane, W, H); }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  auto once = [&](bool flip) {
    const Args &x = flip ? b : a;
    const CUtensorMap &m0 = flip ? md : ms, &m1 = flip ? ms : md;
    if (mode == "copy") copy_kernel<<<148 * 8, 256, 0, st>>>((const float4 *)x.src, (float4 *)x.dst, bytes / 16);
    else if (mode == "copy4") copy4_kernel<<<(unsigned)((bytes / 16 + 1023) / 1024), 256, 0, st>>>((const float4 *)x.src, (float4 *)x.dst, bytes / 16);
    else if (mode == "blob" || mode == "blobldg" || mode == "blobst") {
      const int units = (nx / W) * (ny / H);
#define GL(WW, HH) if (W == WW && H == HH) { if (mode == "blob") blob_kernel<WW, HH, 0><<<units, 32 * HH, 0, st>>>(x); else if (mode == "blobldg") blob_kernel<WW, HH, 1><<<units, 32 * HH, 0, st>>>(x); else blob_kernel<WW, HH, 2><<<units, 32 * HH, 0, st>>>(x); }
      GL(128, 8) GL(128, 4) GL(128, 2) GL(256, 4)
    } else if (mode == "tile" || mode == "rowchunk") {
      const int grid = 148 * bps;
#define GB(WW, HH) if (W == WW && H == HH) { if (mode == "tile") launch_blocked<WW, HH, 0>(x, grid, st); else launch_blocked<WW, HH, 1>(x, grid, st); }
      GB(128, 8) GB(256, 4) GB(256, 8) GB(128, 4) GB(128, 16) GB(64, 16) GB(512, 2) GB(128, 2) GB(256, 2)
    } else if (mode == "ldg") {
      const int tx = W / 4;  // W = tile width here (e.g. 1024 -> 256 lanes)
      dim3 blk(tx, 256 / tx);
      Args y = x; y.rows_iter = H;
      if (bps == 8) ldg_kernel<8><<<148 * 8, blk, 0, st>>>(y);
      else if (bps == 6) ldg_kernel<6><<<148 * 6, blk, 0, st>>>(y);
      else if (bps == 2) ldg_kernel<2><<<148 * 2, blk, 0, st>>>(y);
      else if (bps == 3) ldg_kernel<3><<<148 * 3, blk, 0, st>>>(y);
      else ldg_kernel<4><<<148 * 4, blk, 0, st>>>(y);
    } else {
      const int m = mode == "tma" ? 0 : mode == "tmaws" ? 1 : 2;
      const int grid = 148 * bps;
#define GO(WW, HH) if (W == WW && H == HH) { if (m == 0) launch_tma<WW, HH, 0>(x, m0, m1, grid, st); else if (m == 1) launch_tma<WW, HH, 1>(x, m0, m1, grid, st); else launch_tma<WW, HH, 2>(x, m0, m1, grid, st); }
      GO(128, 8) GO(256, 4) GO(256, 8) GO(128, 4) GO(128, 16) GO(256, 2) GO(256, 16)
    }
    CK(cudaGetLastError());
  };
  for (int i = 0; i < 4; ++i) once(i & 1);
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) once(i & 1);
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms_ = 0; CK(cudaEventElapsedTime(&ms_, e0, e1));
  const double moved = (mode == "copy" || mode == "copy4") ? 2.0 * bytes : 2.0 * 9 * 4 * (double)nx * ny;
  // spot check of the destination (dst of the last a->b pass holds 0x11111111 everywhere it was written)
  float hv = 0; CK(cudaMemcpy(&hv, dst + 3 * plane + (size_t)(ny / 2) * pitch + kCol0 + nx / 2, 4, cudaMemcpyDeviceToHost));
  unsigned hb; memcpy(&hb, &hv, 4);
  printf("%-6s nx=%d ny=%d W=%d H=%d stages=%d bps=%d : %.4f ms/iter  %.1f GB/s  (check %08x)\n", mode.c_str(), nx, ny, W, H, stages, bps,
         ms_ / iters, moved / (ms_ / iters * 1e-3) / 1e9, hb);
  return 0;
}

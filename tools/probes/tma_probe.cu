// Minimal TMA 3-D tile-load probe (development tool): loads one box into shared memory and copies it out.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, const CUtensorMap *gmap, const float *src, float *out, int n, int c0, int c1, int c2, int mode) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ __align__(8) uint64_t bar;
  float *dst = reinterpret_cast<float *>(raw);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (mode == 0) return;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n * 4) : "memory");
    if (mode == 2) asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    if (mode == 3) asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(n * 4), "r"(smem_u32(&bar)) : "memory");
    if (mode == 4) asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(gmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    if (mode == 5) asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  if (mode == 1) return;
  unsigned ok = 0;
  for (long long spin = 0; !ok && spin < 20000000; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  if (!ok) { if (threadIdx.x == 0) out[n] = -1.f; return; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
  if (threadIdx.x == 0) out[n] = (float)(smem_u32(dst) & 1023);
}

int main(int argc, char **argv) {
  const int bw = argc > 1 ? atoi(argv[1]) : 256, bh = argc > 2 ? atoi(argv[2]) : 4, mode = argc > 3 ? atoi(argv[3]) : 2;
  const int pitch = 320, rows = 18, planes = 9;
  const size_t plane = (size_t)pitch * rows;
  std::vector<float> h(plane * planes);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  CK(cudaMalloc(&d, h.size() * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const int n = bw * bh;
  CK(cudaMalloc(&out, (n + 1) * 4));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap map;
  const cuuint64_t dims[3] = {pitch, rows, planes};
  const cuuint64_t strides[2] = {pitch * 4ull, plane * 4ull};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, mode == 5 ? 2 : 3, mode == 5 ? d + plane : d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box=%dx%d -> %d (qres %d)\n", bw, bh, (int)r, (int)q);
  if (r) return 1;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CUtensorMap *gmap;
  CK(cudaMalloc(&gmap, sizeof map));
  CK(cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice));
  probe<<<1, 128, 65536>>>(map, gmap, d + plane + 32, out, n, 31, 0, 1, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d kernel -> %s\n", mode, cudaGetErrorString(e));
  if (e != cudaSuccess || mode < 2) return e != cudaSuccess;
  std::vector<float> o(n + 1);
  CK(cudaMemcpy(o.data(), out, (n + 1) * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int rr = 0; rr < bh; ++rr)
    for (int cc = 0; cc < bw; ++cc)
      bad += o[rr * bw + cc] != (mode == 3 ? h[plane + 32 + rr * bw + cc] : h[1 * plane + (size_t)rr * pitch + 31 + cc]);
  printf("smem offset mod 1024 = %g, mismatches %d of %d\n", o[n], bad, n);
  return bad != 0;
}

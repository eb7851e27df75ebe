// TMA 2-D probe following the CUDA programming guide example as closely as possible (development tool).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, int *out, int x, int y) {
  __shared__ alignas(128) int buf[BH][BW];
  __shared__ alignas(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(blockDim.x) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  uint64_t tok;
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(&buf[0][0])), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(tok) : "r"(smem_u32(&bar)), "r"((unsigned)sizeof(buf)) : "memory");
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(tok) : "r"(smem_u32(&bar)) : "memory");
  }
  unsigned ok = 0;
  for (long long spin = 0; !ok && spin < 20000000; ++spin)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  if (!ok) return;
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&buf[0][0])[i];
}

int main(int argc, char **argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 1024, H = argc > 2 ? atoi(argv[2]) : 1024, promo = argc > 3 ? atoi(argv[3]) : 0;
  const int x0 = argc > 4 ? atoi(argv[4]) : 0, y0 = argc > 5 ? atoi(argv[5]) : 0;
  constexpr int BW = 32, BH = 8;
  std::vector<int> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (int)i;
  int *d, *out;
  CK(cudaMalloc(&d, h.size() * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, BW * BH * 4));
  CK(cudaMemset(out, 0xff, BW * BH * 4));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {BW, BH};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %dx%d promo %d -> %d; map bytes:", W, H, promo, (int)r);
  for (int i = 0; i < 64; ++i) printf(" %02x", ((unsigned char *)&map)[i]);
  printf("\n");
  if (r) return 1;
  probe<BW, BH><<<1, 128>>>(map, out, x0, y0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("x0=%d y0=%d kernel -> %s\n", x0, y0, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<int> o(BW * BH);
  CK(cudaMemcpy(o.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int rr = 0; rr < BH; ++rr) for (int cc = 0; cc < BW; ++cc) bad += o[rr * BW + cc] != h[(size_t)(rr + y0) * W + cc + x0];
  printf("mismatches %d\n", bad);
  return 0;
}

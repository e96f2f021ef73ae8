// tma_probe.cu -- minimal 3-D fp64 TMA box load (the staging step of pom_tma.h) checked against a host copy
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define BW 34
#define BH 17
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap m, double* out, int c0, int c1, int c2) {
  extern __shared__ __align__(128) double sm[];
  uint64_t* bar = (uint64_t*)(sm + 640);
  const bool w0 = (threadIdx.y == 0);
  bool leader;
  if (MODE == 0) leader = (threadIdx.x == 0 && threadIdx.y == 0);
  else {
    uint32_t p = 0;
    if (w0) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    leader = w0 && p;
  }
  if (leader) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (leader) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BW * BH * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(sm)), "l"((uint64_t)&m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)) : "memory");
  for (int e = threadIdx.y * 32 + threadIdx.x; e < BW * BH; e += 512) out[e] = sm[e];
}
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int im = 64, jm = 40, kb = 8;
  size_t n = (size_t)im * jm * kb;
  double* h = (double*)malloc(n * 8);
  for (size_t e = 0; e < n; ++e) h[e] = (double)e;
  double *d, *out; cudaMalloc(&d, n * 8); cudaMalloc(&out, BW * BH * 8);
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[3] = {im, jm, kb}, str[2] = {im * 8, (cuuint64_t)im * jm * 8};
  cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d mode=%d\n", (int)r, mode);
  int cs[6][3] = {{10, 5, 3}, {-2, 5, 3}, {10, -1, 3}, {-2, -1, 0}, {11, 5, 3}, {-1, 5, 3}};
  for (int t = 0; t < 6; ++t) {
    int c0 = cs[t][0], c1 = cs[t][1], c2 = cs[t][2];
    if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); k<0><<<1, dim3(32, 16), 100000>>>(m, out, c0, c1, c2); }
    else { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); k<1><<<1, dim3(32, 16), 100000>>>(m, out, c0, c1, c2); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("run %d (%d,%d,%d): %s\n", t, c0, c1, c2, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    double o[BW * BH]; cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
      int i = c0 + x, j = c1 + y;
      double want = (i < 0 || i >= im || j < 0 || j >= jm) ? 0. : h[(size_t)c2 * im * jm + (size_t)j * im + i];
      if (o[y * BW + x] != want) ++bad;
    }
    printf("  mismatches: %d\n", bad);
  }
  return 0;
}

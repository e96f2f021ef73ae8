// l2scratch_probe.cu -- can per-thread Thomas coefficients (ee/gg of solver.f:1393-1450, 1649-1680,
// 1739-1775) make their down-sweep -> up-sweep round trip through L2 WITHOUT reaching HBM?
// A persistent column kernel streams NIN input fields downward in k, parks NV values per level and
// thread, then sweeps upward reading them back and writing one output field.  Variants of where
// the parked values live:
//   0  per-thread local memory (what the round-1 kernels do)
//   1  explicit global scratch [slot][v][k][thread], plain ld/st
//   2  + L2 evict_last policy on the scratch accesses
//   3  + discard.global.L2 of each scratch line after its last read (no write-back)
//   4  + L2 evict_first policy on the streamed inputs / outputs
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2scratch_probe l2scratch_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define KB 41
#define NIN 7

__device__ __forceinline__ unsigned long long pol_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long pol_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ld_pol(const double* a, unsigned long long p) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(p));
  return v;
}
__device__ __forceinline__ void st_pol(double* a, double v, unsigned long long p) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(p) : "memory");
}
__device__ __forceinline__ void discard128(const void* a) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
}

template <int MODE, int NV>
__global__ void __launch_bounds__(256) probe(const double* __restrict__ in, double* __restrict__ out, double* scratch,
                                             int im, int jm, long n2, int kb) {
  extern __shared__ double pad[];   // occupancy control only
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int nbx = im / 32, ntiles = nbx * (jm / 8);
  const unsigned long long PL = pol_last(), PF = pol_first();
  double* my = scratch + (size_t)blockIdx.x * NV * KB * 256 + tid;
  double loc[MODE == 0 ? NV : 1][MODE == 0 ? 64 : 1];
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int i = (t % nbx) * 32 + threadIdx.x, j = (t / nbx) * 8 + threadIdx.y;
    const long col = i + (long)im * j;
    double carry = 0.;
    for (int k = 0; k < kb; ++k) {
      double x = 0.;
#pragma unroll
      for (int f = 0; f < NIN; ++f) {
        const double* a = in + (size_t)f * n2 * KB + col + n2 * k;
        x += (MODE >= 4) ? ld_pol(a, PF) : __ldcs(a);
      }
      carry = carry * 0.5 + x;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double c = carry * (v + 1);
        if (MODE == 0) loc[v][k] = c;
        else if (MODE == 1) my[(v * KB + k) * 256] = c;
        else st_pol(&my[(v * KB + k) * 256], c, PL);
      }
    }
    double acc = 0.;
    for (int k = kb - 1; k >= 0; --k) {
      double s = 0.;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        double c;
        if (MODE == 0) c = loc[v][k];
        else if (MODE == 1) c = my[(v * KB + k) * 256];
        else c = ld_pol(&my[(v * KB + k) * 256], PL);
        s += c;
      }
      acc = acc * 0.5 + s;
      double* o = out + col + n2 * k;
      if (MODE >= 4) st_pol(o, acc, PF); else __stcs(o, acc);
      if (MODE >= 3) {
        __syncwarp();
        if ((threadIdx.x & 15) == 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) discard128(&my[(v * KB + k) * 256]);
        }
      }
    }
  }
  if (pad[0] == 12345.678) out[0] = pad[1];
}

template <int MODE, int NV>
static void run(const double* in, double* out, double* scratch, int im, int jm, int bps, bool quiet) {
  const long n2 = (long)im * jm;
  size_t smem = (size_t)(227 * 1024) / bps - 2048;
  cudaFuncSetAttribute(probe<MODE, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE, NV>, 256, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<MODE, NV><<<148 * bps, dim3(32, 8), smem>>>(in, out, scratch, im, jm, n2, KB);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  const double alg = (double)(NIN + 1) * n2 * KB * 8;
  const double foot = 148.0 * bps * NV * KB * 256 * 8;
  if (!quiet)
    printf("mode %d NV %d blocks/SM %d (occ %d) scratch %6.1f MB : %7.3f ms  %7.1f GB/s algorithmic  %s\n", MODE, NV, bps, occ,
           foot / 1e6, best, alg / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  const int im = 1024, jm = 1024;
  const long n2 = (long)im * jm;
  int dev = 0; cudaSetDevice(dev);
  int l2 = 0, pers = 0, win = 0;
  cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
  cudaDeviceGetAttribute(&pers, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  printf("L2 %d MB, max persisting %d MB, max access-policy window %d MB\n", l2 >> 20, pers >> 20, win >> 20);
  double *in, *out, *scratch;
  cudaMalloc(&in, (size_t)NIN * n2 * KB * 8);
  cudaMalloc(&out, (size_t)n2 * KB * 8);
  cudaMalloc(&scratch, (size_t)148 * 4 * 4 * KB * 256 * 8);
  cudaMemset(in, 0, (size_t)NIN * n2 * KB * 8);
  const int only = argc > 1 ? atoi(argv[1]) : -1;   // one (mode*10+bps) combination, for ncu
#define RUN(M, V, B) if (only < 0 || only == (M) * 100 + (V) * 10 + (B)) run<M, V>(in, out, scratch, im, jm, B, false);
  for (int b = 1; b <= 4; ++b) {
    RUN(0, 2, b) RUN(1, 2, b) RUN(2, 2, b) RUN(3, 2, b) RUN(4, 2, b)
    RUN(0, 4, b) RUN(1, 4, b) RUN(2, 4, b) RUN(3, 4, b) RUN(4, 4, b)
  }
  cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// fp64_probe.cu -- latency / throughput of the fp64 operations the extPOM kernels are made of,
// on the GPU it runs on.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void chain(double* out, double a, double b, long long* cyc) {
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0 = clock64();
#pragma unroll 16
  for (int n = 0; n < N; ++n) {
    if (OP == 0) x = fma(x, y, a);
    if (OP == 1) x = x * y;
    if (OP == 2) x = x + y;
    if (OP == 3) x = a / (x + y);
    if (OP == 4) x = sqrt(x + y);
    if (OP == 5) x = 1. / (x + y);
    if (OP == 6) { double q0 = x * y; x = fma(fma(-q0, b, x), y, q0); }   // RDiv
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// ILP 4 independent chains per thread
template <int OP>
__global__ void tput(double* out, double a, double b) {
  double x0 = a + threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
#pragma unroll 4
  for (int n = 0; n < N; ++n) {
    if (OP == 0) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
    if (OP == 3) { x0 = a / (x0 + b); x1 = a / (x1 + b); x2 = a / (x2 + b); x3 = a / (x3 + b); }
    if (OP == 4) { x0 = sqrt(x0 + b); x1 = sqrt(x1 + b); x2 = sqrt(x2 + b); x3 = sqrt(x3 + b); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 8 * 1024 * 8); cudaMallocManaged(&cyc, 8);
  const char* nm[] = {"DFMA", "DMUL", "DADD", "a/(x+y)", "sqrt(x+y)", "1/(x+y)", "RDiv(mul+2fma)"};
#define RUNC(OP) chain<OP><<<1, 32>>>(out, 1.0000001, 0.9999999, cyc); cudaDeviceSynchronize(); chain<OP><<<1, 32>>>(out, 1.0000001, 0.9999999, cyc); cudaDeviceSynchronize(); printf("latency %-16s %7.1f cycles/iter (1 warp, dependent chain)\n", nm[OP], (double)*cyc / N);
  RUNC(0) RUNC(1) RUNC(2) RUNC(3) RUNC(4) RUNC(5) RUNC(6)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUNT(OP, ops) { tput<OP><<<148 * 8, 256>>>(out, 1.0000001, 0.9999999); cudaEventRecord(e0); tput<OP><<<148 * 8, 256>>>(out, 1.0000001, 0.9999999); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); \
    double n = 148.0 * 8 * 256 * N * 4; printf("throughput %-12s %8.2f G results/s  (%.1f per clk per SM at 1.9 GHz)\n", nm[OP], n / ms / 1e6, n / ms / 1e6 / 148 / 1.9); }
  RUNT(0, 1) RUNT(3, 1) RUNT(4, 1)
  return 0;
}

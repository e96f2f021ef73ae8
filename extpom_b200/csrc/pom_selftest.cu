// pom_selftest.cu -- device-side self tests of the arithmetic helpers the kernels rely on for
// bit-exactness: pdiv (nvcc's division sequence without its range test) and RDiv (hoisted
// reciprocal + residual correction) against the IEEE `/` on pseudo-random operand pairs.
#include "pom_core.h"

namespace pom {

POM_HD uint64_t mix64(uint64_t x) {   // splitmix64
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// a double with a uniformly random mantissa and sign and a decimal exponent in [-emax, emax]
POM_HD double rnd_double(uint64_t h, int emax) {
  const uint64_t m = h & 0x000FFFFFFFFFFFFFull;
  const int span = (int)(2 * emax * 3.3219) + 1;
  const int e = 1023 - span / 2 + (int)((h >> 52) % (uint64_t)span);
  const uint64_t bits = (h & 0x8000000000000000ull) | ((uint64_t)e << 52) | m;
  double d;
  memcpy(&d, &bits, 8);
  return d;
}

#ifndef POMGPU_EMU
__global__ void pdiv_test_kernel(long n, uint64_t seed, int emax, unsigned long long* bad) {
  unsigned long long nb = 0;
  for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    const uint64_t h = mix64(seed + 2 * (uint64_t)q);
    double a = rnd_double(h, emax);
    const double b = rnd_double(mix64(h), emax);
    if ((q & 1023) == 0) a = 0.;                 // zero numerators are common in the model (state of rest)
    const double want = a / b;
    const double got = pdiv(a, b);
    RDiv rd; rd.set(b);
    const double got2 = rd(a);
    if (__double_as_longlong(want) != __double_as_longlong(got)) ++nb;
    if (__double_as_longlong(want) != __double_as_longlong(got2)) ++nb;
  }
  if (nb) atomicAdd(bad, nb);
}
#endif

// number of (a,b) pairs, out of n, for which pdiv(a,b) or RDiv differ from a/b in any bit;
// operands span 10^-emax .. 10^emax in magnitude
long selftest_pdiv(Ctx* c, long n, unsigned long seed, int emax) {
#ifdef POMGPU_EMU
  (void)c; (void)n; (void)seed; (void)emax;
  return 0;
#else
  cudaSetDevice(c->device);
  unsigned long long* d = (unsigned long long*)c->d_red;
  cudaMemsetAsync(d, 0, 8, (cudaStream_t)c->stream);
  pdiv_test_kernel<<<148 * 8, 256, 0, (cudaStream_t)c->stream>>>(n, (uint64_t)seed, emax, d);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, (cudaStream_t)c->stream);
  if (cudaStreamSynchronize((cudaStream_t)c->stream) != cudaSuccess) return -1;
  return (long)h;
#endif
}

}  // namespace pom

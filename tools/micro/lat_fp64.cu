// microbenchmark: dependent-chain latencies of fp64 ops and the glibc-exact functions on B200
#include <cstdio>
#include <cuda_runtime.h>
#include "../../colate_b200/csrc/glibc_math.cuh"
__global__ void k(const uint64_t* et, const uint64_t* lt, double seed, double* out, long long* cyc)
{
  glm::Tables T{et, lt};
  double x = seed + threadIdx.x * 1e-3;
  long long t0, t1;
  const int N = 2000;
  // dependent DFMA
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __fma_rn(x, 0.999999, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / N;
  // dependent DADD
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __dadd_rn(x, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / N;
  // dependent DDIV
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __ddiv_rn(1.7, x + 0.3);
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / N;
  // dependent glm::exp
  double y = -x;
  t0 = clock64();
  for (int i = 0; i < N; i++) y = -glm::exp(y, T);
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0) / N;
  // dependent glm::log1p (reduce path: arg ~0.6)
  double z = 0.6 + x * 1e-3;
  t0 = clock64();
  for (int i = 0; i < N; i++) z = glm::log1p(z) + 0.2;
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) / N;
  // dependent glm::log1p (k = 0 path: arg ~0.1)
  double z2 = 0.1 + x * 1e-3;
  t0 = clock64();
  for (int i = 0; i < N; i++) z2 = glm::log1p(z2) + 0.01;
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0) / N;
  // dependent lse-like step
  double a = -3.0 + x * 1e-3;
  t0 = clock64();
  for (int i = 0; i < N; i++) { double b = -2.5 + 1e-3 * i; double hi = a > b ? a : b, lo = a > b ? b : a; a = hi + glm::log1p(glm::exp(lo - hi, T)) - 0.7; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0) / N;
  // dependent CUDA exp / log1p
  double w = -x;
  t0 = clock64();
  for (int i = 0; i < N; i++) w = -exp(w);
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0) / N;
  double v = 0.6 + x * 1e-3;
  t0 = clock64();
  for (int i = 0; i < N; i++) v = log1p(v) + 0.2;
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0) / N;
  // glm::log dependent
  double q = 3.0 + x;
  t0 = clock64();
  for (int i = 0; i < N; i++) q = glm::log(q, T) + 2.5;
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = (t1 - t0) / N;
  out[threadIdx.x] = x + y + z + z2 + a + w + v + q;
}
int main()
{
  uint64_t *et, *lt; double* out; long long* cyc;
  cudaMalloc(&et, 2048); cudaMalloc(&lt, 2048); cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 128);
  cudaMemcpy(et, glm::hEXP_TAB, 2048, cudaMemcpyHostToDevice); cudaMemcpy(lt, glm::hLOG_TAB, 2048, cudaMemcpyHostToDevice);
  const char* names[] = {"DFMA", "DADD", "DDIV(+add)", "glm::exp", "glm::log1p reduce", "glm::log1p k=0", "lse step", "cuda exp", "cuda log1p", "glm::log"};
  for (int threads : {1, 32, 256}) {
    k<<<1, threads>>>(et, lt, 0.5, out, cyc);
    long long h[16]; cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
    printf("threads %d:", threads);
    for (int i = 0; i < 10; i++) printf(" %s=%lld", names[i], h[i]);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

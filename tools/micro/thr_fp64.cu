// FP64 pipe throughput on this GPU: independent DFMA / DADD / I2F.F64 chains per thread, all SMs full.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o thr_fp64 thr_fp64.cu && ./thr_fp64
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double seed)
{
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-3 + i;
  int c = threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) a[i] = __fma_rn(a[i], 1.0000001, 1e-9);
      if (OP == 1) a[i] = __dadd_rn(a[i], 1e-9);
      if (OP == 2) { a[i] = (double)(c + i) ; c = __double2hiint(a[i]) + it; }          // I2F.F64.S32 (+ a move back)
      if (OP == 3) a[i] = __dmul_rn(a[i], 1.0000001);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + c;
}
int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const char* names[] = {"DFMA", "DADD", "I2F.F64.S32", "DMUL"};
  for (int op = 0; op < 4; op++) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (op == 0) k<0><<<sms * 8, 256>>>(out, iters, 1.0);
      if (op == 1) k<1><<<sms * 8, 256>>>(out, iters, 1.0);
      if (op == 2) k<2><<<sms * 8, 256>>>(out, iters, 1.0);
      if (op == 3) k<3><<<sms * 8, 256>>>(out, iters, 1.0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * 8 * 256 * iters * 8;
    printf("%-12s %.3f ms  %.2f Tops/s  = %.1f lanes/clk/SM at 1.92 GHz\n", names[op], ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.92e9);
  }
  return 0;
}

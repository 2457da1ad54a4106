// microbenchmark: cost of warp-level histogram primitives on B200 (cycles per warp-instruction per SM)
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void k(int mode, int iters, const int* binsrc, long long* cyc, unsigned* sink)
{
  __shared__ unsigned rows[32][64];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) ((unsigned*)rows)[i] = 0;
  __syncthreads();
  int b = binsrc[(blockIdx.x * blockDim.x + threadIdx.x) & 4095];
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    int bb = ((b + i) & 15) + ((lane >> 3) & (i & 3)) + 40;   // ~16-19 distinct bins, duplicates across lanes
    if (mode == 0) { acc += __match_any_sync(0xffffffffu, bb); }
    else if (mode == 1) { atomicAdd(&rows[warp][bb >> 2], 1u << (8 * (bb & 3))); }
    else if (mode == 2) { acc += __reduce_add_sync(0xffffffffu, (unsigned)bb); }
    else if (mode == 3) { acc += __ballot_sync(0xffffffffu, bb == 45); }
    else if (mode == 4) { acc += __shfl_xor_sync(0xffffffffu, (unsigned)bb, 5); }
    else if (mode == 5) { unsigned m = __match_any_sync(0xffffffffu, bb); if (lane == __ffs(m) - 1) ((unsigned char*)rows[warp])[bb] = __popc(m); }
    else if (mode == 7) { acc += bb; }
    else if (mode == 6) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&rows[warp][bb >> 2])), "r"(1u << (8 * (bb & 3)))); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + rows[warp][lane];
}
int main()
{
  int* src; long long* cyc; unsigned* sink;
  cudaMalloc(&src, 4096 * 4); cudaMalloc(&cyc, 8 * 2048); cudaMalloc(&sink, 4 * 2048 * 1024);
  int h[4096]; for (int i = 0; i < 4096; i++) h[i] = (i * 2654435761u) >> 20;
  cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  const char* names[] = {"match_any", "atomicAdd smem (packed u8)", "redux.add", "ballot", "shfl", "match+leader store", "red.shared.add", "baseline (no op)"};
  for (int warps : {1, 8, 16, 32}) {
    for (int mode = 0; mode < 8; mode++) {
      int iters = 2000;
      k<<<148, warps * 32>>>(mode, iters, src, cyc, sink);
      long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
      printf("warps/SM %2d  %-28s %7.1f cycles per iteration per warp -> %6.2f cycles per warp-instr per SM\n", warps, names[mode], (double)hc / iters, (double)hc / iters / warps);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

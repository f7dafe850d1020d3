// Microbenchmark: DMMA issue rate with the register pattern of a real GEMM inner loop
// (4 A fragments x 8 B fragments -> 32 accumulators per warp), with and without LDS traffic.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>   // 0: operands fixed in registers, 1: operands reloaded from smem each k-step, 2: + software pipelining
__global__ void __launch_bounds__(256, 1) k(double *out, int iters) {
  __shared__ double sa[64 * 36], sb[32 * 68];
  for (int i = threadIdx.x; i < 64 * 36; i += 256) sa[i] = 1.0 + i * 1e-9;
  for (int i = threadIdx.x; i < 32 * 68; i += 256) sb[i] = 0.5 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane >> 2, lc = lane & 3;
  double acc[4][8][2];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  double af[4], bf[8];
  for (int i = 0; i < 4; i++) af[i] = sa[(((warp + i * 8) * 8 + lr) & 63) * 36 + lc];
  for (int j = 0; j < 8; j++) bf[j] = sb[lc * 68 + j * 8 + lr];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      if (MODE >= 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = sa[(((warp + i * 8) * 8 + lr) & 63) * 36 + ks * 4 + lc];
#pragma unroll
        for (int j = 0; j < 8; j++) bf[j] = sb[(ks * 4 + lc) * 68 + j * 8 + lr];
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (MODE == 2) __syncthreads();
  }
  double s = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double *out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  k<MODE><<<148, 256>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s %7.2f TFLOP/s\n", name, 148.0 * 8 * iters * 8 * 32 * 512.0 / (ms * 1e-3) * 1e-12);
}
int main() {
  double *out; cudaMalloc(&out, 148 * 256 * 8);
  run<0>("8 warps, 4x8 frags in registers", out);
  run<1>("8 warps, 12 LDS per 32 DMMA", out);
  run<2>("8 warps, LDS + barrier per 256 DMMA", out);
  return 0;
}

// Microbenchmark: FP64 issue-rate ceilings on this GPU (DFMA vs DMMA shapes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
  double a[16];
  double x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double *out, int iters) {
  double c[8][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double *out, int iters) {
  double c[4][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 0.5 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
          : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
          : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
            "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double *out, int iters) {
  double c[4][4];
  double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  b[0] = 0.5; b[1] = 0.25;
#pragma unroll
  for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
          : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
          : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", p.name, nsm, p.clockRate);
  double *out;
  cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = warps * 32, blocks = nsm * 2;
    double t = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("DFMA      warps/blk %2d: %7.2f TFLOP/s\n", warps, 2.0 * 16 * iters * (double)threads * blocks / t * 1e-12);
    t = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
    printf("DMMA 884  warps/blk %2d: %7.2f TFLOP/s\n", warps, 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * blocks / t * 1e-12);
    t = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
    printf("DMMA 1688 warps/blk %2d: %7.2f TFLOP/s\n", warps, 2.0 * 16 * 8 * 8 * 4 * iters * (double)warps * blocks / t * 1e-12);
    t = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters); });
    printf("DMMA16816 warps/blk %2d: %7.2f TFLOP/s\n", warps, 2.0 * 16 * 8 * 16 * 4 * iters * (double)warps * blocks / t * 1e-12);
  }
  return 0;
}

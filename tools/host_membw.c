// Host memory bandwidth probe: parallel zero-fill and max-abs scan of a 1.76 GB matrix.
// gcc -O3 -fopenmp -march=native tools/host_membw.c -o /tmp/host_membw
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
int main() {
  const size_t n = (size_t)14832 * 14832;
  double *a = (double *)malloc(n * sizeof(double));
  memset(a, 0, n * sizeof(double));
  printf("threads %d\n", omp_get_max_threads());
  for (int rep = 0; rep < 3; rep++) {
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < 14832; c++) memset(a + c * 14832, 0, 14832 * sizeof(double));
    double t1 = omp_get_wtime();
    double mx = 0;
#pragma omp parallel for schedule(static) reduction(max : mx)
    for (size_t i = 0; i < n; i++) mx = fmax(mx, fabs(a[i]));
    double t2 = omp_get_wtime();
    printf("memset %.1f ms (%.1f GB/s)  scan %.1f ms (%.1f GB/s) mx %g\n", 1e3 * (t1 - t0), n * 8e-9 / (t1 - t0),
           1e3 * (t2 - t1), n * 8e-9 / (t2 - t1), mx);
  }
  return 0;
}

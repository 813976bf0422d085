// Which warps of a thread block share an SM sub-partition?  Block of 8 warps on one SM; warps a and b each run an
// FP64-pipe-saturating loop (8 independent DFMA chains), the others exit.  If a and b sit on the same sub-partition
// the pair takes twice as long as one warp alone.   nvcc -arch=sm_100a -O3 -o smsp_probe smsp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int a, int b, int iters, double* out, long long* cyc) {
  const int w = threadIdx.x >> 5;
  if (w != a && w != b) return;
  double v[8];
  for (int i = 0; i < 8; ++i) v[i] = 1e-3 * (threadIdx.x + i);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fma(v[j], 0.999999, 1e-7);
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256 * 8); cudaMallocManaged(&cyc, 8 * 8);
  const int iters = 20000;
  for (int nthreads : {256, 64}) {
    printf("block of %d threads: cycles per DFMA for warp pairs (a alone on the diagonal)\n     ", nthreads);
    const int nw = nthreads / 32;
    for (int b = 0; b < nw; ++b) printf("   b=%d", b);
    printf("\n");
    for (int a = 0; a < nw; ++a) {
      printf("a=%d  ", a);
      for (int b = 0; b < nw; ++b) {
        k<<<1, nthreads>>>(a, b, iters, out, cyc);
        cudaDeviceSynchronize();
        printf(" %5.2f", double(cyc[a]) / (iters * 8.0));
      }
      printf("\n");
    }
  }
  return 0;
}

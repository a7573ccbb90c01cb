// Micro-benchmark: does ONE warp overlap independent FP64 operations on B200? K independent DFMA chains
// per thread, W warps in the CTA (1 CTA): cycles per DFMA warp-instruction issued by each warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_ilp tools/ubench_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void ilp(double *out, long long *cyc, int n, double a, double b) {
  double x[K];
#pragma unroll
  for (int k = 0; k < K; ++k) x[k] = a + threadIdx.x * 1e-9 + k;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < K; ++k) x[k] = fma(x[k], b, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int K>
void run(int warps) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int n = 2048;
  ilp<K><<<1, 32 * warps>>>(out, cyc, n, 0.5, 0.5);
  ilp<K><<<1, 32 * warps>>>(out, cyc, n, 0.5, 0.5);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("chains/thread %d  warps/CTA %2d (%.1f per scheduler): %6.2f cycles per DFMA per warp\n", K, warps, warps / 4.0,
         (double)h / (4.0 * n * K));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w); run<2>(w); run<4>(w); run<8>(w);
  }
  return 0;
}

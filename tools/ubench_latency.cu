// Micro-benchmark: dependent-issue latencies that bound a one-warp-per-scheduler rollout on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench tools/ubench_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../predictive-multi-agent-framework_b200/csrc/pmaf_math.cuh"
using namespace pmaf;

template <int OP>
__global__ void chain(double *out, long long *cyc, int n, double a, double b) {
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    if (OP == 0) { x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); }
    if (OP == 1) { x = x + y; x = x + y; x = x + y; x = x + y; }
    if (OP == 2) { x = x * y; x = x * y; x = x * y; x = x * y; }
    if (OP == 3) { x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
                   x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); }
    if (OP == 4) { x = sqrt(x) + y; x = sqrt(x) + y; x = sqrt(x) + y; x = sqrt(x) + y; }
    if (OP == 5) { FastMath m; x = m.sqrt_(x) + y; x = m.sqrt_(x) + y; x = m.sqrt_(x) + y; x = m.sqrt_(x) + y; if (m.bad()) x = 0; }
    if (OP == 6) { x = y / x + y; x = y / x + y; x = y / x + y; x = y / x + y; }
    if (OP == 7) { FastMath m; x = m.div_(y, x) + y; x = m.div_(y, x) + y; x = m.div_(y, x) + y; x = m.div_(y, x) + y; if (m.bad()) x = 0; }
    if (OP == 8) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y;
                   asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y;
                   asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y;
                   asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y; }
    if (OP == 9) { float f = (float)x; f = f * 1.0001f + 0.5f; f = f * 1.0001f + 0.5f; f = f * 1.0001f + 0.5f; f = f * 1.0001f + 0.5f; x = f; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, double a, double b, int extra_per_iter) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  const int n = 4096;
  chain<OP><<<1, 32>>>(out, cyc, n, a, b);
  chain<OP><<<1, 32>>>(out, cyc, n, a, b);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %7.1f cycles per op (incl. %d dependent DADD)\n", name, (double)h / (4.0 * n), extra_per_iter);
}

int main() {
  run<0>("DFMA dependent", 0.5, 0.5, 0);
  run<1>("DADD dependent", 0.5, 1e-9, 0);
  run<2>("DMUL dependent", 1.0, 1.0000001, 0);
  run<3>("SHFL (double = 2 SHFL) dependent", 0.5, 0.5, 0);
  run<4>("sqrt() built-in + DADD", 2.0, 1.5, 1);
  run<5>("FastMath sqrt_ + DADD", 2.0, 1.5, 1);
  run<6>("division built-in + DADD", 2.0, 1.5, 1);
  run<7>("FastMath div_ + DADD", 2.0, 1.5, 1);
  run<8>("MUFU.RSQ64H + DADD", 2.0, 1.5, 1);
  run<9>("FFMA dependent (4) + 2 cvt", 2.0, 1.5, 0);
  return 0;
}

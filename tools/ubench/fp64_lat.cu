// Microbenchmark: dependent-chain latency and throughput of DFMA / DADD / DMUL and of shared-memory LDS.128 on
// this GPU (cycles per instruction), to size the per-mode phase of the LaplaceCyclic rows kernels.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lds(double* out, long long* cyc, int iters) {
  __shared__ double2 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2(i ^ 1, 0.0);
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    idx = static_cast<int>(buf[idx & 1023].x);   // dependent loads
  }
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    k_dfma<1><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP=1 warps/SM=%2d: %.2f cycles per dependent step\n", warps, double(h) / iters);
    k_dfma<2><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP=2 warps/SM=%2d: %.2f cycles per step (2 instr)\n", warps, double(h) / iters);
    k_dfma<4><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP=4 warps/SM=%2d: %.2f cycles per step (4 instr)\n", warps, double(h) / iters);
    k_dfma<8><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP=8 warps/SM=%2d: %.2f cycles per step (8 instr) -> %.2f warp-instr/cycle/SM\n", warps, double(h) / iters,
           8.0 * warps * iters / double(h));
  }
  k_lds<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent: %.2f cycles per load (+ F2I)\n", double(h) / iters);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock %d kHz\n", clk);
  return 0;
}

// FP64 latency / throughput probes (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dfma_chain(double* out, double a, double b, int iters) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[4096] = (double)(t1 - t0) / iters;
}
__global__ void lds_chain(int* out, int iters) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 1) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = s[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) out[64] = (int)((t1 - t0) / iters);
}
__global__ void dmma_chain(double* out, int iters) {
  double d0 = 0, d1 = 0, a = threadIdx.x, b = 1.0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[threadIdx.x] = d0 + d1;
  if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / iters;
}
int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 8192);
  double h;
  #define RUN(CH, W) dfma_chain<CH><<<1, 32 * W>>>(out, 1.0000001, 1e-9, 4096); cudaMemcpy(&h, out + 4096, 8, cudaMemcpyDeviceToHost); \
     printf("DFMA: %d independent chains/thread, %d warps: %.1f cycles per round (%.1f per DFMA per warp)\n", CH, W, h, h / CH);
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(4, 4) RUN(8, 8) RUN(8, 16)
  int* io; cudaMalloc(&io, 4096); int hi;
  lds_chain<<<1, 32>>>(io, 4096); cudaMemcpy(&hi, io + 64, 4, cudaMemcpyDeviceToHost); printf("dependent LDS latency: %d cycles\n", hi);
  dmma_chain<<<1, 32>>>(out, 4096); cudaMemcpy(&h, out + 64, 8, cudaMemcpyDeviceToHost); printf("dependent DMMA.8x8x4 latency: %.1f cycles\n", h);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

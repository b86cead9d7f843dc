// Latency probes used to calibrate the persistent local-update kernel's design (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
#define N (1 << 20)   // 16 MiB of double2
__global__ void chase(const int* __restrict__ next, int steps, long long* out, int mode) {
  int p = threadIdx.x + blockIdx.x * 7919;
  p %= N;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    int q;
    if (mode == 0) q = __ldcg(next + p);
    else asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(q) : "l"(next + p));
    p = q;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = p; }
}
// block 0 publishes a counter, block 1 polls it and answers; measures the store->visible->load round trip across SMs
__global__ void pingpong(volatile int* a, volatile int* b, int rounds, long long* out) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  if (blockIdx.x == 0) {
    for (int r = 1; r <= rounds; ++r) { *a = r; while (*b != r) { } }
  } else {
    for (int r = 1; r <= rounds; ++r) { while (*a != r) { } *b = r; }
  }
  out[blockIdx.x] = clock64() - t0;
}
int main() {
  int* h = new int[N];
  unsigned long long x = 88172645463325252ull;
  int* perm = new int[N];
  for (int i = 0; i < N; ++i) perm[i] = i;
  for (int i = N - 1; i > 0; --i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; int j = x % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
  for (int i = 0; i < N; ++i) h[perm[i]] = perm[(i + 1) % N];
  int* d; long long* out; int* flags;
  cudaMalloc(&d, sizeof(int) * N); cudaMalloc(&out, sizeof(long long) * 1024); cudaMalloc(&flags, 512);
  cudaMemcpy(d, h, sizeof(int) * N, cudaMemcpyHostToDevice); cudaMemset(flags, 0, 512);
  long long ho[1024];
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {1, 128}) {
      chase<<<grid, 32>>>(d, 2000, out, mode);   // warm L2
      chase<<<grid, 32>>>(d, 2000, out, mode);
      cudaMemcpy(ho, out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
      printf("%s dependent load latency, %3d CTAs: %.0f cycles\n", mode ? "ld.volatile" : "ld.global.cg", grid, ho[0] / 2000.0);
    }
  pingpong<<<2, 32>>>(flags, flags + 64, 1000, out);
  cudaMemcpy(ho, out, sizeof(long long) * 2, cudaMemcpyDeviceToHost);
  printf("cross-SM store->poll round trip (2 hops): %.0f cycles per round\n", ho[0] / 1000.0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}

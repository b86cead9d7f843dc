// Cluster exchange latency probes (not part of the product): ping-pong between CTA 0 and CTA 1 of an 8-CTA cluster.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__global__ void __cluster_dims__(8, 1, 1) probe(long long* out, int rounds) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank();
  __shared__ __align__(16) double2 slot[2];
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ volatile unsigned long long flag;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    flag = 0;
  }
  __syncthreads();
  cl.sync();
  // (a) st.async + mbarrier
  if (threadIdx.x == 0 && rank < 2) {
    const uint32_t peer = rank ^ 1;
    const uint32_t rslot = mapa_u32(smem_u32(&slot[0]), peer), rbar = mapa_u32(smem_u32(&bar[0]), peer), lbar = smem_u32(&bar[0]);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(lbar) : "memory");
      if (rank == 0) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(rslot), "d"(1.0), "d"(2.0), "r"(rbar) : "memory");
      uint32_t done = 0;
      while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(lbar), "r"(r & 1) : "memory");
      if (rank == 1) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(rslot), "d"(1.0), "d"(2.0), "r"(rbar) : "memory");
    }
    out[rank] = clock64() - t0;
  }
  cl.sync();
  // (b) plain remote store + local poll
  if (threadIdx.x == 0 && rank < 2) {
    const uint32_t peer = rank ^ 1;
    const uint32_t rflag = mapa_u32(smem_u32((const void*)&flag), peer);
    long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
      if (rank == 0) {
        asm volatile("st.volatile.shared::cluster.u64 [%0], %1;" ::"r"(rflag), "l"((unsigned long long)r) : "memory");
        while (flag != (unsigned long long)r) { }
      } else {
        while (flag != (unsigned long long)r) { }
        asm volatile("st.volatile.shared::cluster.u64 [%0], %1;" ::"r"(rflag), "l"((unsigned long long)r) : "memory");
      }
    }
    out[2 + rank] = clock64() - t0;
  }
  cl.sync();
  // (c) cluster barrier, all threads
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) cl.sync();
  if (threadIdx.x == 0 && rank == 0) out[4] = clock64() - t0;
  // (d) arrive.release / wait.acquire split, all threads
  t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0 && rank == 0) out[5] = clock64() - t0;
  // (e) __syncthreads
  t0 = clock64();
  for (int r = 0; r < rounds; ++r) __syncthreads();
  if (threadIdx.x == 0 && rank == 0) out[6] = clock64() - t0;
}
__global__ void fp64_lat(double* out, double a) {
  double x = a; long long t0 = clock64();
  for (int i = 0; i < 256; ++i) x = sqrt(x + 1.5);
  long long t1 = clock64();
  for (int i = 0; i < 256; ++i) x = 1.0 / (x + 1.5);
  long long t2 = clock64();
  for (int i = 0; i < 256; ++i) x = rsqrt(x + 1.5);
  long long t3 = clock64();
  for (int i = 0; i < 256; ++i) x = fma(x, a, 1.5);
  long long t4 = clock64();
  for (int i = 0; i < 256; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  long long t5 = clock64();
  out[0] = x; out[1] = (t1 - t0) / 256.0; out[2] = (t2 - t1) / 256.0; out[3] = (t3 - t2) / 256.0; out[4] = (t4 - t3) / 256.0; out[5] = (t5 - t4) / 256.0;
}
int main() {
  long long* out; cudaMalloc(&out, 64 * 8); long long h[8];
  for (int thr : {32, 256}) {
    probe<<<8, thr>>>(out, 1000); probe<<<8, thr>>>(out, 1000);
    cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    printf("threads/CTA %d: st.async+mbarrier round trip %.0f cyc (2 hops); remote st + poll round trip %.0f cyc; cluster.sync %.0f cyc; relaxed arrive+wait %.0f; __syncthreads %.0f\n",
           thr, h[0] / 1000.0, h[2] / 1000.0, h[4] / 1000.0, h[5] / 1000.0, h[6] / 1000.0);
  }
  double* d; cudaMalloc(&d, 64); double hd[8];
  fp64_lat<<<1, 32>>>(d, 1.0000001); cudaMemcpy(hd, d, 48, cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles): sqrt %.0f, div %.0f, rsqrt %.0f, dfma %.0f, shfl(double)+add %.0f\n", hd[1], hd[2], hd[3], hd[4], hd[5]);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

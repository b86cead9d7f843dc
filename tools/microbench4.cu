// All-to-all exchange probes inside an 8-CTA cluster, same traffic pattern as the QR panel kernel (not part of the product):
// every CTA sends 32 x 16 B to every CTA per round and then consumes the 8 x 32 values it received.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra.uni WD;\nbra.uni WL;\nWD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
#define SENT 0xFFFFFFFFFFFFFFFFull
template <int MODE>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256) xchg(long long* out, double* sink, int rounds) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank(), tid = threadIdx.x, g = tid >> 5, c = tid & 31;
  __shared__ __align__(16) double2 xch[2][8][32];
  __shared__ __align__(16) double2 stage[2][32];
  __shared__ __align__(8) unsigned long long full[2];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < 2 * 8 * 32; e += 256) { ((unsigned long long*)xch)[2 * e] = SENT; ((unsigned long long*)xch)[2 * e + 1] = SENT; }
  __syncthreads();
  const uint32_t r_xch = mapa_u32(smem_u32(&xch[0][rank][c]), g), r_bar = mapa_u32(smem_u32(&full[0]), g), l_bar = smem_u32(&full[0]);
  const uint32_t r_xrow = mapa_u32(smem_u32(&xch[0][rank][0]), c & 7);
  const uint32_t r_bar7 = mapa_u32(smem_u32(&full[0]), c & 7);
  cl.sync();
  double acc = tid;
  long long t0 = clock64();
  for (int j = 0; j < rounds; ++j) {
    const int par = j & 1;
    double2 v = make_double2(acc * 1e-3 + j, 1.0);
    if (MODE == 0) {            // st.async, 16 B per thread, complete_tx per store
      if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(l_bar + 8 * par), "r"(8 * 32 * 16) : "memory");
      __syncthreads();
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(r_xch + par * 8 * 32 * 16), "d"(v.x), "d"(v.y), "r"(r_bar + 8 * par) : "memory");
      mbar_wait(l_bar + 8 * par, (j >> 1) & 1);
      double s = 0;
#pragma unroll
      for (int src = 0; src < 8; ++src) s += xch[par][src][c].x;
      acc = s;
    } else if (MODE == 1) {     // plain remote stores, consumers poll on the data (sentinel), then re-arm
      __syncthreads();
      asm volatile("st.volatile.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(r_xch + par * 8 * 32 * 16), "d"(v.x), "d"(v.y) : "memory");
      double s = 0;
#pragma unroll
      for (int src = 0; src < 8; ++src) {
        unsigned long long x, y;
        const uint32_t a = smem_u32(&xch[par][src][c]);
        do { asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a) : "memory"); } while (x == SENT || y == SENT);
        s += __longlong_as_double(x);
      }
      __syncthreads();          // (all warps read the same slots here; the real kernel would own slots per warp)
      if (g == 0) {
#pragma unroll
        for (int src = 0; src < 8; ++src) asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %1};" ::"r"(smem_u32(&xch[par][src][c])), "l"(SENT) : "memory");
      }
      acc = s;
    } else {                    // one bulk copy of 512 B per destination, issued by 8 lanes of warp 0
      if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(l_bar + 8 * par), "r"(8 * 32 * 16) : "memory");
      if (g == 0) stage[par][c] = v;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid < 8)
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(r_xrow + par * 8 * 32 * 16), "r"(smem_u32(&stage[par][0])), "r"(512), "r"(r_bar7 + 8 * par) : "memory");
      mbar_wait(l_bar + 8 * par, (j >> 1) & 1);
      double s = 0;
#pragma unroll
      for (int src = 0; src < 8; ++src) s += xch[par][src][c].x;
      acc = s;
    }
  }
  long long t1 = clock64();
  sink[blockIdx.x * 256 + tid] = acc;
  if (tid == 0 && rank == 0) out[MODE] = t1 - t0;
  cl.sync();
}
int main() {
  long long* out; double* sink; cudaMalloc(&out, 64); cudaMalloc(&sink, 8 * 256 * 8 * 8); long long h[4];
  const int R = 2000;
  xchg<0><<<8, 256>>>(out, sink, R); xchg<0><<<8, 256>>>(out, sink, R);
  printf("st.async: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  xchg<1><<<8, 256>>>(out, sink, R); xchg<1><<<8, 256>>>(out, sink, R);
  printf("poll: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  xchg<2><<<8, 256>>>(out, sink, R); xchg<2><<<8, 256>>>(out, sink, R);
  printf("bulk: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  printf("cycles per all-to-all round (8 CTAs, 4 KB in per CTA): st.async+mbarrier %.0f, remote st + sentinel poll %.0f, bulk copy + mbarrier %.0f\n",
         h[0] / (double)R, h[1] / (double)R, h[2] / (double)R);
}

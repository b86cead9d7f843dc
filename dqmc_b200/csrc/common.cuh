// Shared device/host helpers for libdqmc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

typedef double2 cplx;  // interleaved (re, im) == Julia ComplexF64

#define DQMC_HD __host__ __device__ __forceinline__
#define DQMC_D __device__ __forceinline__

DQMC_HD cplx cmake(double r, double i) { return make_double2(r, i); }
DQMC_HD cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
DQMC_HD cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
DQMC_HD cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DQMC_HD cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
DQMC_HD cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
DQMC_HD cplx cneg(cplx a) { return make_double2(-a.x, -a.y); }
// acc += a*b
DQMC_HD void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
DQMC_HD void cfma_conj(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
DQMC_HD cplx cdiv(cplx a, cplx b) {
  double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
DQMC_HD double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

enum { OP_N = 0, OP_T = 1, OP_C = 2, OP_J = 3 };  // none, transpose, conj-transpose, conj

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      snprintf(g_errbuf, sizeof(g_errbuf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,    \
               cudaGetErrorString(_e));                                                     \
      return -1;                                                                            \
    }                                                                                       \
  } while (0)

extern char g_errbuf[512];
extern long long g_launches;   // kernels launched by this library (all contexts)

// Opt a kernel into the largest dynamic shared-memory size the device allows next to its static usage.
template <typename K>
static inline int set_max_dynamic_smem(K kernel, size_t* limit_out) {
  cudaFuncAttributes fa;
  CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
  const size_t lim = (size_t)227 * 1024 - fa.sharedSizeBytes;
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
  if (limit_out) *limit_out = lim;
  return 0;
}

// Function attributes are per DEVICE, and a process may hold contexts on several GPUs: every opt-in is memoised per device
// ordinal (a process-wide flag would leave the kernels of a second device at the 48 KB default).
#define DQMC_MAX_DEVICES 64
struct SmemMemo { size_t lim[DQMC_MAX_DEVICES]; };   // zero-initialised static at every call site
static inline int current_device_slot(int* dev) {
  CUDA_TRY(cudaGetDevice(dev));
  if (*dev < 0 || *dev >= DQMC_MAX_DEVICES) { snprintf(g_errbuf, sizeof(g_errbuf), "device ordinal %d out of range", *dev); return -1; }
  return 0;
}
// largest dynamic shared memory next to the kernel's static usage; limit of the current device in *limit_out
template <typename K>
static inline int ensure_max_dynamic_smem(K kernel, SmemMemo& memo, size_t* limit_out) {
  int dev = 0;
  if (current_device_slot(&dev)) return -1;
  if (memo.lim[dev] == 0) {
    size_t lim = 0;
    if (set_max_dynamic_smem(kernel, &lim)) return -1;
    memo.lim[dev] = lim;
  }
  if (limit_out) *limit_out = memo.lim[dev];
  return 0;
}
// fixed dynamic shared-memory size `bytes`
template <typename K>
static inline int ensure_dynamic_smem(K kernel, SmemMemo& memo, size_t bytes) {
  int dev = 0;
  if (current_device_slot(&dev)) return -1;
  if (memo.lim[dev] < bytes) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    memo.lim[dev] = bytes;
  }
  return 0;
}

// FP64 tensor-core MMA (DMMA.8x8x4): D(8x8) += A(8x4,row) * B(4x8,col).
// lane holds a = A[lane/4][lane%4], b = B[lane%4][lane/4], d0/d1 = D[lane/4][2*(lane%4) + {0,1}].
DQMC_D void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

DQMC_D double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DQMC_D double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Grid-wide barrier for kernels launched cooperatively (all CTAs co-resident).
// `bar` points to two zero-initialised unsigned ints {count, generation}.
DQMC_D void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* gen = bar + 1;
    unsigned int g = *gen;
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      bar[0] = 0;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == g) { }
    }
    __threadfence();
  }
  __syncthreads();
}

// Same, with a monotonic arrival counter and no generation word: CTA k-th barrier of a launch waits for the counter to reach
// k * nblocks, so the release is the last arrival's atomic itself (one L2 round trip less than the count + generation scheme).
// The counter must be zero at launch (the local-update kernel alternates between two counters and clears the idle one).
// The same barrier in two halves, for work that may run between a CTA's arrival and the release (it must neither read what
// other CTAs write before the barrier nor be needed by them): arrive = everything this CTA wrote so far is published.
DQMC_D void grid_barrier_mono_arrive(unsigned int* ctr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
  }
}
DQMC_D void grid_barrier_mono_wait(unsigned int* ctr, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}
DQMC_D void grid_barrier_mono(unsigned int* ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

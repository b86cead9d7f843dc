#include "misc.cuh"

__global__ void colnorm2_kernel(const cplx* __restrict__ A, int lda, int n, double* __restrict__ out) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int r = threadIdx.x; r < n; r += blockDim.x) acc += cabs2(A[(size_t)c * lda + r]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    out[c] = t;
  }
}
int colnorm2(cudaStream_t st, const cplx* A, int lda, int n, double* out) {
  colnorm2_kernel<<<n, 256, 0, st>>>(A, lda, n, out);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// Bitonic sort of (key, index) pairs, descending by key, ties by ascending index (= stable sort).
__global__ void __launch_bounds__(1024) argsort_desc_kernel(const double* __restrict__ key, int n, int* __restrict__ perm) {
  __shared__ double k[2048];
  __shared__ int ix[2048];
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    k[i] = i < n ? key[i] : -1.0;   // keys are squared norms (>= 0); padding sorts last
    ix[i] = i;
  }
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < np2 / 2; t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        double ka = k[lo], kb = k[hi];
        int ia = ix[lo], ib = ix[hi];
        bool a_first = (ka > kb) || (ka == kb && ia < ib);   // "a belongs before b" in descending order
        if (a_first != desc) { k[lo] = kb; k[hi] = ka; ix[lo] = ib; ix[hi] = ia; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = ix[i];
}
int argsort_desc(cudaStream_t st, const double* key, int n, int* perm) {
  if (n > 2048) { snprintf(g_errbuf, sizeof(g_errbuf), "argsort: n=%d > 2048", n); return -1; }
  argsort_desc_kernel<<<1, 1024, 0, st>>>(key, n, perm);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

__global__ void gather_cols_kernel(const cplx* __restrict__ in, int ldi, int n, const int* __restrict__ perm,
                                   cplx* __restrict__ out, int ldo) {
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    const int src = perm[c];
    for (int r = threadIdx.x; r < n; r += blockDim.x) out[(size_t)c * ldo + r] = in[(size_t)src * ldi + r];
  }
}
int gather_cols(cudaStream_t st, const cplx* in, int ldi, int n, const int* perm, cplx* out, int ldo, int num_sms) {
  gather_cols_kernel<<<min(n, num_sms * 8), 256, 0, st>>>(in, ldi, n, perm, out, ldo);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

__global__ void build_T_kernel(const cplx* __restrict__ R, int ldr, int n, const double* __restrict__ dabs,
                               const int* __restrict__ perm, cplx* __restrict__ T, int ldt) {
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const int dst = perm[j];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      cplx v = cmake(0.0, 0.0);
      if (i <= j) v = cscale(R[(size_t)j * ldr + i], 1.0 / dabs[i]);
      T[(size_t)dst * ldt + i] = v;
    }
  }
}
int build_T(cudaStream_t st, const cplx* R, int ldr, int n, const double* dabs, const int* perm, cplx* T, int ldt,
            int num_sms) {
  build_T_kernel<<<min(n, num_sms * 8), 256, 0, st>>>(R, ldr, n, dabs, perm, T, ldt);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// 32x32 tiles through shared memory so that both the read of Ul (transposed) and the writes are coalesced.
__global__ void __launch_bounds__(256)
loh_assemble_kernel(int n, const cplx* __restrict__ M1, const cplx* __restrict__ M2, const double* __restrict__ Dl,
                    const double* __restrict__ Dr, const cplx* __restrict__ Ul, cplx* __restrict__ inner,
                    cplx* __restrict__ rhs, double* __restrict__ drp_inv) {
  __shared__ cplx tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ntile = (n + 31) / 32;
  for (int t = blockIdx.x; t < ntile * ntile; t += gridDim.x) {
    const int bi = t % ntile, bj = t / ntile;
    // inner block (rows bi, cols bj)
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = bi * 32 + tx, c = bj * 32 + cc;
      if (r < n && c < n) {
        const double dl = Dl[r], dr = Dr[c];
        const double dlp = fmax(dl, 1.0), dlm = fmin(dl, 1.0), drp = fmax(dr, 1.0), drm = fmin(dr, 1.0);
        const cplx m1 = M1[(size_t)c * n + r], m2 = M2[(size_t)c * n + r];
        const double s1 = 1.0 / dlp / drp, s2 = dlm * drm;
        inner[(size_t)c * n + r] = cmake(m1.x * s1 + m2.x * s2, m1.y * s1 + m2.y * s2);
      }
    }
    // rhs block (rows bi, cols bj) = conj(Ul[cols bj, rows bi])^T / Dlp[row]
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = bj * 32 + tx, c = bi * 32 + cc;   // element of Ul
      tile[cc][tx] = (r < n && c < n) ? Ul[(size_t)c * n + r] : cmake(0.0, 0.0);
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = bi * 32 + tx, c = bj * 32 + cc;
      if (r < n && c < n) {
        cplx u = tile[tx][cc];                        // Ul[c, r]
        const double s = 1.0 / fmax(Dl[r], 1.0);
        rhs[(size_t)c * n + r] = cmake(u.x * s, -u.y * s);
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) drp_inv[i] = 1.0 / fmax(Dr[i], 1.0);
}
int loh_assemble(cudaStream_t st, int n, const cplx* M1, const cplx* M2, const double* Dl, const double* Dr,
                 const cplx* Ul, cplx* inner, cplx* rhs, double* drp_inv, int num_sms) {
  loh_assemble_kernel<<<num_sms * 4, 256, 0, st>>>(n, M1, M2, Dl, Dr, Ul, inner, rhs, drp_inv);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

__global__ void max_abs_diff_kernel(const cplx* __restrict__ A, const cplx* __restrict__ B, size_t count,
                                    unsigned long long* __restrict__ out) {
  double m = 0.0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x) {
    cplx d = csub(A[e], B[e]);
    m = fmax(m, sqrt(cabs2(d)));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));  // m >= 0: bit order == value order
}
int max_abs_diff(cudaStream_t st, const cplx* A, const cplx* B, size_t count, double* out, int num_sms) {
  CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double), st));
  max_abs_diff_kernel<<<num_sms * 4, 256, 0, st>>>(A, B, count, reinterpret_cast<unsigned long long*>(out));
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

__global__ void logdet_kernel(int n, const double* __restrict__ Dl, const double* __restrict__ Dr,
                              const double* __restrict__ dabs, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += log(fmax(Dl[i], 1.0)) + log(fmax(Dr[i], 1.0)) + log(dabs[i]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    out[0] = -t;
  }
}
int logdet_from_factors(cudaStream_t st, int n, const double* Dl, const double* Dr, const double* dabs, double* out) {
  logdet_kernel<<<1, 1024, 0, st>>>(n, Dl, Dr, dabs, out);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

__global__ void set_identity_kernel2(cplx* Q, int ldq, int n) {
  size_t tot = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(e % n), c = (int)(e / n);
    Q[(size_t)c * ldq + r] = cmake(r == c ? 1.0 : 0.0, 0.0);
  }
}
int set_identity(cudaStream_t st, cplx* Q, int ldq, int n, int num_sms) {
  set_identity_kernel2<<<num_sms * 4, 256, 0, st>>>(Q, ldq, n);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
__global__ void fill_ones_kernel(double* d, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = 1.0;
}
int fill_ones(cudaStream_t st, double* d, int n) {
  fill_ones_kernel<<<(n + 255) / 256, 256, 0, st>>>(d, n);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// out = alpha * ( f_a .* op_a(A) + f_b .* op_b(B) ),   f[i][j] = rowfac(i) * colfac(j) with factors derived from a
// real vector D by mode: 0 -> 1, 1 -> D, 2 -> 1/max(D,1), 3 -> min(D,1).  op = identity or conjugate transpose
// (staged through a shared-memory tile so both sides stay coalesced).  In place (out == A) is allowed when
// A is not transposed.  Used by the time-displaced Green's function path (linalg.jl:512-567).
__device__ __forceinline__ double ew_fac(const double* D, int mode, int i) {
  if (mode == 0 || D == nullptr) return 1.0;
  const double d = D[i];
  return mode == 1 ? d : (mode == 2 ? 1.0 / fmax(d, 1.0) : fmin(d, 1.0));
}
__global__ void __launch_bounds__(256) ew_combine_kernel(int n, EwTerm a, EwTerm b, double alpha, cplx* __restrict__ out) {
  __shared__ cplx tile[2][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ntile = (n + 31) / 32;
  for (int t = blockIdx.x; t < ntile * ntile; t += gridDim.x) {
    const int bi = t % ntile, bj = t / ntile;      // output rows bi*32.., cols bj*32..
    const EwTerm* terms[2] = {&a, &b};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const EwTerm& tm = *terms[k];
      if (tm.M != nullptr && tm.trans) {
        for (int cc = ty; cc < 32; cc += 8) {
          const int r = bj * 32 + tx, c = bi * 32 + cc;   // element M[r, c] feeds out[c, r]
          tile[k][cc][tx] = (r < n && c < n) ? tm.M[(size_t)c * n + r] : cmake(0.0, 0.0);
        }
      }
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = bi * 32 + tx, c = bj * 32 + cc;
      if (r < n && c < n) {
        double accx = 0.0, accy = 0.0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const EwTerm& tm = *terms[k];
          if (tm.M == nullptr) continue;
          cplx v;
          if (tm.trans) { v = tile[k][tx][cc]; v.y = -v.y; }
          else v = tm.M[(size_t)c * n + r];
          const double f = ew_fac(tm.rd, tm.rmode, r) * ew_fac(tm.cd, tm.cmode, c);
          accx += v.x * f; accy += v.y * f;
        }
        out[(size_t)c * n + r] = cmake(alpha * accx, alpha * accy);
      }
    }
    __syncthreads();
  }
}
int ew_combine(cudaStream_t st, int n, EwTerm a, EwTerm b, double alpha, cplx* out, int num_sms) {
  ew_combine_kernel<<<num_sms * 4, 256, 0, st>>>(n, a, b, alpha, out);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// Lower half of a matrix with the antiunitary flavour symmetry of the O(3) model from its upper half:
// G = [[A, B], [-conj(B), conj(A)]]  (oracle/experiments/antiunitary_symmetry.py).
__global__ void mirror_lower_half_kernel(cplx* __restrict__ G, int n) {
  const int h = n >> 1;
  const size_t total = (size_t)h * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e % h), c = (int)(e / h);
    const cplx v = G[(size_t)c * n + r];
    const bool left = c < h;
    G[(size_t)(left ? c + h : c - h) * n + r + h] = left ? cmake(v.x, -v.y) : cmake(-v.x, v.y);
  }
}
int mirror_lower_half(cudaStream_t st, cplx* G, int n, int num_sms) {
  mirror_lower_half_kernel<<<num_sms * 4, 256, 0, st>>>(G, n);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// out[0] = max over the upper half of |G[mirror(r,c)] - mirror value of G[r,c]|, out[1] = max |G|: how far a matrix is from
// the antiunitary flavour symmetry above (dqmc_set_greens / dqmc_calculate_greens_from decide with it whether the
// half-matrix shortcuts may be used on caller-supplied data).
__global__ void sym_violation_kernel(const cplx* __restrict__ G, int n, unsigned long long* __restrict__ out) {
  const int h = n >> 1;
  const size_t total = (size_t)h * n;
  double mv = 0.0, mg = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e % h), c = (int)(e / h);
    const cplx v = G[(size_t)c * n + r];
    const bool left = c < h;
    const cplx w = G[(size_t)(left ? c + h : c - h) * n + r + h];
    const cplx d = left ? cmake(w.x - v.x, w.y + v.y) : cmake(w.x + v.x, w.y - v.y);
    mv = fmax(mv, sqrt(cabs2(d)));
    mg = fmax(mg, fmax(sqrt(cabs2(v)), sqrt(cabs2(w))));
  }
  mv = warp_max(mv);
  mg = warp_max(mg);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(mv));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(mg));
  }
}
int sym_violation(cudaStream_t st, const cplx* G, int n, double* out2, int num_sms) {
  CUDA_TRY(cudaMemsetAsync(out2, 0, 2 * sizeof(double), st));
  sym_violation_kernel<<<num_sms * 4, 256, 0, st>>>(G, n, reinterpret_cast<unsigned long long*>(out2));
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

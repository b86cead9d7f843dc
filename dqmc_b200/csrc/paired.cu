#include "paired.cuh"

#define LAUNCHED() do { CUDA_TRY(cudaGetLastError()); g_launches++; } while (0)

__global__ void gather_interleave_cols_kernel(const cplx* __restrict__ in, int ldi, int n, const int* __restrict__ perm,
                                              cplx* __restrict__ out, int ldo) {
  const int h = n >> 1;
  for (int c = blockIdx.x; c < h; c += gridDim.x) {
    const cplx* src = in + (size_t)perm[c] * ldi;
    cplx* dst = out + (size_t)c * ldo;
    for (int r = threadIdx.x; r < n; r += blockDim.x) dst[r] = src[(r >> 1) + (r & 1) * h];
  }
}
int gather_interleave_cols(cudaStream_t st, const cplx* in, int ldi, int n, const int* perm, cplx* out, int ldo, int num_sms) {
  gather_interleave_cols_kernel<<<min(n / 2, num_sms * 8), 256, 0, st>>>(in, ldi, n, perm, out, ldo);
  LAUNCHED();
  return 0;
}

__global__ void set_identity_lint_kernel(cplx* __restrict__ out, int ldo, int n) {
  const int h = n >> 1;
  const size_t tot = (size_t)n * h;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e % n), c = (int)(e / n);
    out[(size_t)c * ldo + r] = cmake(r == 2 * c ? 1.0 : 0.0, 0.0);
  }
}
int set_identity_lint(cudaStream_t st, cplx* out, int ldo, int n, int num_sms) {
  set_identity_lint_kernel<<<num_sms * 4, 256, 0, st>>>(out, ldo, n);
  LAUNCHED();
  return 0;
}

// U = Q = (Q^H)^H.  With p(r) = 2r (r < h), 2(r-h)+1 (r >= h):  U[r, c] = conj(QH_int[p(c), p(r)]); for r < h the column p(r) = 2r of
// QH_int is column r of QHL, so the upper half is a conjugate transpose of QHL with permuted rows, and the lower half is its
// mirror image.  32x32 tiles through shared memory keep both sides coalesced.
__global__ void __launch_bounds__(256) build_U_paired_kernel(const cplx* __restrict__ QHL, int ldq, int n, cplx* __restrict__ U, int ldu) {
  __shared__ cplx tile[32][33];
  const int h = n >> 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nti = h / 32, ntj = n / 32;          // tiles over (r < h, c < n)
  for (int t = blockIdx.x; t < nti * ntj; t += gridDim.x) {
    const int bi = t % nti, bj = t / nti;
    // rows of QHL needed: p(c) for c in tile bj.  The tile lies entirely in one half (h % 32 == 0): p(c) = 2 (c mod h) + (c >= h)
    const int cbase = bj * 32, cpar = cbase >= h ? 1 : 0, cq = cbase - cpar * h;
    for (int cc = ty; cc < 32; cc += 8) {
      // element QHL[p(cbase + tx), bi*32 + cc]
      tile[cc][tx] = QHL[(size_t)(bi * 32 + cc) * ldq + 2 * (cq + tx) + cpar];
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = bi * 32 + tx, c = cbase + cc;
      const cplx v = tile[tx][cc];               // QHL[p(c), r]
      const cplx u = cmake(v.x, -v.y);           // U[r, c]
      U[(size_t)c * ldu + r] = u;
      // lower half: U[r+h, c+h] = conj(U[r, c]) (c < h);  U[r+h, c-h] = -conj(U[r, c]) (c >= h)
      if (c < h) U[(size_t)(c + h) * ldu + r + h] = cmake(u.x, -u.y);
      else U[(size_t)(c - h) * ldu + r + h] = cmake(-u.x, u.y);
    }
    __syncthreads();
  }
}
int build_U_paired(cudaStream_t st, const cplx* QHL, int ldq, int n, cplx* U, int ldu, int num_sms) {
  build_U_paired_kernel<<<num_sms * 4, 256, 0, st>>>(QHL, ldq, n, U, ldu);
  LAUNCHED();
  return 0;
}

__global__ void build_T_paired_kernel(const cplx* __restrict__ RL, int ldr, int n, const double* __restrict__ dabs,
                                      const int* __restrict__ perm, cplx* __restrict__ T, int ldt) {
  const int h = n >> 1;
  for (int c = blockIdx.x; c < h; c += gridDim.x) {
    const int dst = perm[c];
    const cplx* src = RL + (size_t)c * ldr;
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
      cplx e = cmake(0.0, 0.0), o = e;
      if (i <= c) {
        const double s = 1.0 / dabs[i];
        e = cscale(src[2 * i], s);
        o = cscale(src[2 * i + 1], s);
      }
      T[(size_t)dst * ldt + i] = e;
      T[(size_t)dst * ldt + i + h] = o;
      T[(size_t)(dst + h) * ldt + i] = cmake(-o.x, o.y);
      T[(size_t)(dst + h) * ldt + i + h] = cmake(e.x, -e.y);
    }
  }
}
int build_T_paired(cudaStream_t st, const cplx* RL, int ldr, int n, const double* dabs, const int* perm, cplx* T, int ldt, int num_sms) {
  build_T_paired_kernel<<<min(n / 2, num_sms * 8), 256, 0, st>>>(RL, ldr, n, dabs, perm, T, ldt);
  LAUNCHED();
  return 0;
}

__global__ void mirror_right_half_kernel(cplx* __restrict__ M, int ldm, int n) {
  const int h = n >> 1;
  const size_t total = (size_t)h * h;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % h), c = (int)(e / h);
    const cplx t = M[(size_t)c * ldm + i], b = M[(size_t)c * ldm + i + h];
    M[(size_t)(c + h) * ldm + i] = cmake(-b.x, b.y);
    M[(size_t)(c + h) * ldm + i + h] = cmake(t.x, -t.y);
  }
}
int mirror_right_half(cudaStream_t st, cplx* M, int ldm, int n, int num_sms) {
  mirror_right_half_kernel<<<num_sms * 4, 256, 0, st>>>(M, ldm, n);
  LAUNCHED();
  return 0;
}

// 32x32 tiles; the rhs part reads Ul transposed through shared memory.  Output rows are pair-interleaved: natural row r of a
// tile in the upper (lower) half goes to row 2r (2(r-h)+1), a stride-2 store.
__global__ void __launch_bounds__(256)
loh_assemble_paired_kernel(int n, const cplx* __restrict__ M1, const cplx* __restrict__ M2, const double* __restrict__ Dl,
                           const double* __restrict__ Dr, const cplx* __restrict__ Ul, cplx* __restrict__ inner,
                           cplx* __restrict__ rhs, double* __restrict__ drp_inv) {
  __shared__ cplx tile[32][33];
  const int h = n >> 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nti = n / 32, ntj = h / 32;
  for (int t = blockIdx.x; t < nti * ntj; t += gridDim.x) {
    const int bi = t % nti, bj = t / nti;
    const int r = bi * 32 + tx;
    const int ro = r < h ? 2 * r : 2 * (r - h) + 1;
    const double dl = Dl[r], dlp = fmax(dl, 1.0), dlm = fmin(dl, 1.0);
    for (int cc = ty; cc < 32; cc += 8) {
      const int c = bj * 32 + cc;
      const double dr = Dr[c], drp = fmax(dr, 1.0), drm = fmin(dr, 1.0);
      const cplx m1 = M1[(size_t)c * n + r], m2 = M2[(size_t)c * n + r];
      const double s1 = 1.0 / dlp / drp, s2 = dlm * drm;
      inner[(size_t)c * n + ro] = cmake(m1.x * s1 + m2.x * s2, m1.y * s1 + m2.y * s2);
    }
    for (int cc = ty; cc < 32; cc += 8) {
      const int ur = bj * 32 + tx, uc = bi * 32 + cc;   // element Ul[ur, uc] feeds rhs[uc, ur]
      tile[cc][tx] = Ul[(size_t)uc * n + ur];
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
      const int c = bj * 32 + cc;
      const cplx u = tile[tx][cc];                      // Ul[c, r]
      const double s = 1.0 / dlp;
      rhs[(size_t)c * n + ro] = cmake(u.x * s, -u.y * s);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) drp_inv[i] = 1.0 / fmax(Dr[i], 1.0);
}
int loh_assemble_paired(cudaStream_t st, int n, const cplx* M1L, const cplx* M2L, const double* Dl, const double* Dr, const cplx* Ul,
                        cplx* innerL, cplx* rhsL, double* drp_inv, int num_sms) {
  loh_assemble_paired_kernel<<<num_sms * 4, 256, 0, st>>>(n, M1L, M2L, Dl, Dr, Ul, innerL, rhsL, drp_inv);
  LAUNCHED();
  return 0;
}

__global__ void expand_R_paired_kernel(const cplx* __restrict__ RL, int ldr, int n, cplx* __restrict__ Rf, int ldf) {
  const int h = n >> 1;
  for (int c = blockIdx.x; c < h; c += gridDim.x) {
    const cplx* src = RL + (size_t)c * ldr;
    cplx* d0 = Rf + (size_t)(2 * c) * ldf;
    cplx* d1 = d0 + ldf;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const bool nz = (r >> 1) <= c;
      const cplx v = nz ? src[r] : cmake(0.0, 0.0);
      d0[r] = v;
      // partner: rows (2q, 2q+1) = (-conj(v[2q+1]), conj(v[2q])): my value goes to the other row of the pair
      d1[r ^ 1] = (r & 1) ? cmake(-v.x, v.y) : cmake(v.x, -v.y);
    }
  }
}
int expand_R_paired(cudaStream_t st, const cplx* RL, int ldr, int n, cplx* Rfull, int ldf, int num_sms) {
  expand_R_paired_kernel<<<min(n / 2, num_sms * 8), 256, 0, st>>>(RL, ldr, n, Rfull, ldf);
  LAUNCHED();
  return 0;
}

__global__ void uninterleave_rows_kernel(const cplx* __restrict__ in, int ldi, int n, const double* __restrict__ rowscale,
                                         cplx* __restrict__ out, int ldo) {
  const int h = n >> 1;
  for (int c = blockIdx.x; c < h; c += gridDim.x) {
    const cplx* src = in + (size_t)c * ldi;
    cplx* dst = out + (size_t)c * ldo;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const int nat = (r >> 1) + (r & 1) * h;
      const cplx v = src[r];
      dst[nat] = rowscale ? cscale(v, rowscale[nat]) : v;
    }
  }
}
int uninterleave_rows(cudaStream_t st, const cplx* in, int ldi, int n, const double* rowscale, cplx* out, int ldo, int num_sms) {
  uninterleave_rows_kernel<<<min(n / 2, num_sms * 8), 256, 0, st>>>(in, ldi, n, rowscale, out, ldo);
  LAUNCHED();
  return 0;
}

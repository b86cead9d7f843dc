#include "blockops.cuh"

__device__ __forceinline__ void cp_async_16(cplx* smem_dst, const cplx* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

// TMA bulk copies (cp.async.bulk, SASS UBLKCP): a column of the column-major matrix is one contiguous run of n x 16 bytes, so the
// COLS pass moves its panel with one bulk copy per column in each direction (completion on an mbarrier for the loads, a bulk
// group for the stores) instead of n/256 16-byte cp.async per thread and column.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "BLK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni BLK_DONE;\n"
      "bra.uni BLK_WAIT;\n"
      "BLK_DONE:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
}

// One CTA stages `nvec` vectors of length n (columns for COLS, rows for ROWS) in shared memory,
// applies every step of the chain in place, and writes them back.
template <bool ROWS>
__global__ void __launch_bounds__(256) apply_chain_kernel(cplx* __restrict__ mat, int n, int ld, int nvec_cta,
                                                          Chain chain, const double* __restrict__ hsfield,
                                                          int nsites, double lam_dtau,
                                                          const double* __restrict__ colscale,
                                                          double* __restrict__ colnorm2, int nvtot, int mirror) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ldx = ROWS ? n + 1 : n;
  cplx* x = reinterpret_cast<cplx*>(smem_raw);
  double4* tab = reinterpret_cast<double4*>(x + (size_t)nvec_cta * ldx);
  __shared__ double red[8];

  const int v0 = blockIdx.x * nvec_cta;
  const int nvec = min(nvec_cta, nvtot - v0);   // nvtot vectors in all (n, or the n/2 left-half columns of a symmetric matrix)
  const int tid = threadIdx.x;

  // ---- global -> shared.  COLS: one TMA bulk copy per column; ROWS: 16-byte cp.async copies, all in flight at once
  __shared__ __align__(8) unsigned long long ld_bar;
  if (!ROWS) {
    if (tid == 0) mbar_init(&ld_bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&ld_bar, (unsigned)(nvec * n * sizeof(cplx)));
      for (int v = 0; v < nvec; ++v) bulk_g2s(x + (size_t)v * ldx, mat + (size_t)(v0 + v) * ld, (unsigned)(n * sizeof(cplx)), &ld_bar);
    }
    mbar_wait(&ld_bar, 0);
  } else {
    // element (row v0+v, col j): v fastest so that each column contributes nvec*16 contiguous bytes
    const int tot = nvec * n;
    for (int e = tid; e < tot; e += blockDim.x) {
      int v = e % nvec, j = e / nvec;
      cp_async_16(x + (size_t)v * ldx + j, mat + (size_t)j * ld + v0 + v);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // the 4x4 block of the NEXT stored-operator step is fetched (L2 latency) while the current step runs; with n/4 <= 256
  // blocks every thread owns exactly one block per step
  cplx mnx[4][4];
  int inx[4] = {0, 0, 0, 0};
  auto prefetch_block = [&](int st) {
    if (st >= chain.nsteps) return;
    const ChainStep& s = chain.s[st];
    if (s.kind != 0 || tid >= s.nblk) return;
    const cplx* vp = s.val + 16 * (size_t)tid;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        cplx t = (s.mode == OP_N || s.mode == OP_J) ? vp[4 * r + c] : vp[4 * c + r];
        if (s.mode == OP_C || s.mode == OP_J) t.y = -t.y;
        mnx[r][c] = t;
      }
#pragma unroll
    for (int k = 0; k < 4; ++k) inx[k] = s.idx[4 * tid + k];
  };
  prefetch_block(0);
  for (int st = 0; st < chain.nsteps; ++st) {
    const ChainStep s = chain.s[st];
    if (s.kind == 0) {
      for (int b = tid; b < s.nblk; b += blockDim.x) {
        int i0, i1, i2, i3;
        cplx m[4][4];
        if (b == tid) {
          i0 = inx[0]; i1 = inx[1]; i2 = inx[2]; i3 = inx[3];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) m[r][c] = mnx[r][c];
          prefetch_block(st + 1);
        } else {
          i0 = s.idx[4 * b + 0]; i1 = s.idx[4 * b + 1]; i2 = s.idx[4 * b + 2]; i3 = s.idx[4 * b + 3];
          const cplx* vp = s.val + 16 * (size_t)b;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              cplx t = (s.mode == OP_N || s.mode == OP_J) ? vp[4 * r + c] : vp[4 * c + r];
              if (s.mode == OP_C || s.mode == OP_J) t.y = -t.y;
              m[r][c] = t;
            }
        }
        for (int v = 0; v < nvec; ++v) {
          cplx* xv = x + (size_t)v * ldx;
          cplx a0 = xv[i0], a1 = xv[i1], a2 = xv[i2], a3 = xv[i3];
          cplx y[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            cplx acc = cmul(m[r][0], a0);
            cfma(acc, m[r][1], a1);
            cfma(acc, m[r][2], a2);
            cfma(acc, m[r][3], a3);
            y[r] = acc;
          }
          xv[i0] = y[0]; xv[i1] = y[1]; xv[i2] = y[2]; xv[i3] = y[3];
        }
      }
      if (tid >= s.nblk) prefetch_block(st + 1);
      __syncthreads();
    } else {
      // interaction exponential of one time slice (interactions.jl:35-88):
      // C = cosh(lam*dtau*|phi|), S = (i phi2 - phi1) sign sinh/|phi|, R = -phi3 sign sinh/|phi|
      prefetch_block(st + 1);
      for (int i = tid; i < nsites; i += blockDim.x) {
        const double* h = hsfield + 3 * ((size_t)i + (size_t)nsites * s.slice);
        double p1 = h[0], p2 = h[1], p3 = h[2];
        double nrm = sqrt(p1 * p1 + p2 * p2 + p3 * p3);
        double sh = s.sign * sinh(lam_dtau * nrm) / nrm;
        double sim = p2 * sh;
        if (s.mode == OP_T || s.mode == OP_J) sim = -sim;  // E^T = conj(E) (E is Hermitian)
        tab[i] = make_double4(cosh(lam_dtau * nrm), -p1 * sh, sim, -p3 * sh);
      }
      __syncthreads();
      for (int i = tid; i < nsites; i += blockDim.x) {
        double4 t = tab[i];
        const double C = t.x, R = t.w;
        const cplx S = cmake(t.y, t.z), cS = cmake(t.y, -t.z);
        for (int v = 0; v < nvec; ++v) {
          cplx* xv = x + (size_t)v * ldx;
          cplx a0 = xv[i], a1 = xv[i + nsites], a2 = xv[i + 2 * nsites], a3 = xv[i + 3 * nsites];
          cplx y0 = cscale(a0, C); cfma(y0, S, a1); y0.x = fma(R, a3.x, y0.x); y0.y = fma(R, a3.y, y0.y);
          cplx y1 = cscale(a1, C); cfma(y1, cS, a0); y1.x = fma(-R, a2.x, y1.x); y1.y = fma(-R, a2.y, y1.y);
          cplx y2 = cscale(a2, C); cfma(y2, cS, a3); y2.x = fma(-R, a1.x, y2.x); y2.y = fma(-R, a1.y, y2.y);
          cplx y3 = cscale(a3, C); cfma(y3, S, a2); y3.x = fma(R, a0.x, y3.x); y3.y = fma(R, a0.y, y3.y);
          xv[i] = y0; xv[i + nsites] = y1; xv[i + 2 * nsites] = y2; xv[i + 3 * nsites] = y3;
        }
      }
      __syncthreads();
    }
  }

  // ---- shared -> global (optionally scale columns and emit their squared norms)
  if (!ROWS && colscale == nullptr && colnorm2 == nullptr) {
    // plain write-back: TMA bulk stores straight from the panel (the generic-proxy writes of the last step must be made visible
    // to the async proxy first)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      for (int v = 0; v < nvec; ++v) bulk_s2g(mat + (size_t)(v0 + v) * ld, x + (size_t)v * ldx, (unsigned)(n * sizeof(cplx)));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (mirror) {
      // the matrix has the antiunitary flavour symmetry [[A, B], [-conj(B), conj(A)]] and only its left-half columns were
      // processed: column c + n/2 is the partner (-conj(bottom half); conj(top half)) of column c
      const int h = n >> 1;
      for (int v = 0; v < nvec; ++v) {
        const cplx* xv = x + (size_t)v * ldx;
        cplx* dst = mat + (size_t)(v0 + v + h) * ld;
        for (int j = tid; j < n; j += blockDim.x) {
          const cplx t = xv[j < h ? j + h : j - h];
          dst[j] = j < h ? cmake(-t.x, t.y) : cmake(t.x, -t.y);
        }
      }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (!ROWS) {
    for (int v = 0; v < nvec; ++v) {
      const double sc = colscale ? colscale[v0 + v] : 1.0;
      double acc = 0.0;
      for (int j = tid; j < n; j += blockDim.x) {
        cplx t = cscale(x[(size_t)v * ldx + j], sc);
        acc += cabs2(t);
        mat[(size_t)(v0 + v) * ld + j] = t;
      }
      if (colnorm2) {
        acc = warp_sum(acc);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
          double t = 0.0;
          for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
          colnorm2[v0 + v] = t;
        }
      }
    }
  } else {
    const int tot = nvec * n;
    const int h = n >> 1;
    for (int e = tid; e < tot; e += blockDim.x) {
      int v = e % nvec, j = e / nvec;
      const cplx t = x[(size_t)v * ldx + j];
      mat[(size_t)j * ld + v0 + v] = t;
      // upper-half rows only were processed: row r + n/2 is (-conj(right half), conj(left half)) of row r
      if (mirror) mat[(size_t)(j < h ? j + h : j - h) * ld + v0 + v + h] = j < h ? cmake(t.x, -t.y) : cmake(-t.x, t.y);
    }
  }
}

int launch_apply_chain(bool rows, cplx* mat, int n, int ld, const Chain& chain, const double* hsfield,
                       int nsites, double lam_dtau, const double* colscale, double* colnorm2,
                       int num_sms, cudaStream_t stream, int nvtot, int mirror) {
  if (nvtot <= 0 || nvtot > n) nvtot = n;
  if (mirror) {
    if (n % 2 != 0 || colscale != nullptr || colnorm2 != nullptr) { snprintf(g_errbuf, sizeof(g_errbuf), "apply_chain: mirror mode needs even n and a plain write-back"); return -1; }
    nvtot = n / 2;
  }
  const size_t smem_cap = 200 * 1024;
  const size_t tab_bytes = sizeof(double4) * (size_t)nsites;
  const int ldx = rows ? n + 1 : n;
  int nvec_max = (int)((smem_cap - tab_bytes) / (sizeof(cplx) * (size_t)ldx));
  if (nvec_max < 1) {
    snprintf(g_errbuf, sizeof(g_errbuf), "apply_chain: n=%d too large for one shared-memory vector", n);
    return -1;
  }
  int nvec;
  if (rows) {
    nvec = 8;                                   // 128 B contiguous per column
    while (nvec > 1 && ((nvtot + nvec - 1) / nvec < num_sms / 2 || nvec > nvec_max)) nvec >>= 1;
  } else {
    nvec = (nvtot + num_sms - 1) / num_sms;
    if (nvec > nvec_max) nvec = nvec_max;
    if (nvec < 1) nvec = 1;
  }
  const int grid = (nvtot + nvec - 1) / nvec;
  const size_t smem = sizeof(cplx) * (size_t)nvec * ldx + tab_bytes;
  static SmemMemo memo_cols, memo_rows;
  if (ensure_max_dynamic_smem(apply_chain_kernel<false>, memo_cols, nullptr)) return -1;
  if (ensure_max_dynamic_smem(apply_chain_kernel<true>, memo_rows, nullptr)) return -1;
  if (rows)
    apply_chain_kernel<true><<<grid, 256, smem, stream>>>(mat, n, ld, nvec, chain, hsfield, nsites, lam_dtau, colscale, colnorm2, nvtot, mirror);
  else
    apply_chain_kernel<false><<<grid, 256, smem, stream>>>(mat, n, ld, nvec, chain, hsfield, nsites, lam_dtau, colscale, colnorm2, nvtot, mirror);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// Device helpers shared by the two local-update kernels (local_updates.cu: per-site lookahead; local_updates_blk.cu: block
// lookahead): sentinel loads/stores of the published update factors, the proposal evaluation (randuniform,
// calc_boson_action_diff, interaction_matrix_exp_op!) and small complex helpers.
#pragma once
#include "local_updates.cuh"

#define LU_SENT 0xFFFFFFFFFFFFFFFFull

struct Prep {
  double nw[3];      // proposed field value
  double e_dS;       // exp(-dS)
  double mlog;       // -log(exp(-dS))   (local_updates.jl:34)
  double u3;         // the accept draw (used only if p_acc <= 1)
  cplx D[16];        // Delta = e^{+dtau V(old)} e^{-dtau V(new)} - 1, row-major
};

__device__ __forceinline__ cplx ld_valid(const cplx* p) {
  unsigned long long x, y;
  do {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
  } while (x == LU_SENT || y == LU_SENT);
  return make_double2(__longlong_as_double((long long)x), __longlong_as_double((long long)y));
}
// Publishing store of one element of the update factors.  LU_WEAK_PUB: a plain 16-byte store to L2 (one transaction, so a reader
// sees the sentinel or the value); consumers spin on the data, so no ordering is needed -- and a CTA barrier that follows does not
// have to wait for a strong (volatile) store to be acknowledged by L2.
__device__ __forceinline__ void st_pub(cplx* p, cplx v) {
#ifdef LU_WEAK_PUB
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#else
  asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#endif
}
__device__ __forceinline__ void st_sent(cplx* p) {
#ifdef LU_WEAK_PUB
  asm volatile("st.global.cg.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(LU_SENT) : "memory");
#else
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(LU_SENT) : "memory");
#endif
}
// L2 load issued in program order relative to the other volatile asm statements (no memory clobber: ordinary accesses may move)
__device__ __forceinline__ cplx ld_cg_issue(const cplx* p) {
  cplx v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ cplx ldcg2(const cplx* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

__device__ __forceinline__ void cp_async16(cplx* smem_dst, const cplx* gmem_src, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 16 : 0;   // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}

// 4x4 complex e^{-power*dtau*V(op)} (interactions.jl:102-141), element (r,c):  [C S 0 R; cS C -R 0; 0 -R C cS; R 0 S C]
__device__ __forceinline__ cplx evop_elem(int r, int c, double C, cplx S, double R) {
  if (r == c) return cmake(C, 0.0);
  switch (r * 4 + c) {
    case 1: case 14: return S;
    case 4: case 11: return cconj(S);
    case 3: case 12: return cmake(R, 0.0);
    case 6: case 9: return cmake(-R, 0.0);
    default: return cmake(0.0, 0.0);
  }
}

__device__ __forceinline__ cplx det3(cplx a, cplx b, cplx c, cplx d, cplx e, cplx f, cplx g, cplx h, cplx i) {
  cplx t1 = csub(cmul(e, i), cmul(f, h));
  cplx t2 = csub(cmul(d, i), cmul(f, g));
  cplx t3 = csub(cmul(d, h), cmul(e, g));
  return cadd(csub(cmul(a, t1), cmul(b, t2)), cmul(c, t3));
}

// rare path (|x| > 1), kept out of line so that it does not sit in the instruction stream of the site loop
static __device__ __noinline__ void cosh_sinhc_large(double z, double* ch, double* shc) {
  const double x = sqrt(z), em1 = expm1(x), q = em1 / (em1 + 1.0);
  *ch = 1.0 + 0.5 * em1 * q;
  *shc = 0.5 * (em1 + q) / x;
}

// cosh(x) and sinh(x)/x as functions of z = x^2 (Taylor series, |x| <= 1: truncation < 1e-18), so that neither a square
// root nor a division by |phi| is needed; larger arguments fall back to expm1.
__device__ __forceinline__ void cosh_sinhc(double z, double* ch, double* shc) {
  if (z <= 1.0) {
    double c = 1.0 / 2432902008176640000.0, s = 1.0 / 51090942171709440000.0;   // 1/20!, 1/21!
    c = fma(c, z, 1.0 / 6402373705728000.0);  s = fma(s, z, 1.0 / 121645100408832000.0);   // 18!, 19!
    c = fma(c, z, 1.0 / 20922789888000.0);    s = fma(s, z, 1.0 / 355687428096000.0);      // 16!, 17!
    c = fma(c, z, 1.0 / 87178291200.0);       s = fma(s, z, 1.0 / 1307674368000.0);        // 14!, 15!
    c = fma(c, z, 1.0 / 479001600.0);         s = fma(s, z, 1.0 / 6227020800.0);           // 12!, 13!
    c = fma(c, z, 1.0 / 3628800.0);           s = fma(s, z, 1.0 / 39916800.0);             // 10!, 11!
    c = fma(c, z, 1.0 / 40320.0);             s = fma(s, z, 1.0 / 362880.0);               // 8!, 9!
    c = fma(c, z, 1.0 / 720.0);               s = fma(s, z, 1.0 / 5040.0);                 // 6!, 7!
    c = fma(c, z, 1.0 / 24.0);                s = fma(s, z, 1.0 / 120.0);                  // 4!, 5!
    c = fma(c, z, 0.5);                       s = fma(s, z, 1.0 / 6.0);                    // 2!, 3!
    *ch = fma(c, z, 1.0);
    *shc = fma(s, z, 1.0);
  } else {
    cosh_sinhc_large(z, ch, shc);
  }
}

// One warp evaluates the proposal at `site`: proposal draws uw[off..off+2], accept draw uw[off+3] (uw = this slice's
// window of the uniform stream in shared memory).  prev_site / pn*: a site whose field value must be read as pn*
// instead of fs[] (or -1).  tn = phi(l+1) + phi(l-1) per site, nbr = spatial neighbour table, both in shared memory.
__device__ __forceinline__ void do_prep(const LUArgs& a, const double* fs, const double* tn, const int* nbr, const double* uw,
                                        int off, int navail, int site, int prev_site, double pn1, double pn2, double pn3,
                                        Prep* out, int* exhausted) {
  const int lane = threadIdx.x & 31;
  double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
  if (off + 4 <= navail) { u0 = uw[off]; u1 = uw[off + 1]; u2 = uw[off + 2]; u3 = uw[off + 3]; }
  else *exhausted = 1;
  const double o1 = fs[3 * site], o2 = fs[3 * site + 1], o3 = fs[3 * site + 2];
  // randuniform (dqmc_framework.jl:628): -b + 2*b*rand(); no FMA contraction so the field stays bit-identical
  const double b2 = __dmul_rn(2.0, a.box);
  const double n1 = __dadd_rn(o1, __dadd_rn(-a.box, __dmul_rn(b2, u0)));
  const double n2 = __dadd_rn(o2, __dadd_rn(-a.box, __dmul_rn(b2, u1)));
  const double n3 = __dadd_rn(o3, __dadd_rn(-a.box, __dmul_rn(b2, u2)));
  // calc_boson_action_diff (action.jl:57-101)
  const double d1 = n1 - o1, d2 = n2 - o2, d3 = n3 - o3;
  const double osq = o1 * o1 + o2 * o2 + o3 * o3, nsq = n1 * n1 + n2 * n2 + n3 * n3;
  const double sq_diff = nsq - osq, pow4_diff = nsq * nsq - osq * osq;
  double dS;
  if (!a.edrun) {
    const double t1 = tn[3 * site], t2 = tn[3 * site + 1], t3 = tn[3 * site + 2];
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int j = nbr[4 * site + nb];
      const bool sub = (j == prev_site);
      s1 += sub ? pn1 : fs[3 * j];
      s2 += sub ? pn2 : fs[3 * j + 1];
      s3 += sub ? pn3 : fs[3 * j + 2];
    }
    dS = a.inv_dtau_c2 * (sq_diff - (t1 * d1 + t2 * d2 + t3 * d3));
    dS += 0.5 * a.dtau * (4.0 * sq_diff - 2.0 * (s1 * d1 + s2 * d2 + s3 * d3));
    dS += a.dtau * (0.5 * a.r * sq_diff + 0.25 * a.u * pow4_diff);
  } else {
    dS = a.dtau * (0.5 * a.r * sq_diff);
  }
  const double e_dS = exp(-dS);
  // interaction_matrix_exp_op!: old value with power -1, new value with power +1;  sinh(x)/|phi| = lam*dtau * sinh(x)/x
  const double l2 = a.lam_dtau * a.lam_dtau;
  double C1, q1, C2, q2;
#ifdef LU_PREP_SPLIT_LANES
  {   // lanes 0-15 evaluate the series for the old value, lanes 16-31 for the new one; one exchange instead of a second pass
    double Cm, qm;
    cosh_sinhc(l2 * (lane < 16 ? osq : nsq), &Cm, &qm);
    const double Co = __shfl_xor_sync(0xffffffffu, Cm, 16), qo = __shfl_xor_sync(0xffffffffu, qm, 16);
    C1 = lane < 16 ? Cm : Co; q1 = lane < 16 ? qm : qo;
    C2 = lane < 16 ? Co : Cm; q2 = lane < 16 ? qo : qm;
  }
#else
  cosh_sinhc(l2 * osq, &C1, &q1);
  cosh_sinhc(l2 * nsq, &C2, &q2);
#endif
  const double sh1 = -a.lam_dtau * q1, sh2 = a.lam_dtau * q2;
  const cplx S1 = cmake(-o1 * sh1, o2 * sh1), S2 = cmake(-n1 * sh2, n2 * sh2);
  const double R1 = -o3 * sh1, R2 = -n3 * sh2;
  // E = C*1 + Y(S,R),  Y = [0 S 0 R; cS 0 -R 0; 0 -R 0 cS; R 0 S 0],  so  E1 E2 - 1 = (C1 C2 - 1) + C1 Y2 + C2 Y1 + Y1 Y2 with
  // Y1 Y2 = diag(a, ca, ca, a) + b (e02 + e31) - cb (e13 + e20),  a = S1 cS2 + R1 R2,  b = R1 S2 - S1 R2.
  if (lane < 16) {
    const int r = lane >> 2, c = lane & 3;
    const cplx av = cmake(S1.x * S2.x + S1.y * S2.y + R1 * R2, S1.y * S2.x - S1.x * S2.y);
    const cplx bv = cmake(R1 * S2.x - S1.x * R2, R1 * S2.y - S1.y * R2);
    const cplx ys = cmake(C1 * S2.x + C2 * S1.x, C1 * S2.y + C2 * S1.y);     // S-type entry of C1 Y2 + C2 Y1
    const double yr = C1 * R2 + C2 * R1;                                      // R-type entry
    // kind per (r,c): 0 a, 1 conj a, 2 S, 3 conj S, 4 R, 5 -R, 6 b, 7 -conj b
    const unsigned long long kinds = 0x264315775134620ull;   // nibble (r*4+c): see table below
    // (0,0)0 (0,1)2 (0,2)6 (0,3)4 | (1,0)3 (1,1)1 (1,2)5 (1,3)7 | (2,0)7 (2,1)5 (2,2)1 (2,3)3 | (3,0)4 (3,1)6 (3,2)2 (3,3)0
    const int kind = (int)((kinds >> (4 * lane)) & 7ull);
    cplx v;
    switch (kind) {
      case 0: v = cmake(C1 * C2 - 1.0 + av.x, av.y); break;
      case 1: v = cmake(C1 * C2 - 1.0 + av.x, -av.y); break;
      case 2: v = ys; break;
      case 3: v = cconj(ys); break;
      case 4: v = cmake(yr, 0.0); break;
      case 5: v = cmake(-yr, 0.0); break;
      case 6: v = bv; break;
      default: v = cmake(-bv.x, bv.y); break;
    }
    (void)r; (void)c;
    out->D[lane] = v;
  }
  if (lane == 0) {
    out->nw[0] = n1; out->nw[1] = n2; out->nw[2] = n3;
    // the reference accumulates -log(exp(-dS)) (local_updates.jl:34) = dS up to one rounding of exp/log
    out->e_dS = e_dS; out->mlog = dS; out->u3 = u3;
  }
}

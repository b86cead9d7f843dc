// C ABI of libdqmc_b200: context, operator ingestion, the propagate state machine of the reference
// (stack.jl:391-499) driving the CUDA kernels, and the sweep entry points.  See include/dqmc_b200.h.
#include <dlfcn.h>

#include <algorithm>
#include <map>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/dqmc_b200.h"
#include "blockops.cuh"
#include "common.cuh"
#include "local_updates.cuh"
#include "misc.cuh"
#include "paired.cuh"
#include "qr.cuh"
#include "zgemm.cuh"

char g_errbuf[512] = {0};
long long g_launches = 0;

enum { F_HBH = 0, F_HA, F_HBH_MU, F_HBHINV, F_HAINV, F_MUINV_HBHINV, F_HAH, F_HAHINV, F_COUNT };
enum { TM_WRAP = 0, TM_LOCAL, TM_UDT, TM_GREENS, TM_SWEEP, TM_COUNT };

struct HostCSC {
  bool set = false;
  int n = 0;
  std::vector<int64_t> colptr, rowval;
  std::vector<cplx> nz;
};

struct TimerRec { int cat; cudaEvent_t a, b; };
// one stabilization step of propagate (a fixed (slice, direction) always touches the same slabs with the same kernels), captured
// once as a CUDA graph and replayed: ~300 launches on two streams become one graph launch
struct BlockGraph { cudaGraphExec_t exec; long long launches; int slice_after, dir_after; bool symG_after; };

struct dqmc_ctx {
  dqmc_params p;
  int n, N, M, sm, nel, kmax;
  int num_sms;
  cudaStream_t st;
  QrAsync qra;
  bool lookahead;
  char err[512];
  // state (1-based like the reference)
  int current_slice, direction;
  // device buffers
  cplx *G, *Gtmp, *u_stack, *t_stack;
  double* d_stack;
  cplx *gb_G, *gb_u_stack, *gb_t_stack;   // global-update backups (stack.jl:134-143), allocated on first use
  double *gb_d_stack, *gb_hs;
  double gb_log_det;
  double* d_action;                       // [0] result, [1..] per-block partials
  cplx *Ul, *Ur, *Tl, *Tr;
  double *Dl, *Dr;
  cplx* W[5];
  cplx *tau, *tfac, *trsm_work;
  double *dabs, *drp_inv, *colnorm;
  int *perm, *pos;
  double* hs;
  double* hs_bak;
  int* nbr;
  cplx *At, *Bm;
  double* unif;
  long long unif_cap, unif_n, unif_pos_bound;   // unif_pos_bound: host-side upper bound of the device stream position
  long long* d_pos;
  long long* d_acc;
  double* d_dS;
  int* d_flags;
  unsigned int* d_bar;
  long long* d_prof;
  bool lu_prof;
  double* d_logdet;
  double* d_check;     // [0] running max, [1] scratch
  // time-displaced Green's functions (fermion_measurements.jl:1343-1541), allocated on first use
  cplx *Gt0, *G0t;                 // [M][n*n]
  cplx *td_u[4], *td_t[4];         // UDT chains BT0Inv, BBetaT, BT0, BBetaTInv: [K][n*n]
  double* td_d[4];                 // [K][n]
  cplx* td_eye; double* td_ones;
  bool have_nbr, ops_ready;
  // antiunitary flavour symmetry X = [[A, B], [-conj(B), conj(A)]]: the half-matrix shortcuts are used
  // only while it is known to hold.  sym_model: every operator handed to dqmc_set_operator has it (checked on the host);
  // sym_G: the current G has it (true after every calculate_greens of a symmetric model, measured for caller-supplied G).
  bool sym_lu_opt, sym_greens_opt, sym_model, sym_G;
  bool wrap_sym_opt;                // half-matrix wrap (DQMC_WRAP_SYM=0: off)
  bool paired_opt;                  // half-matrix stabilization (paired Householder QR) for symmetric models (DQMC_PAIRED=0: off)
  double* d_sym;                    // [2] scratch of sym_violation
  HostCSC csc[DQMC_OP_COUNT];
  QuadOp fop[F_COUNT];
  int lu_grid, lu_rpc;
  bool udt_qrcp_sweep;              // DQMC_UDT=qrcp: the sweep's UDTs use the column-pivoted QR too (experiments / bisecting)
  bool lu_blk;                      // block-lookahead local-update kernel (default when rows per CTA <= 16; DQMC_LU_KERNEL=site: the older one)
  int lu_bar_mode, lu_bar_parity;   // grid barrier of the local-update kernel: 1 = monotonic counters alternating per launch
  int timing;                       // 0 off, 1 sweep timer only (CUDA graphs stay on), 2 all phase timers (graphs off)
  bool use_graphs, capturing;
  std::map<long long, BlockGraph> graphs;
  std::map<long long, int> graph_seen;
  std::vector<TimerRec> trecs;
  std::vector<cudaEvent_t> evpool;
  double tacc[TM_COUNT];
};

#define CTX_FAIL(ctx, ...)                                   \
  do {                                                       \
    snprintf(g_errbuf, sizeof(g_errbuf), __VA_ARGS__);       \
    if (ctx) memcpy((ctx)->err, g_errbuf, sizeof(g_errbuf)); \
    return -1;                                               \
  } while (0)
#define TRY(ctx, expr)                                         \
  do {                                                         \
    if ((expr) != 0) {                                         \
      if (ctx) memcpy((ctx)->err, g_errbuf, sizeof(g_errbuf)); \
      return -1;                                               \
    }                                                          \
  } while (0)
#define CU(ctx, expr)                                                                                        \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) CTX_FAIL(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

static const cplx ONE = {1.0, 0.0}, ZERO = {0.0, 0.0};

// ---------------------------------------------------------------------------------------------- timers
static cudaEvent_t ev_get(dqmc_ctx* c) {
  if (!c->evpool.empty()) { cudaEvent_t e = c->evpool.back(); c->evpool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
struct ScopedTimer {
  dqmc_ctx* c; int cat; cudaEvent_t a;
  ScopedTimer(dqmc_ctx* c_, int cat_) : c(c_), cat(cat_), a(nullptr) {
    if (!c->capturing && c->timing >= (cat == TM_SWEEP ? 1 : 2)) { a = ev_get(c); cudaEventRecord(a, c->st); }
  }
  ~ScopedTimer() {
    if (a) { cudaEvent_t b = ev_get(c); cudaEventRecord(b, c->st); c->trecs.push_back({cat, a, b}); }
  }
};
static void timers_resolve(dqmc_ctx* c) {
  for (auto& r : c->trecs) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    cudaEventElapsedTime(&ms, r.a, r.b);
    c->tacc[r.cat] += ms;
    c->evpool.push_back(r.a);
    c->evpool.push_back(r.b);
  }
  c->trecs.clear();
}

// ---------------------------------------------------------------------------------------------- create / destroy
template <typename T>
static int dmalloc(dqmc_ctx* c, T** p, size_t count) {
  CU(c, cudaMalloc((void**)p, sizeof(T) * count));
  CU(c, cudaMemsetAsync(*p, 0, sizeof(T) * count, c->st));
  return 0;
}

extern "C" int dqmc_create(dqmc_ctx** out, const dqmc_params* p) {
  if (!out || !p) CTX_FAIL((dqmc_ctx*)nullptr, "dqmc_create: null argument");
  if (p->opdim != 3 || p->flv != 4) CTX_FAIL((dqmc_ctx*)nullptr, "dqmc_create: only the O(3) model (opdim=3, flv=4) is implemented");
  if (p->slices <= 0 || p->safe_mult <= 0 || p->slices % p->safe_mult != 0)
    CTX_FAIL((dqmc_ctx*)nullptr, "dqmc_create: slices must be a positive multiple of safe_mult (stack.jl:189)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    CTX_FAIL((dqmc_ctx*)nullptr, "dqmc_create: no CUDA device (libdqmc_b200 has no CPU fallback)");
  if (p->device < 0 || p->device >= ndev) CTX_FAIL((dqmc_ctx*)nullptr, "dqmc_create: device %d out of range", p->device);
  dqmc_ctx* c = new dqmc_ctx();
  memset(c->err, 0, sizeof(c->err));
  c->p = *p;
  c->N = p->L * p->L;
  c->n = p->flv * c->N;
  c->M = p->slices;
  c->sm = p->safe_mult;
  c->nel = c->M / c->sm + 1;
  c->kmax = p->delay > 0 ? p->delay : 16;
  if (c->kmax > 16) c->kmax = 16;   // the flush stages at most two k-chunks of 32 columns
  c->have_nbr = c->ops_ready = false;
  c->timing = 0;
  c->capturing = false;
  { const char* e = getenv("DQMC_GRAPHS"); c->use_graphs = e ? atoi(e) != 0 : true; }
  for (int i = 0; i < TM_COUNT; ++i) c->tacc[i] = 0.0;
  c->current_slice = c->M + 1;
  c->direction = -1;
  if (c->n % 8 != 0) { delete c; CTX_FAIL((dqmc_ctx*)nullptr, "dqmc_create: n = 4 L^2 must be a multiple of 8"); }
  CU(c, cudaSetDevice(p->device));
  cudaDeviceProp prop;
  CU(c, cudaGetDeviceProperties(&prop, p->device));
  c->num_sms = prop.multiProcessorCount;
  // The main stream carries the latency-bound critical path (panel chain, narrow updates), the second one the wide trailing updates
  // that run in its shadow: when both have a kernel ready, the critical path's CTAs must get the free SMs first (a cluster of 8 needs
  // them within one GPC, which a 126-CTA wide update would otherwise occupy).  DQMC_STREAM_PRIO=0: equal priorities.
  {
    int prio_lo = 0, prio_hi = 0;
    CU(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // numerically lower = higher priority
    const char* e = getenv("DQMC_STREAM_PRIO");
    const bool prio = e ? atoi(e) != 0 : true;
    CU(c, cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, prio ? prio_hi : 0));
    CU(c, cudaStreamCreateWithPriority(&c->qra.st2, cudaStreamNonBlocking, prio ? prio_lo : 0));
  }
  CU(c, cudaEventCreateWithFlags(&c->qra.eA, cudaEventDisableTiming));
  CU(c, cudaEventCreateWithFlags(&c->qra.eB, cudaEventDisableTiming));
  c->lookahead = getenv("DQMC_NO_LOOKAHEAD") == nullptr;
  { const char* e = getenv("DQMC_LU_BAR"); c->lu_bar_mode = e ? atoi(e) : 1; c->lu_bar_parity = 0; }
  { const char* e = getenv("DQMC_LU_SYM"); c->sym_lu_opt = e ? atoi(e) != 0 : true; }
  { const char* e = getenv("DQMC_GREENS_SYM"); c->sym_greens_opt = e ? atoi(e) != 0 : true; }
  c->sym_model = false; c->sym_G = false;
  { const char* e = getenv("DQMC_PAIRED"); c->paired_opt = e ? atoi(e) != 0 : true; }
  { const char* e = getenv("DQMC_WRAP_SYM"); c->wrap_sym_opt = e ? atoi(e) != 0 : true; }
  const size_t n = c->n, nn = n * n;
  TRY(c, dmalloc(c, &c->G, nn));
  TRY(c, dmalloc(c, &c->Gtmp, nn));
  TRY(c, dmalloc(c, &c->u_stack, nn * c->nel));
  TRY(c, dmalloc(c, &c->t_stack, nn * c->nel));
  TRY(c, dmalloc(c, &c->d_stack, n * c->nel));
  TRY(c, dmalloc(c, &c->Ul, nn)); TRY(c, dmalloc(c, &c->Ur, nn));
  TRY(c, dmalloc(c, &c->Tl, nn)); TRY(c, dmalloc(c, &c->Tr, nn));
  TRY(c, dmalloc(c, &c->Dl, n)); TRY(c, dmalloc(c, &c->Dr, n));
  for (int i = 0; i < 5; ++i) TRY(c, dmalloc(c, &c->W[i], nn));
  TRY(c, dmalloc(c, &c->tau, n));
  const size_t nblk = (n + QR_NB - 1) / QR_NB;
  TRY(c, dmalloc(c, &c->tfac, nblk * QR_NB * QR_NB));
  TRY(c, dmalloc(c, &c->trsm_work, nblk * QR_NB * QR_NB));
  TRY(c, dmalloc(c, &c->dabs, n)); TRY(c, dmalloc(c, &c->drp_inv, n)); TRY(c, dmalloc(c, &c->colnorm, n));
  TRY(c, dmalloc(c, &c->perm, n)); TRY(c, dmalloc(c, &c->pos, n));
  { const char* e = getenv("DQMC_UDT"); c->udt_qrcp_sweep = e && strcmp(e, "qrcp") == 0; }
  TRY(c, dmalloc(c, &c->hs, (size_t)3 * c->N * c->M));
  TRY(c, dmalloc(c, &c->hs_bak, (size_t)3 * c->N * c->M));
  TRY(c, dmalloc(c, &c->nbr, (size_t)4 * c->N));
  // two buffers each (double-buffered by flush batch), pre-filled with the all-ones NaN sentinel the kernel spins on
  const size_t pend = sizeof(cplx) * 2 * 4 * (size_t)(c->kmax + LU_BLOCK_SITES) * n;
  CU(c, cudaMalloc((void**)&c->At, pend));
  CU(c, cudaMalloc((void**)&c->Bm, pend));
  CU(c, cudaMemsetAsync(c->At, 0xFF, pend, c->st));
  CU(c, cudaMemsetAsync(c->Bm, 0xFF, pend, c->st));
  c->unif = nullptr; c->unif_cap = c->unif_n = c->unif_pos_bound = 0;
  TRY(c, dmalloc(c, &c->d_pos, 1)); TRY(c, dmalloc(c, &c->d_acc, 1)); TRY(c, dmalloc(c, &c->d_dS, 1));
  TRY(c, dmalloc(c, &c->d_flags, 4)); TRY(c, dmalloc(c, &c->d_bar, 8));
  TRY(c, dmalloc(c, &c->d_prof, 32)); c->lu_prof = false;
  TRY(c, dmalloc(c, &c->d_logdet, 1)); TRY(c, dmalloc(c, &c->d_check, 2)); TRY(c, dmalloc(c, &c->d_sym, 2));
  c->Gt0 = c->G0t = nullptr; c->td_eye = nullptr; c->td_ones = nullptr;
  for (int i = 0; i < 4; ++i) { c->td_u[i] = c->td_t[i] = nullptr; c->td_d[i] = nullptr; }
  c->gb_G = c->gb_u_stack = c->gb_t_stack = nullptr; c->gb_d_stack = c->gb_hs = nullptr; c->gb_log_det = 0.0;
  TRY(c, dmalloc(c, &c->d_action, 1 + 1024));
  for (int i = 0; i < F_COUNT; ++i) { c->fop[i].nblk = 0; c->fop[i].idx = nullptr; c->fop[i].val = nullptr; }
  c->lu_grid = local_updates_grid(c->n, c->num_sms, &c->lu_rpc);
  {   // the batch depth is halved until the kernel's shared memory fits (large lattices)
    const char* e = getenv("DQMC_LU_KERNEL");
    // block kernel: at most 16 rows per CTA, and L >= 4 (it writes an accepted value into the field while the proposals of the
    // site after next are being evaluated: site i must not be a neighbour of site i+2)
    c->lu_blk = c->lu_rpc <= 16 && c->p.L >= 4 && !(e && strcmp(e, "site") == 0);
    LUArgs probe; probe.nsites = c->N; probe.rpc = c->lu_rpc;
    for (probe.kmax = c->kmax; probe.kmax > 2 && (c->lu_blk ? lu_block_smem(probe) : local_updates_smem(probe)) > (size_t)(c->lu_blk ? 221 : 210) * 1024; probe.kmax /= 2) { }
    c->kmax = probe.kmax;
  }
  CU(c, cudaStreamSynchronize(c->st));
  *out = c;
  return 0;
}

static void drop_graphs(dqmc_ctx* c) {
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second.exec);
  c->graphs.clear();
  c->graph_seen.clear();
}

extern "C" int dqmc_destroy(dqmc_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->p.device);
  cudaStreamSynchronize(c->st);
  timers_resolve(c);
  for (auto e : c->evpool) cudaEventDestroy(e);
  drop_graphs(c);
  void* ptrs[] = {c->G, c->Gtmp, c->u_stack, c->t_stack, c->d_stack, c->Ul, c->Ur, c->Tl, c->Tr, c->Dl, c->Dr,
                  c->W[0], c->W[1], c->W[2], c->W[3], c->W[4], c->tau, c->tfac, c->trsm_work, c->dabs, c->drp_inv,
                  c->colnorm, c->perm, c->pos, c->hs, c->hs_bak, c->nbr, c->At, c->Bm, c->unif, c->d_pos, c->d_acc, c->d_dS,
                  c->d_flags, c->d_bar, c->d_logdet, c->d_check, c->d_sym, c->gb_G, c->gb_u_stack, c->gb_t_stack, c->gb_d_stack,
                  c->gb_hs, c->d_action, c->d_prof, c->Gt0, c->G0t, c->td_eye, c->td_ones, c->td_u[0], c->td_u[1], c->td_u[2],
                  c->td_u[3], c->td_t[0], c->td_t[1], c->td_t[2], c->td_t[3], c->td_d[0], c->td_d[1], c->td_d[2], c->td_d[3]};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < F_COUNT; ++i) { if (c->fop[i].idx) cudaFree(c->fop[i].idx); if (c->fop[i].val) cudaFree(c->fop[i].val); }
  cudaStreamDestroy(c->st);
  cudaStreamDestroy(c->qra.st2);
  cudaEventDestroy(c->qra.eA);
  cudaEventDestroy(c->qra.eB);
  delete c;
  return 0;
}

extern "C" const char* dqmc_last_error(dqmc_ctx* c) { return c ? c->err : g_errbuf; }
extern "C" int64_t dqmc_kernel_launches(dqmc_ctx*) { return g_launches; }

// ---------------------------------------------------------------------------------------------- operators
struct HostQuad { std::vector<int> idx; std::vector<cplx> val; };

static int find_root(std::vector<int>& par, int x) {
  while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; }
  return x;
}

// CSC -> disjoint index groups (connected components of the sparsity pattern) packed into 4-index blocks.
static int csc_to_quad(dqmc_ctx* c, const HostCSC& m, HostQuad* q) {
  const int n = m.n;
  std::vector<int> par(n);
  std::iota(par.begin(), par.end(), 0);
  for (int j = 0; j < n; ++j)
    for (int64_t k = m.colptr[j] - 1; k < m.colptr[j + 1] - 1; ++k) {
      int r = (int)m.rowval[k] - 1;
      int a = find_root(par, r), b = find_root(par, j);
      if (a != b) par[a] = b;
    }
  std::vector<std::vector<int>> comp(n);
  for (int i = 0; i < n; ++i) comp[find_root(par, i)].push_back(i);
  std::vector<std::vector<int>> by_size[5];
  for (int i = 0; i < n; ++i) {
    if (comp[i].empty()) continue;
    if (comp[i].size() > 4)
      CTX_FAIL(c, "operator couples %zu indices in one group; only checkerboard factors made of <=4-site blocks are "
                  "supported (folded CBGeneric groups are out of scope)", comp[i].size());
    by_size[comp[i].size()].push_back(comp[i]);
  }
  std::vector<std::vector<int>> blocks = by_size[4];
  auto merge = [&](std::vector<std::vector<int>>& src, size_t per) -> int {
    if (src.size() % per) return -1;
    for (size_t i = 0; i < src.size(); i += per) {
      std::vector<int> b;
      for (size_t k = 0; k < per; ++k) b.insert(b.end(), src[i + k].begin(), src[i + k].end());
      blocks.push_back(b);
    }
    return 0;
  };
  // 3+1, then 2+2, then 1+1+1+1
  while (!by_size[3].empty() && !by_size[1].empty()) {
    std::vector<int> b = by_size[3].back(); by_size[3].pop_back();
    b.push_back(by_size[1].back()[0]); by_size[1].pop_back();
    blocks.push_back(b);
  }
  if (!by_size[3].empty()) CTX_FAIL(c, "operator block structure cannot be packed into 4-index blocks");
  if (by_size[2].size() % 2 == 1 && by_size[1].size() >= 2) {
    std::vector<int> b = by_size[2].back(); by_size[2].pop_back();
    b.push_back(by_size[1].back()[0]); by_size[1].pop_back();
    b.push_back(by_size[1].back()[0]); by_size[1].pop_back();
    blocks.push_back(b);
  }
  if (merge(by_size[2], 2) || merge(by_size[1], 4)) CTX_FAIL(c, "operator block structure cannot be packed into 4-index blocks");
  const int nblk = (int)blocks.size();
  q->idx.assign((size_t)4 * nblk, 0);
  q->val.assign((size_t)16 * nblk, ZERO);
  std::vector<int> where(n), slot(n);
  for (int b = 0; b < nblk; ++b) {
    std::sort(blocks[b].begin(), blocks[b].end());
    for (int k = 0; k < 4; ++k) { q->idx[4 * b + k] = blocks[b][k]; where[blocks[b][k]] = b; slot[blocks[b][k]] = k; }
  }
  for (int j = 0; j < n; ++j)
    for (int64_t k = m.colptr[j] - 1; k < m.colptr[j + 1] - 1; ++k) {
      int r = (int)m.rowval[k] - 1;
      q->val[(size_t)16 * where[r] + 4 * slot[r] + slot[j]] = m.nz[k];
    }
  return 0;
}

static int upload_quad(dqmc_ctx* c, const HostQuad& q, QuadOp* d) {
  if (d->idx) cudaFree(d->idx);
  if (d->val) cudaFree(d->val);
  d->nblk = (int)(q.idx.size() / 4);
  CU(c, cudaMalloc((void**)&d->idx, sizeof(int) * q.idx.size()));
  CU(c, cudaMalloc((void**)&d->val, sizeof(cplx) * q.val.size()));
  CU(c, cudaMemcpy(d->idx, q.idx.data(), sizeof(int) * q.idx.size(), cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(d->val, q.val.data(), sizeof(cplx) * q.val.size(), cudaMemcpyHostToDevice));
  return 0;
}

static int csc_diag(dqmc_ctx* c, const HostCSC& m, std::vector<cplx>* d) {
  d->assign(m.n, ZERO);
  for (int j = 0; j < m.n; ++j)
    for (int64_t k = m.colptr[j] - 1; k < m.colptr[j + 1] - 1; ++k) {
      if ((int)m.rowval[k] - 1 != j) CTX_FAIL(c, "chkr_mu operator must be diagonal");
      (*d)[j] = m.nz[k];
    }
  return 0;
}

// Does the operator have the antiunitary flavour symmetry X[r+h, c+h] = conj(X[r, c]), X[r+h, c-h] = -conj(X[r, c]) (h = n/2)?
static bool csc_is_antiunitary_symmetric(const HostCSC& m) {
  const int n = m.n, h = n / 2;
  if (n % 2) return false;
  auto at = [&](int r, int cc) -> cplx {
    for (int64_t k = m.colptr[cc] - 1; k < m.colptr[cc + 1] - 1; ++k) if ((int)m.rowval[k] - 1 == r) return m.nz[k];
    return ZERO;
  };
  double vmax = 0.0, viol = 0.0;
  for (int cc = 0; cc < n; ++cc)
    for (int64_t k = m.colptr[cc] - 1; k < m.colptr[cc + 1] - 1; ++k) {
      const int r = (int)m.rowval[k] - 1;
      const cplx v = m.nz[k];
      vmax = std::max(vmax, sqrt(cabs2(v)));
      const int r2 = r < h ? r + h : r - h, c2 = cc < h ? cc + h : cc - h;
      const cplx w = at(r2, c2);
      const bool diag_block = (r < h) == (cc < h);
      const cplx d = diag_block ? cmake(w.x - v.x, w.y + v.y) : cmake(w.x + v.x, w.y - v.y);
      viol = std::max(viol, sqrt(cabs2(d)));
    }
  return viol <= 1e-13 * vmax;
}

static int finalize_operators(dqmc_ctx* c) {
  for (int i = 0; i <= DQMC_OP_MU_INV; ++i) if (!c->csc[i].set) return 0;   // wait until the six factors of B are there
  c->sym_model = true;
  for (int i = 0; i < DQMC_OP_COUNT; ++i)
    if (c->csc[i].set && !csc_is_antiunitary_symmetric(c->csc[i])) c->sym_model = false;
  c->sym_G = false;
  HostQuad hbh, ha, hbhinv, hainv;
  TRY(c, csc_to_quad(c, c->csc[DQMC_OP_HOP_HALF_B], &hbh));
  TRY(c, csc_to_quad(c, c->csc[DQMC_OP_HOP_A], &ha));
  TRY(c, csc_to_quad(c, c->csc[DQMC_OP_HOP_HALF_INV_B], &hbhinv));
  TRY(c, csc_to_quad(c, c->csc[DQMC_OP_HOP_INV_A], &hainv));
  std::vector<cplx> mu, muinv;
  TRY(c, csc_diag(c, c->csc[DQMC_OP_MU], &mu));
  TRY(c, csc_diag(c, c->csc[DQMC_OP_MU_INV], &muinv));
  // fold the diagonal chemical-potential factor into the neighbouring hopping factor:
  //   B = hBh * hA * (hBh * mu) * e^{-dtau V},   B^-1 = e^{+dtau V} * (mu^-1 * hBh^-1) * hA^-1 * hBh^-1
  HostQuad hbh_mu = hbh, muinv_hbhinv = hbhinv;
  for (size_t b = 0; b < hbh.idx.size() / 4; ++b)
    for (int r = 0; r < 4; ++r)
      for (int cc = 0; cc < 4; ++cc) hbh_mu.val[16 * b + 4 * r + cc] = cmul(hbh.val[16 * b + 4 * r + cc], mu[hbh.idx[4 * b + cc]]);
  for (size_t b = 0; b < hbhinv.idx.size() / 4; ++b)
    for (int r = 0; r < 4; ++r)
      for (int cc = 0; cc < 4; ++cc)
        muinv_hbhinv.val[16 * b + 4 * r + cc] = cmul(muinv[hbhinv.idx[4 * b + r]], hbhinv.val[16 * b + 4 * r + cc]);
  TRY(c, upload_quad(c, hbh, &c->fop[F_HBH]));
  TRY(c, upload_quad(c, ha, &c->fop[F_HA]));
  TRY(c, upload_quad(c, hbh_mu, &c->fop[F_HBH_MU]));
  TRY(c, upload_quad(c, hbhinv, &c->fop[F_HBHINV]));
  TRY(c, upload_quad(c, hainv, &c->fop[F_HAINV]));
  TRY(c, upload_quad(c, muinv_hbhinv, &c->fop[F_MUINV_HBHINV]));
  // optional: the half-step factors of group A, used only by effective_greens2greens! (fermion_measurements.jl:1125-1142)
  if (c->csc[DQMC_OP_HOP_HALF_A].set && c->csc[DQMC_OP_HOP_HALF_INV_A].set) {
    HostQuad hah, hahinv;
    TRY(c, csc_to_quad(c, c->csc[DQMC_OP_HOP_HALF_A], &hah));
    TRY(c, csc_to_quad(c, c->csc[DQMC_OP_HOP_HALF_INV_A], &hahinv));
    TRY(c, upload_quad(c, hah, &c->fop[F_HAH]));
    TRY(c, upload_quad(c, hahinv, &c->fop[F_HAHINV]));
  }
  c->ops_ready = true;
  return 0;
}

extern "C" int dqmc_set_operator(dqmc_ctx* c, int which, int64_t m, int64_t n, const int64_t* colptr,
                                 const int64_t* rowval, const void* nzval, int nz_is_complex) {
  if (!c) return -1;
  if (which < 0 || which >= DQMC_OP_COUNT) CTX_FAIL(c, "dqmc_set_operator: unknown operator %d", which);
  if (m != c->n || n != c->n) CTX_FAIL(c, "dqmc_set_operator: operator is %lldx%lld, expected %dx%d", (long long)m, (long long)n, c->n, c->n);
  HostCSC& h = c->csc[which];
  h.n = c->n;
  h.colptr.assign(colptr, colptr + n + 1);
  const int64_t nnz = colptr[n] - 1;
  h.rowval.assign(rowval, rowval + nnz);
  h.nz.resize(nnz);
  for (int64_t k = 0; k < nnz; ++k)
    h.nz[k] = nz_is_complex ? ((const cplx*)nzval)[k] : cmake(((const double*)nzval)[k], 0.0);
  h.set = true;
  c->ops_ready = false;
  CU(c, cudaSetDevice(c->p.device));
  return finalize_operators(c);
}

extern "C" int dqmc_set_neighbors(dqmc_ctx* c, const int64_t* neighbors) {
  if (!c) return -1;
  std::vector<int> nb((size_t)4 * c->N);
  for (size_t i = 0; i < nb.size(); ++i) {
    if (neighbors[i] < 1 || neighbors[i] > c->N) CTX_FAIL(c, "dqmc_set_neighbors: index out of range");
    nb[i] = (int)neighbors[i] - 1;
  }
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaMemcpy(c->nbr, nb.data(), sizeof(int) * nb.size(), cudaMemcpyHostToDevice));
  c->have_nbr = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------- simple get/set
#define NEED_OPS(c) do { if (!(c)->ops_ready) CTX_FAIL(c, "operators not set (dqmc_set_operator for all six factors)"); } while (0)

extern "C" int dqmc_set_hsfield(dqmc_ctx* c, const double* h) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaMemcpyAsync(c->hs, h, sizeof(double) * 3 * c->N * c->M, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}
extern "C" int dqmc_get_hsfield(dqmc_ctx* c, double* h) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaMemcpyAsync(h, c->hs, sizeof(double) * 3 * c->N * c->M, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}
// sym_G <- does the G on the device have the antiunitary flavour symmetry (to 1e-10 of max|G|)?  Synchronises.
static int measure_sym_G(dqmc_ctx* c) {
  double v[2] = {0.0, 0.0};
  if (c->n % 2) { c->sym_G = false; CU(c, cudaStreamSynchronize(c->st)); return 0; }
  TRY(c, sym_violation(c->st, c->G, c->n, c->d_sym, c->num_sms));
  CU(c, cudaMemcpyAsync(v, c->d_sym, sizeof(v), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  c->sym_G = v[0] <= 1e-10 * v[1];
  return 0;
}
extern "C" int dqmc_set_greens(dqmc_ctx* c, const double* g) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaMemcpyAsync(c->G, g, sizeof(cplx) * c->n * c->n, cudaMemcpyHostToDevice, c->st));
  TRY(c, measure_sym_G(c));    // a G without the flavour symmetry switches the local updates to the full flush
  return 0;
}
extern "C" int dqmc_get_greens(dqmc_ctx* c, double* g) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaMemcpyAsync(g, c->G, sizeof(cplx) * c->n * c->n, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}
extern "C" int dqmc_get_state(dqmc_ctx* c, int32_t* slice, int32_t* direction) {
  if (slice) *slice = c->current_slice;
  if (direction) *direction = c->direction;
  return 0;
}
extern "C" int dqmc_set_state(dqmc_ctx* c, int32_t slice, int32_t direction) {
  if (slice < 0 || slice > c->M + 1 || (direction != 1 && direction != -1)) CTX_FAIL(c, "dqmc_set_state: bad state");
  c->current_slice = slice;
  c->direction = direction;
  return 0;
}
extern "C" int dqmc_sync(dqmc_ctx* c) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------- B chains
static void push_op(Chain& ch, const QuadOp& op, int mode) {
  ChainStep& s = ch.s[ch.nsteps++];
  s.kind = 0; s.mode = mode; s.nblk = op.nblk; s.slice = 0; s.sign = 0.0; s.idx = op.idx; s.val = op.val;
}
static void push_ev(Chain& ch, int slice0, double sign, int mode) {
  ChainStep& s = ch.s[ch.nsteps++];
  s.kind = 1; s.mode = mode; s.nblk = 0; s.slice = slice0; s.sign = sign; s.idx = nullptr; s.val = nullptr;
}
// append the steps of one multiply_B_* (slice_matrices.jl:101-226); slice 1-based
static void push_B(dqmc_ctx* c, Chain& ch, int op, int slice) {
  const int s0 = slice - 1;
  switch (op) {
    case DQMC_B_LEFT:         // M <- hBh hA hBh mu eV M
      push_ev(ch, s0, 1.0, OP_N); push_op(ch, c->fop[F_HBH_MU], OP_N); push_op(ch, c->fop[F_HA], OP_N); push_op(ch, c->fop[F_HBH], OP_N);
      break;
    case DQMC_B_INV_LEFT:     // M <- eV^-1 mu^-1 hBh^-1 hA^-1 hBh^-1 M
      push_op(ch, c->fop[F_HBHINV], OP_N); push_op(ch, c->fop[F_HAINV], OP_N); push_op(ch, c->fop[F_MUINV_HBHINV], OP_N); push_ev(ch, s0, -1.0, OP_N);
      break;
    case DQMC_B_DAGGER_LEFT:  // M <- eV mu hBh^† hA^† hBh^† M
      push_op(ch, c->fop[F_HBH], OP_C); push_op(ch, c->fop[F_HA], OP_C); push_op(ch, c->fop[F_HBH_MU], OP_C); push_ev(ch, s0, 1.0, OP_C);
      break;
    case DQMC_B_RIGHT:        // M <- M hBh hA hBh mu eV   (row vector x -> F^T x)
      push_op(ch, c->fop[F_HBH], OP_T); push_op(ch, c->fop[F_HA], OP_T); push_op(ch, c->fop[F_HBH_MU], OP_T); push_ev(ch, s0, 1.0, OP_T);
      break;
    case DQMC_B_INV_RIGHT:    // M <- M eV^-1 mu^-1 hBh^-1 hA^-1 hBh^-1
      push_ev(ch, s0, -1.0, OP_T); push_op(ch, c->fop[F_MUINV_HBHINV], OP_T); push_op(ch, c->fop[F_HAINV], OP_T); push_op(ch, c->fop[F_HBHINV], OP_T);
      break;
  }
}
static bool op_is_rows(int op) { return op == DQMC_B_RIGHT || op == DQMC_B_INV_RIGHT; }

static int run_chain(dqmc_ctx* c, bool rows, cplx* mat, const Chain& ch, const double* colscale, double* colnorm2, int nvtot = 0,
                     int mirror = 0) {
  return launch_apply_chain(rows, mat, c->n, c->n, ch, c->hs, c->N, c->p.lambda * c->p.delta_tau, colscale, colnorm2,
                            c->num_sms, c->st, nvtot, mirror);
}

static int apply_B(dqmc_ctx* c, int op, int slice, cplx* mat, int mirror = 0) {
  Chain ch; ch.nsteps = 0;
  push_B(c, ch, op, slice);
  return run_chain(c, op_is_rows(op), mat, ch, nullptr, nullptr, 0, mirror);
}

// wrap_greens! (stack.jl:316-325)
static int wrap_greens_dev(dqmc_ctx* c, cplx* g, int slice, int dir, bool treat_as_G = false) {
  ScopedTimer t(c, TM_WRAP);
  // B, B^-1 and (while sym_G holds) G have the antiunitary flavour symmetry, hence B G and (B G) B^-1 too: each pass works on
  // half of the vectors and writes the other half as their mirror image (DQMC_WRAP_SYM=0: full passes)
  const int mir = (c->wrap_sym_opt && c->sym_model && c->sym_G && (g == c->G || g == c->Gtmp || treat_as_G) && c->n % 2 == 0) ? 1 : 0;
  if (dir == -1) {
    TRY(c, apply_B(c, DQMC_B_INV_LEFT, slice - 1, g, mir));
    TRY(c, apply_B(c, DQMC_B_RIGHT, slice - 1, g, mir));
  } else {
    TRY(c, apply_B(c, DQMC_B_LEFT, slice, g, mir));
    TRY(c, apply_B(c, DQMC_B_INV_RIGHT, slice, g, mir));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- UDT / greens
// decompose_udt! (linalg.jl:20-39): X (in c->W[0], destroyed; column norms^2 in c->colnorm) -> U, D, T(W[2]).
// Pivoting = one stable sort of the columns by norm, then unpivoted blocked Householder QR.
// (the stack must come from the paired UDT too: its D is exactly paired, D[i] = D[i + n/2], which the half-matrix calculate_greens relies on;
// the column-pivoted UDTs of DQMC_UDT=qrcp order their columns differently)
static bool use_paired(const dqmc_ctx* c) { return c->paired_opt && c->sym_model && !c->udt_qrcp_sweep && c->n % 32 == 0 && c->n >= 64; }

// The same for a matrix with the antiunitary flavour symmetry (every matrix of the stabilization path of a symmetric model):
// only the left-half columns are sorted and factored, by PAIRED Householder steps (qr.cu), n/2 sequential steps instead of n and
// half the flops; U and T are rebuilt in full from their halves.  oracle/experiments/{quaternion_qr,paired_panel_spec}.py.
static int udt_paired_dev(dqmc_ctx* c, cplx* Uout, double* Dout) {
  const int n = c->n, h = n / 2;
  cplx* AL = c->W[1];
  cplx* QHL = c->W[1] + (size_t)n * h;
  TRY(c, argsort_desc(c->st, c->colnorm, h, c->perm));
  TRY(c, gather_interleave_cols(c->st, c->W[0], n, n, c->perm, AL, n, c->num_sms));
  TRY(c, set_identity_lint(c->st, QHL, n, n, c->num_sms));
  TRY(c, qr_factor_paired(c->st, AL, n, n, c->W[3], n, Dout, c->tfac, QHL, n, h, c->lookahead ? &c->qra : nullptr));
  TRY(c, build_U_paired(c->st, QHL, n, n, Uout, n, c->num_sms));
  TRY(c, build_T_paired(c->st, AL, n, n, Dout, c->perm, c->W[2], n, c->num_sms));
  return 0;
}

static int udt_dev(dqmc_ctx* c, cplx* Uout, double* Dout) {
  const int n = c->n;
  if (use_paired(c)) return udt_paired_dev(c, Uout, Dout);
  TRY(c, argsort_desc(c->st, c->colnorm, n, c->perm));
  TRY(c, gather_cols(c->st, c->W[0], n, n, c->perm, c->W[1], n, c->num_sms));
  // Q^H is accumulated on the second stream while the (latency-bound) panel chain runs: Q^H = H_k^H ... H_1^H 1 is the
  // "right-hand side" of the factorization; U = (Q^H)^H afterwards.  (Forming Q by backward accumulation after the
  // factorization costs half the flops but cannot overlap with anything.)
  if (c->lookahead) {
    TRY(c, set_identity(c->st, c->W[0], n, n, c->num_sms));
    TRY(c, qr_factor(c->st, c->W[1], n, n, c->tau, Dout, c->tfac, c->W[0], n, n, c->num_sms, &c->qra));
    TRY(c, ew_combine(c->st, n, EwTerm{c->W[0], 1, nullptr, 0, nullptr, 0}, EwTerm{nullptr, 0, nullptr, 0, nullptr, 0}, 1.0, Uout, c->num_sms));
  } else {
    TRY(c, qr_factor(c->st, c->W[1], n, n, c->tau, Dout, c->tfac, nullptr, 0, 0, c->num_sms, nullptr));
    TRY(c, qr_form_q(c->st, c->W[1], n, n, c->tfac, Uout, n, c->num_sms));
  }
  TRY(c, build_T(c->st, c->W[1], n, n, Dout, c->perm, c->W[2], n, c->num_sms));
  return 0;
}

// decompose_udt! with true column pivoting (zgeqp3 semantics, qrcp.cu): X in W[0] (destroyed) -> U, D, T (W[2]).  Slower than
// udt_dev (level-2, one grid barrier per column) but T is as well conditioned as the reference's; used where T^-1 is applied
// (time-displaced Green's functions), for the public dqmc_decompose_udt, and for the sweep when DQMC_UDT=qrcp.
static int udt_qrcp_dev(dqmc_ctx* c, cplx* Uout, double* Dout) {
  const int n = c->n;
  TRY(c, qrcp_udt(c->st, c->W[0], n, n, c->W[1], n, c->W[2], n, Dout, c->colnorm, c->drp_inv, c->perm, c->pos, c->d_bar + 3 + 1, c->num_sms));
  TRY(c, ew_combine(c->st, n, EwTerm{c->W[1], 1, nullptr, 0, nullptr, 0}, EwTerm{nullptr, 0, nullptr, 0, nullptr, 0}, 1.0, Uout, c->num_sms));
  return 0;
}

static cplx* uslab(dqmc_ctx* c, int idx1) { return c->u_stack + (size_t)(idx1 - 1) * c->n * c->n; }
static cplx* tslab(dqmc_ctx* c, int idx1) { return c->t_stack + (size_t)(idx1 - 1) * c->n * c->n; }
static double* dslab(dqmc_ctx* c, int idx1) { return c->d_stack + (size_t)(idx1 - 1) * c->n; }

// add_slice_sequence_left / _right (stack.jl:278-313); idx 1-based.
static int add_slice_sequence(dqmc_ctx* c, int idx, bool left) {
  ScopedTimer t(c, TM_UDT);
  const int n = c->n;
  const size_t nn = (size_t)n * n;
  const int src = left ? idx : idx + 1, dst = left ? idx + 1 : idx;
  // half-matrix path: B ... B U D is symmetric and the paired UDT reads its left-half columns only
  const bool paired = !c->udt_qrcp_sweep && use_paired(c);
  const int ncols = paired ? n / 2 : n;
  CU(c, cudaMemcpyAsync(c->W[0], uslab(c, src), sizeof(cplx) * (paired ? nn / 2 : nn), cudaMemcpyDeviceToDevice, c->st));
  const int lo = 1 + (idx - 1) * c->sm, hi = idx * c->sm;       // s.ranges[idx]
  Chain ch; ch.nsteps = 0;
  for (int k = 0; k < c->sm; ++k) {
    const int slice = left ? lo + k : hi - k;
    const bool last = (k == c->sm - 1);
    push_B(c, ch, left ? DQMC_B_LEFT : DQMC_B_DAGGER_LEFT, slice);
    if (last || ch.nsteps + 4 > DQMC_MAX_CHAIN) {
      TRY(c, run_chain(c, false, c->W[0], ch, last ? dslab(c, src) : nullptr, last ? c->colnorm : nullptr, ncols));
      ch.nsteps = 0;
    }
  }
  TRY(c, c->udt_qrcp_sweep ? udt_qrcp_dev(c, uslab(c, dst), dslab(c, dst)) : udt_dev(c, uslab(c, dst), dslab(c, dst)));
  if (!c->udt_qrcp_sweep && use_paired(c)) {     // T_dst = T_new T_src is symmetric: left half + mirror
    TRY(c, zgemm(c->st, OP_N, OP_N, n, n / 2, n, ONE, c->W[2], n, tslab(c, src), n, ZERO, tslab(c, dst), n, c->num_sms));
    TRY(c, mirror_right_half(c->st, tslab(c, dst), n, n, c->num_sms));
    return 0;
  }
  TRY(c, zgemm(c->st, OP_N, OP_N, n, n, n, ONE, c->W[2], n, tslab(c, src), n, ZERO, tslab(c, dst), n, c->num_sms));
  return 0;
}

// calculate_greens (stack.jl:338-369): G = [1 + Ul Dl Tl (Ur Dr Tr)^†]^-1, evaluated as
//   G = Ur Drp^-1 [ Dlp^-1 Ul^† Ur Drp^-1 + Dlm Tl Tr^† Drm ]^-1 Dlp^-1 Ul^†      (Dp = max(D,1), Dm = min(D,1))
// with one Householder QR of the bracket (Q^† applied to the right-hand side on the fly) and a triangular solve.
static int calculate_greens_dev(dqmc_ctx* c, bool allow_sym = true) {
  ScopedTimer t(c, TM_GREENS);
  const int n = c->n;
  if (allow_sym && c->sym_greens_opt && use_paired(c)) {
    // half-matrix evaluation (oracle/experiments/half_matrix_greens.py): every operand is symmetric, so the bracket, its
    // right-hand side, the solution and G are formed as left halves (n x n/2) and mirrored; the bracket is factored by the paired QR
    const int h = n / 2;
    cplx *innerL = c->W[2], *rhsL = c->W[2] + (size_t)n * h;
    TRY(c, zgemm(c->st, OP_C, OP_N, n, h, n, ONE, c->Ul, n, c->Ur, n, ZERO, c->W[0], n, c->num_sms));
    TRY(c, zgemm(c->st, OP_N, OP_C, n, h, n, ONE, c->Tl, n, c->Tr, n, ZERO, c->W[1], n, c->num_sms));
    TRY(c, loh_assemble_paired(c->st, n, c->W[0], c->W[1], c->Dl, c->Dr, c->Ul, innerL, rhsL, c->drp_inv, c->num_sms));
    TRY(c, qr_factor_paired(c->st, innerL, n, n, c->W[3], n, c->dabs, c->tfac, rhsL, n, h, c->lookahead ? &c->qra : nullptr));
    TRY(c, expand_R_paired(c->st, innerL, n, n, c->W[4], n, c->num_sms));
    TRY(c, trsm_upper(c->st, c->W[4], n, n, rhsL, n, h, c->trsm_work, nullptr, c->num_sms, 1));
    TRY(c, uninterleave_rows(c->st, rhsL, n, n, c->drp_inv, c->W[0], n, c->num_sms));
    TRY(c, zgemm(c->st, OP_N, OP_N, n, h, n, ONE, c->Ur, n, c->W[0], n, ZERO, c->G, n, c->num_sms));
    TRY(c, mirror_right_half(c->st, c->G, n, n, c->num_sms));
    c->sym_G = true;
    return 0;
  }
  TRY(c, zgemm(c->st, OP_C, OP_N, n, n, n, ONE, c->Ul, n, c->Ur, n, ZERO, c->W[0], n, c->num_sms));
  TRY(c, zgemm(c->st, OP_N, OP_C, n, n, n, ONE, c->Tl, n, c->Tr, n, ZERO, c->W[1], n, c->num_sms));
  TRY(c, loh_assemble(c->st, n, c->W[0], c->W[1], c->Dl, c->Dr, c->Ul, c->W[2], c->W[3], c->drp_inv, c->num_sms));
  TRY(c, qr_factor(c->st, c->W[2], n, n, c->tau, c->dabs, c->tfac, c->W[3], n, n, c->num_sms, c->lookahead ? &c->qra : nullptr));
  TRY(c, trsm_upper(c->st, c->W[2], n, n, c->W[3], n, n, c->trsm_work, c->drp_inv, c->num_sms));
  // G has the antiunitary flavour symmetry [[A, B], [-conj(B), conj(A)]]: the last product is formed for the upper half of the
  // rows only and mirrored (DQMC_GREENS_SYM=0: full product)
  // and only for a model whose operators have the symmetry; caller-supplied UDTs (dqmc_calculate_greens_from) get the full one)
  const bool sym = allow_sym && c->sym_greens_opt && c->sym_model;
  c->sym_G = c->sym_model;
  if (sym && n % 2 == 0) {
    TRY(c, zgemm(c->st, OP_N, OP_N, n / 2, n, n, ONE, c->Ur, n, c->W[3], n, ZERO, c->G, n, c->num_sms));
    TRY(c, mirror_lower_half(c->st, c->G, n, c->num_sms));
  } else {
    TRY(c, zgemm(c->st, OP_N, OP_N, n, n, n, ONE, c->Ur, n, c->W[3], n, ZERO, c->G, n, c->num_sms));
  }
  return 0;
}

static int calculate_logdet_dev(dqmc_ctx* c) {
  return logdet_from_factors(c->st, c->n, c->Dl, c->Dr, c->dabs, c->d_logdet);
}

static int load_udt(dqmc_ctx* c, int idx1, cplx* U, double* D, cplx* T) {
  const size_t nn = (size_t)c->n * c->n;
  CU(c, cudaMemcpyAsync(U, uslab(c, idx1), sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
  CU(c, cudaMemcpyAsync(D, dslab(c, idx1), sizeof(double) * c->n, cudaMemcpyDeviceToDevice, c->st));
  CU(c, cudaMemcpyAsync(T, tslab(c, idx1), sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
  return 0;
}
static int set_udt_identity(dqmc_ctx* c, cplx* U, double* D, cplx* T) {
  TRY(c, set_identity(c->st, U, c->n, c->n, c->num_sms));
  TRY(c, fill_ones(c->st, D, c->n));
  TRY(c, set_identity(c->st, T, c->n, c->n, c->num_sms));
  return 0;
}

static int check_begin(dqmc_ctx* c) {   // copyto!(s.greens_temp, s.greens)
  if (!c->p.all_checks) return 0;
  CU(c, cudaMemcpyAsync(c->Gtmp, c->G, sizeof(cplx) * c->n * c->n, cudaMemcpyDeviceToDevice, c->st));
  return 0;
}
__global__ void check_accumulate_kernel(double* chk) { if (chk[1] > chk[0]) chk[0] = chk[1]; }
static int check_end(dqmc_ctx* c) {     // maximum(absdiff(s.greens_temp, s.greens))
  if (!c->p.all_checks) return 0;
  TRY(c, max_abs_diff(c->st, c->Gtmp, c->G, (size_t)c->n * c->n, c->d_check + 1, c->num_sms));
  check_accumulate_kernel<<<1, 1, 0, c->st>>>(c->d_check);
  g_launches++;
  return 0;
}

extern "C" int dqmc_build_stack(dqmc_ctx* c) {
  NEED_OPS(c);
  CU(c, cudaSetDevice(c->p.device));
  TRY(c, set_udt_identity(c, uslab(c, 1), dslab(c, 1), tslab(c, 1)));
  for (int i = 1; i <= c->nel - 1; ++i) TRY(c, add_slice_sequence(c, i, true));
  c->current_slice = c->M + 1;
  c->direction = -1;
  return 0;
}

// propagate (stack.jl:391-499); everything is enqueued on the context's stream, nothing synchronises.
static int propagate_dev(dqmc_ctx* c);
static int propagate_body(dqmc_ctx* c) {
  const int M = c->M, sm = c->sm;
  if (c->direction == 1) {
    if (c->current_slice % sm == 0) {
      c->current_slice += 1;
      if (c->current_slice == 1) {
        TRY(c, load_udt(c, 1, c->Ur, c->Dr, c->Tr));
        TRY(c, set_udt_identity(c, uslab(c, 1), dslab(c, 1), tslab(c, 1)));
        TRY(c, set_udt_identity(c, c->Ul, c->Dl, c->Tl));
        TRY(c, calculate_greens_dev(c));
        TRY(c, calculate_logdet_dev(c));
      } else if (c->current_slice <= M) {
        const int idx = (c->current_slice - 1) / sm;
        TRY(c, load_udt(c, idx + 1, c->Ur, c->Dr, c->Tr));
        TRY(c, add_slice_sequence(c, idx, true));
        TRY(c, load_udt(c, idx + 1, c->Ul, c->Dl, c->Tl));
        if (c->p.all_checks) {
          TRY(c, check_begin(c));
          TRY(c, wrap_greens_dev(c, c->Gtmp, c->current_slice - 1, 1));
        }
        TRY(c, calculate_greens_dev(c));
        TRY(c, check_end(c));
      } else {
        TRY(c, add_slice_sequence(c, c->nel - 1, true));
        c->direction = -1;
        c->current_slice = M + 1;
        return propagate_body(c);
      }
    } else {
      TRY(c, wrap_greens_dev(c, c->G, c->current_slice, 1));
      c->current_slice += 1;
    }
  } else {
    if ((c->current_slice - 1) % sm == 0) {
      c->current_slice -= 1;
      if (c->current_slice == M) {
        TRY(c, load_udt(c, c->nel, c->Ul, c->Dl, c->Tl));
        TRY(c, set_udt_identity(c, uslab(c, c->nel), dslab(c, c->nel), tslab(c, c->nel)));
        TRY(c, set_udt_identity(c, c->Ur, c->Dr, c->Tr));
        TRY(c, calculate_greens_dev(c));
        TRY(c, calculate_logdet_dev(c));
        TRY(c, wrap_greens_dev(c, c->G, c->current_slice + 1, -1));
      } else if (c->current_slice > 0) {
        const int idx = c->current_slice / sm + 1;
        TRY(c, load_udt(c, idx, c->Ul, c->Dl, c->Tl));
        TRY(c, add_slice_sequence(c, idx, false));
        TRY(c, load_udt(c, idx, c->Ur, c->Dr, c->Tr));
        TRY(c, check_begin(c));
        TRY(c, calculate_greens_dev(c));
        TRY(c, check_end(c));
        TRY(c, wrap_greens_dev(c, c->G, c->current_slice + 1, -1));
      } else {
        TRY(c, add_slice_sequence(c, 1, false));
        c->direction = 1;
        c->current_slice = 0;
        return propagate_body(c);
      }
    } else {
      TRY(c, wrap_greens_dev(c, c->G, c->current_slice, -1));
      c->current_slice -= 1;
    }
  }
  return 0;
}

static int propagate_dev(dqmc_ctx* c) {
  const bool stab = c->direction == 1 ? (c->current_slice % c->sm == 0) : ((c->current_slice - 1) % c->sm == 0);
  if (!stab || !c->use_graphs || c->timing >= 2 || c->capturing) return propagate_body(c);
  const long long key = (long long)c->current_slice * 4 + (c->direction + 1);
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    // first visit: capture (relaxed mode: the one-off kernel attribute opt-ins of a first launch are legal inside it)
    const long long l0 = g_launches;
    const int s_before = c->current_slice, d_before = c->direction;
    CU(c, cudaStreamBeginCapture(c->st, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    const int rc = propagate_body(c);
    c->capturing = false;
    cudaGraph_t g = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(c->st, &g);
    if (rc != 0 || ee != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      c->use_graphs = false;                       // fall back to eager launches for the rest of the run
      c->current_slice = s_before; c->direction = d_before;
      if (rc != 0) return -1;
      return propagate_body(c);
    }
    BlockGraph bg;
    const cudaError_t ei = cudaGraphInstantiate(&bg.exec, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) {
      cudaGetLastError();
      c->use_graphs = false;
      c->current_slice = s_before; c->direction = d_before;
      return propagate_body(c);
    }
    bg.launches = g_launches - l0; bg.slice_after = c->current_slice; bg.dir_after = c->direction; bg.symG_after = c->sym_G;
    g_launches = l0;
    it = c->graphs.emplace(key, bg).first;
  }
  CU(c, cudaGraphLaunch(it->second.exec, c->st));
  g_launches += it->second.launches;
  c->current_slice = it->second.slice_after; c->direction = it->second.dir_after; c->sym_G = it->second.symG_after;
  return 0;
}

extern "C" int dqmc_propagate(dqmc_ctx* c, int32_t* slice, int32_t* direction) {
  NEED_OPS(c);
  CU(c, cudaSetDevice(c->p.device));
  TRY(c, propagate_dev(c));
  if (slice) *slice = c->current_slice;
  if (direction) *direction = c->direction;
  return 0;
}

extern "C" int dqmc_wrap_greens(dqmc_ctx* c, double* g, int32_t slice, int32_t direction) {
  NEED_OPS(c);
  if (direction != 1 && direction != -1) CTX_FAIL(c, "dqmc_wrap_greens: direction must be +1 or -1");
  // the slice matrix used is B(slice) going up and B(slice - 1) going down (stack.jl:316-325); the reference throws a
  // BoundsError outside 1..slices
  const int bs = direction == 1 ? slice : slice - 1;
  if (bs < 1 || bs > c->M) CTX_FAIL(c, "dqmc_wrap_greens: slice %d with direction %d needs B(%d), outside 1..%d", slice, direction, bs, c->M);
  CU(c, cudaSetDevice(c->p.device));
  const size_t bytes = sizeof(cplx) * c->n * c->n;
  if (!g) { TRY(c, wrap_greens_dev(c, c->G, slice, direction)); return 0; }
  CU(c, cudaMemcpyAsync(c->W[4], g, bytes, cudaMemcpyHostToDevice, c->st));
  TRY(c, wrap_greens_dev(c, c->W[4], slice, direction));
  CU(c, cudaMemcpyAsync(g, c->W[4], bytes, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int dqmc_multiply_B(dqmc_ctx* c, int op, int32_t slice, double* m) {
  NEED_OPS(c);
  if (op < 0 || op > DQMC_B_DAGGER_LEFT) CTX_FAIL(c, "dqmc_multiply_B: bad op %d", op);
  if (slice < 1 || slice > c->M) CTX_FAIL(c, "dqmc_multiply_B: slice %d out of range", slice);
  CU(c, cudaSetDevice(c->p.device));
  const size_t bytes = sizeof(cplx) * c->n * c->n;
  CU(c, cudaMemcpyAsync(c->W[4], m, bytes, cudaMemcpyHostToDevice, c->st));
  TRY(c, apply_B(c, op, slice, c->W[4]));
  CU(c, cudaMemcpyAsync(m, c->W[4], bytes, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int dqmc_calculate_greens_from(dqmc_ctx* c, const double* Ul, const double* Dl, const double* Tl,
                                          const double* Ur, const double* Dr, const double* Tr, double* g) {
  CU(c, cudaSetDevice(c->p.device));
  const size_t nn = sizeof(cplx) * c->n * c->n, nd = sizeof(double) * c->n;
  CU(c, cudaMemcpyAsync(c->Ul, Ul, nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Dl, Dl, nd, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Tl, Tl, nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Ur, Ur, nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Dr, Dr, nd, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Tr, Tr, nn, cudaMemcpyHostToDevice, c->st));
  TRY(c, calculate_greens_dev(c, false));   // arbitrary UDTs: full product, then measure what came out
  TRY(c, calculate_logdet_dev(c));
  if (g) CU(c, cudaMemcpyAsync(g, c->G, nn, cudaMemcpyDeviceToHost, c->st));
  TRY(c, measure_sym_G(c));
  return 0;
}

extern "C" int dqmc_logdet(dqmc_ctx* c, double* logdet) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaMemcpyAsync(logdet, c->d_logdet, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int dqmc_decompose_udt(dqmc_ctx* c, const double* x, double* U, double* D, double* T) {
  CU(c, cudaSetDevice(c->p.device));
  const int n = c->n;
  const size_t nn = sizeof(cplx) * n * n;
  CU(c, cudaMemcpyAsync(c->W[0], x, nn, cudaMemcpyHostToDevice, c->st));
  TRY(c, udt_qrcp_dev(c, c->W[3], c->dabs));
  CU(c, cudaMemcpyAsync(U, c->W[3], nn, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(D, c->dabs, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(T, c->W[2], nn, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------- local updates / sweep
static int upload_uniforms(dqmc_ctx* c, const double* u, int64_t nu) {
  if (nu > c->unif_cap) {
    if (c->unif) cudaFree(c->unif);
    CU(c, cudaMalloc((void**)&c->unif, sizeof(double) * (size_t)nu));
    c->unif_cap = nu;
  }
  CU(c, cudaMemcpyAsync(c->unif, u, sizeof(double) * (size_t)nu, cudaMemcpyHostToDevice, c->st));
  c->unif_n = nu;
  c->unif_pos_bound = 0;
  CU(c, cudaMemsetAsync(c->d_pos, 0, sizeof(long long), c->st));
  return 0;
}

extern "C" int dqmc_set_uniforms(dqmc_ctx* c, const double* u, int64_t nu) {
  CU(c, cudaSetDevice(c->p.device));
  TRY(c, upload_uniforms(c, u, nu));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

static int local_updates_dev(dqmc_ctx* c, double box) {
  ScopedTimer t(c, TM_LOCAL);
  if (!c->have_nbr) CTX_FAIL(c, "neighbour table not set (dqmc_set_neighbors)");
  if (c->current_slice < 1 || c->current_slice > c->M) CTX_FAIL(c, "local_updates: current_slice %d is not a physical slice", c->current_slice);
  // the kernel reads a window of up to 4 N draws starting at the (device-side) stream position; positions advance by at most
  // 4 N per slice, so a host-side upper bound is enough to refuse a launch that could run dry (it would have mutated G and
  // the field before the error surfaced)
  if (c->unif_n - c->unif_pos_bound < (long long)4 * c->N)
    CTX_FAIL(c, "uniform stream exhausted: %lld draws left (upper bound), a slice may consume %d", c->unif_n - c->unif_pos_bound, 4 * c->N);
  c->unif_pos_bound += (long long)4 * c->N;
  LUArgs a;
  a.n = c->n; a.nsites = c->N; a.nslices = c->M; a.slice = c->current_slice - 1;
  a.kmax = c->kmax; a.rpc = c->lu_rpc; a.edrun = c->p.edrun;
  a.box = box; a.dtau = c->p.delta_tau; a.lam_dtau = c->p.lambda * c->p.delta_tau;
  a.inv_dtau_c2 = 1.0 / (c->p.delta_tau * c->p.c * c->p.c); a.r = c->p.r; a.u = c->p.u;
  a.G = c->G; a.At = c->At; a.Bm = c->Bm; a.hs = c->hs; a.nbr = c->nbr;
  a.unif = c->unif; a.nunif = c->unif_n; a.pos = c->d_pos; a.accepted = c->d_acc; a.dS = c->d_dS;
  a.flags = c->d_flags; a.bar = c->d_bar; a.prof = c->lu_prof ? c->d_prof : nullptr;
  a.sym = (c->sym_lu_opt && c->sym_model && c->sym_G && c->n % 2 == 0) ? 1 : 0;
  a.bar_mode = c->lu_bar_mode; a.bar_parity = c->lu_bar_parity; c->lu_bar_parity ^= 1;
  TRY(c, c->lu_blk ? launch_lu_block(c->st, a, c->lu_grid) : launch_local_updates(c->st, a, c->lu_grid));
  return 0;
}

static int counters_reset(dqmc_ctx* c) {
  CU(c, cudaMemsetAsync(c->d_acc, 0, sizeof(long long), c->st));
  CU(c, cudaMemsetAsync(c->d_dS, 0, sizeof(double), c->st));
  CU(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->st));   // keep [1] (non-real counter) running
  return 0;
}
static int counters_read(dqmc_ctx* c, int64_t* consumed, int64_t* accepted, double* dS) {
  long long pos = 0, acc = 0; double ds = 0.0; int flags[2] = {0, 0};
  CU(c, cudaMemcpyAsync(&pos, c->d_pos, sizeof(pos), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(&acc, c->d_acc, sizeof(acc), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(&ds, c->d_dS, sizeof(ds), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(flags, c->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  c->unif_pos_bound = pos;     // the true position replaces the upper bound
  if (consumed) *consumed = pos;
  if (accepted) *accepted = acc;
  if (dS) *dS = ds;
  if (flags[0]) CTX_FAIL(c, "uniform stream exhausted: supply at least 4*N draws per slice");
  return 0;
}

extern "C" int dqmc_local_updates(dqmc_ctx* c, double box, const double* u, int64_t nu, int64_t* consumed,
                                  int64_t* accepted, double* dS_total) {
  NEED_OPS(c);
  CU(c, cudaSetDevice(c->p.device));
  if (u) TRY(c, upload_uniforms(c, u, nu));
  TRY(c, counters_reset(c));
  TRY(c, local_updates_dev(c, box));
  return counters_read(c, consumed, accepted, dS_total);
}

extern "C" int dqmc_sweep(dqmc_ctx* c, int32_t nupdates, double box, const double* u, int64_t nu, int64_t* consumed,
                          int64_t* accepted, double* dS_total) {
  NEED_OPS(c);
  CU(c, cudaSetDevice(c->p.device));
  if (u) TRY(c, upload_uniforms(c, u, nu));
  TRY(c, counters_reset(c));
  {
    ScopedTimer t(c, TM_SWEEP);
    for (int k = 0; k < nupdates; ++k) {
      TRY(c, propagate_dev(c));
      TRY(c, local_updates_dev(c, box));
    }
  }
  return counters_read(c, consumed, accepted, dS_total);
}

// ---------------------------------------------------------------------------------------------- boson action / global update
// calc_boson_action (action.jl:1-52): every (site, slice) adds its forward time difference, its "up" and "right" spatial
// differences, and its mass / quartic terms; per-block partial sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256) boson_action_kernel(const double* __restrict__ hs, const int* __restrict__ nbr, int N, int M,
                                                           double dtau, double inv_c2, double r, double u, int edrun,
                                                           double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  const long long tot = (long long)N * M;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % N), s = (int)(e / N);
    const double* h = hs + 3 * e;
    const double sq = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    double t = dtau * r * 0.5 * sq;
    if (!edrun) {
      t += dtau * u * 0.25 * sq * sq;
      const double* hl = hs + 3 * ((long long)i + (long long)N * ((s + 1) % M));
      const double* hu = hs + 3 * ((long long)nbr[4 * i + 0] + (long long)N * s);
      const double* hr = hs + 3 * ((long long)nbr[4 * i + 1] + (long long)N * s);
      double dt2 = 0.0, ds2 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double a = h[k] - hl[k], b = h[k] - hu[k], c = h[k] - hr[k];
        dt2 += a * a;
        ds2 += b * b + c * c;
      }
      t += 0.5 / dtau * inv_c2 * dt2 + 0.5 * dtau * ds2;
    }
    acc += t;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[1 + blockIdx.x] = t;
  }
}
__global__ void boson_action_final_kernel(double* partial, int nblocks) {
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partial[1 + b];
  partial[0] = t;
}
__global__ void shift_field_kernel(double* hs, long long count, double s0, double s1, double s2) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % 3);
    hs[e] = __dadd_rn(hs[e], k == 0 ? s0 : (k == 1 ? s1 : s2));
  }
}

static int boson_action_dev(dqmc_ctx* c, double* out_host) {
  if (!c->have_nbr) CTX_FAIL(c, "neighbour table not set (dqmc_set_neighbors)");
  const int nblocks = 256;
  boson_action_kernel<<<nblocks, 256, 0, c->st>>>(c->hs, c->nbr, c->N, c->M, c->p.delta_tau, 1.0 / (c->p.c * c->p.c), c->p.r,
                                                  c->p.u, c->p.edrun, c->d_action);
  boson_action_final_kernel<<<1, 1, 0, c->st>>>(c->d_action, nblocks);
  CU(c, cudaGetLastError());
  g_launches += 2;
  CU(c, cudaMemcpyAsync(out_host, c->d_action, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int dqmc_calc_boson_action(dqmc_ctx* c, double* S) {
  CU(c, cudaSetDevice(c->p.device));
  return boson_action_dev(c, S);
}

static void global_update_backup_swap(dqmc_ctx* c) {   // global_updates.jl:1-8 (pointer swaps, no copies)
  drop_graphs(c);                                        // the captured graphs hold the old slab pointers
  std::swap(c->gb_u_stack, c->u_stack);
  std::swap(c->gb_d_stack, c->d_stack);
  std::swap(c->gb_t_stack, c->t_stack);
  std::swap(c->gb_G, c->G);
}

// global_update (global_updates.jl:18-59).  u[0..2]: shift draws (randuniform(box_global) per component), u[3]: accept draw,
// consumed only if p_acc <= 1.  S_old = mc.p.boson_action; on acceptance *S_new replaces it.
extern "C" int dqmc_global_update(dqmc_ctx* c, double box_global, const double* u, double S_old, double* S_new,
                                  int32_t* accepted, int32_t* consumed) {
  NEED_OPS(c);
  CU(c, cudaSetDevice(c->p.device));
  if (!(c->current_slice == c->M && c->direction == -1))
    CTX_FAIL(c, "global_update: state must be (slices, -1) (global_updates.jl:22), is (%d, %d)", c->current_slice, c->direction);
  const size_t n = c->n, nn = n * n, nh = (size_t)3 * c->N * c->M;
  if (!c->gb_G) {
    TRY(c, dmalloc(c, &c->gb_G, nn));
    TRY(c, dmalloc(c, &c->gb_u_stack, nn * c->nel));
    TRY(c, dmalloc(c, &c->gb_t_stack, nn * c->nel));
    TRY(c, dmalloc(c, &c->gb_d_stack, n * c->nel));
    TRY(c, dmalloc(c, &c->gb_hs, nh));
  }
  double ld_old = 0.0;
  CU(c, cudaMemcpyAsync(&ld_old, c->d_logdet, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(c->gb_hs, c->hs, sizeof(double) * nh, cudaMemcpyDeviceToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  c->gb_log_det = ld_old;
  global_update_backup_swap(c);
  const double b2 = 2.0 * box_global;
  const double s0 = -box_global + b2 * u[0], s1 = -box_global + b2 * u[1], s2 = -box_global + b2 * u[2];
  shift_field_kernel<<<c->num_sms * 4, 256, 0, c->st>>>(c->hs, (long long)nh, s0, s1, s2);
  CU(c, cudaGetLastError());
  g_launches++;
  TRY(c, dqmc_build_stack(c));
  TRY(c, propagate_dev(c));
  double Snew = 0.0, ld_new = 0.0;
  CU(c, cudaMemcpyAsync(&ld_new, c->d_logdet, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  TRY(c, boson_action_dev(c, &Snew));
  const double p_acc = exp(-(Snew - S_old)) * exp(ld_old - ld_new);
  int acc, used = 3;
  if (p_acc > 1.0) acc = 1;
  else { acc = u[3] < p_acc; used = 4; }
  if (!acc) {   // undo the move: field back, stacks / G / logdet swapped back
    CU(c, cudaMemcpyAsync(c->hs, c->gb_hs, sizeof(double) * nh, cudaMemcpyDeviceToDevice, c->st));
    global_update_backup_swap(c);
    CU(c, cudaMemcpyAsync(c->d_logdet, &ld_old, sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaStreamSynchronize(c->st));
  }
  if (S_new) *S_new = acc ? Snew : S_old;
  if (accepted) *accepted = acc;
  if (consumed) *consumed = used;
  return 0;
}

// ---------------------------------------------------------------------------------------------- boson susceptibility
// measure_chi_dynamic (boson_measurements.jl:6-10, 48-56): chi(qy,qx,iw) = dtau/(N M) sum_k |FT phi_k|^2 on the rfft grid
// qy, qx in 0..L/2, w in 0..M/2.  Separable DFT: time first (the long axis), then the two short lattice axes.
__global__ void __launch_bounds__(128) chi_time_dft_kernel(const double* __restrict__ hs, int N, int M, int nt,
                                                           cplx* __restrict__ F1) {
  // F1[(k + 3*i)*nt + w] = sum_s phi[k,i,s] exp(-2 pi i s w / M); one block per (k,i) row, threads over w
  extern __shared__ double row[];
  const int ki = blockIdx.x;
  for (int s = threadIdx.x; s < M; s += blockDim.x) row[s] = hs[(size_t)ki + (size_t)3 * N * s];
  __syncthreads();
  for (int w = threadIdx.x; w < nt; w += blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int s = 0; s < M; ++s) {
      const long long ph = ((long long)s * w) % M;           // exact phase reduction
      double sn, cs;
      sincospi(-2.0 * (double)ph / (double)M, &sn, &cs);
      re = fma(row[s], cs, re);
      im = fma(row[s], sn, im);
    }
    F1[(size_t)ki * nt + w] = cmake(re, im);
  }
}
__global__ void __launch_bounds__(128) chi_space_dft_kernel(const cplx* __restrict__ F1, int L, int nt, int nq, double scale,
                                                            double* __restrict__ chi) {
  // chi[qy + nq*(qx + nq*w)] = scale * sum_k | sum_{y,x} F1[k,(y,x),w] exp(-2 pi i (y qy + x qx)/L) |^2
  const int tot = nq * nq * nt;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < tot; o += gridDim.x * blockDim.x) {
    const int qy = o % nq, qx = (o / nq) % nq, w = o / (nq * nq);
    double acc = 0.0;
    for (int k = 0; k < 3; ++k) {
      double re = 0.0, im = 0.0;
      for (int x = 0; x < L; ++x)
        for (int y = 0; y < L; ++y) {
          const int ph = (y * qy + x * qx) % L;
          double sn, cs;
          sincospi(-2.0 * (double)ph / (double)L, &sn, &cs);
          const cplx f = F1[(size_t)(k + 3 * (y + L * x)) * nt + w];
          re += f.x * cs - f.y * sn;
          im += f.x * sn + f.y * cs;
        }
      acc += re * re + im * im;
    }
    chi[o] = scale * acc;
  }
}

extern "C" int dqmc_measure_chi_dynamic(dqmc_ctx* c, double* chi) {
  CU(c, cudaSetDevice(c->p.device));
  const int L = c->p.L, N = c->N, M = c->M, nq = L / 2 + 1, nt = M / 2 + 1;
  const size_t need = (size_t)3 * N * nt;                       // complex scratch: fits in one n x n work matrix?
  cplx* F1 = c->W[0];
  cplx* tmpbuf = nullptr;
  if (need + (size_t)nq * nq * nt > (size_t)c->n * c->n) {
    CU(c, cudaMalloc((void**)&tmpbuf, sizeof(cplx) * (need + (size_t)nq * nq * nt)));
    F1 = tmpbuf;
  }
  double* dchi = reinterpret_cast<double*>(F1 + need);
  chi_time_dft_kernel<<<3 * N, 128, sizeof(double) * M, c->st>>>(c->hs, N, M, nt, F1);
  chi_space_dft_kernel<<<(nq * nq * nt + 127) / 128, 128, 0, c->st>>>(F1, L, nt, nq, c->p.delta_tau / ((double)N * M), dchi);
  cudaError_t e = cudaGetLastError();
  g_launches += 2;
  if (e == cudaSuccess) e = cudaMemcpyAsync(chi, dchi, sizeof(double) * nq * nq * nt, cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
  if (tmpbuf) cudaFree(tmpbuf);
  if (e != cudaSuccess) CTX_FAIL(c, "dqmc_measure_chi_dynamic: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int dqmc_timers(dqmc_ctx* c, double* ms, int32_t n) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaStreamSynchronize(c->st));
  timers_resolve(c);
  for (int i = 0; i < n; ++i) ms[i] = i < TM_COUNT ? c->tacc[i] : 0.0;
  for (int i = 0; i < TM_COUNT; ++i) c->tacc[i] = 0.0;
  return 0;
}

extern "C" int dqmc_set_timing(dqmc_ctx* c, int32_t level) {
  c->timing = level < 0 ? 0 : (level > 2 ? 2 : level);
  return 0;
}

// debug hook: per-phase cycle counters of the QR panel kernel (rank 0, thread 0), accumulated while enabled
extern long long* g_qr_prof;
extern int g_qr_prof_rank;
extern "C" int dqmc_qr_profile(dqmc_ctx* c, int32_t enable, int64_t* out8 /* 16 values */) {
  if (enable > 0) g_qr_prof_rank = enable - 1;     // enable = 1 + cluster rank to stamp (paired panel kernel; the unpaired one stamps rank 0)
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaStreamSynchronize(c->st));
  if (out8) CU(c, cudaMemcpy(out8, c->d_prof + 8, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
  CU(c, cudaMemset(c->d_prof + 8, 0, sizeof(long long) * 16));
  g_qr_prof = enable ? c->d_prof + 8 : nullptr;
  return 0;
}

// debug hook: cycle counters of the last local_updates launch (CTA 0); enable with enable != 0
extern "C" int dqmc_lu_profile(dqmc_ctx* c, int32_t enable, int64_t* out16) {
  CU(c, cudaSetDevice(c->p.device));
  c->lu_prof = enable != 0;
  if (out16) {
    CU(c, cudaMemcpyAsync(out16, c->d_prof, sizeof(long long) * 32, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
  }
  return 0;
}

extern "C" int dqmc_checks(dqmc_ctx* c, double* max_propagation_error, int64_t* nonreal) {
  CU(c, cudaSetDevice(c->p.device));
  double chk[2]; int flags[2];
  CU(c, cudaMemcpyAsync(chk, c->d_check, sizeof(chk), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(flags, c->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemsetAsync(c->d_check, 0, sizeof(double) * 2, c->st));
  CU(c, cudaMemsetAsync(c->d_flags + 1, 0, sizeof(int), c->st));
  CU(c, cudaStreamSynchronize(c->st));
  if (max_propagation_error) *max_propagation_error = chk[0];
  if (nonreal) *nonreal = flags[1];
  return 0;
}

// ---------------------------------------------------------------------------------------------- time-displaced G
// [Ua Da Ta + Ub Db Tb]^-1 (inv_sum_udts_scalettar!, linalg.jl:512-567), all operands on the device, into `res`.
// Same scale separation as the reference (Dp = max(D,1), Dm = min(D,1)):
//   res = Tb^-1 Dbp^-1 [ Dam (Ta Tb^-1) Dbp^-1 + Dap^-1 (Ua^H Ub) Dbm ]^-1 Dap^-1 Ua^H
// with every inverse applied through a Householder QR (Q^H folded into the right-hand side) and a triangular solve,
// as in calculate_greens_dev; the reference's two intermediate pivoted UDTs + LU solves are replaced by that.
static int inv_sum_udts_dev(dqmc_ctx* c, const cplx* Ua, const double* Da, const cplx* Ta, const cplx* Ub, const double* Db,
                            const cplx* Tb, cplx* res) {
  const int n = c->n;
  const size_t nn = (size_t)n * n;
  const EwTerm none = {nullptr, 0, nullptr, 0, nullptr, 0};
  const QrAsync* la = c->lookahead ? &c->qra : nullptr;
  // (Ta Tb^-1)^H = (Tb^H)^-1 Ta^H
  TRY(c, ew_combine(c->st, n, EwTerm{Tb, 1, nullptr, 0, nullptr, 0}, none, 1.0, c->W[0], c->num_sms));
  TRY(c, ew_combine(c->st, n, EwTerm{Ta, 1, nullptr, 0, nullptr, 0}, none, 1.0, c->W[1], c->num_sms));
  TRY(c, qr_factor(c->st, c->W[0], n, n, c->tau, c->dabs, c->tfac, c->W[1], n, n, c->num_sms, la));
  TRY(c, trsm_upper(c->st, c->W[0], n, n, c->W[1], n, n, c->trsm_work, nullptr, c->num_sms));
  TRY(c, zgemm(c->st, OP_C, OP_N, n, n, n, ONE, Ua, n, Ub, n, ZERO, c->W[2], n, c->num_sms));
  TRY(c, ew_combine(c->st, n, EwTerm{c->W[1], 1, Da, 3, Db, 2}, EwTerm{c->W[2], 0, Da, 2, Db, 3}, 1.0, c->W[0], c->num_sms));
  // res <- Dbp^-1 [mat]^-1 Dap^-1 Ua^H
  TRY(c, ew_combine(c->st, n, EwTerm{Ua, 1, Da, 2, nullptr, 0}, none, 1.0, res, c->num_sms));
  TRY(c, qr_factor(c->st, c->W[0], n, n, c->tau, c->dabs, c->tfac, res, n, n, c->num_sms, la));
  TRY(c, trsm_upper(c->st, c->W[0], n, n, res, n, n, c->trsm_work, nullptr, c->num_sms));
  TRY(c, ew_combine(c->st, n, EwTerm{res, 0, Db, 2, nullptr, 0}, none, 1.0, res, c->num_sms));
  // res <- Tb^-1 res
  CU(c, cudaMemcpyAsync(c->W[0], Tb, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
  TRY(c, qr_factor(c->st, c->W[0], n, n, c->tau, c->dabs, c->tfac, res, n, n, c->num_sms, la));
  TRY(c, trsm_upper(c->st, c->W[0], n, n, res, n, n, c->trsm_work, nullptr, c->num_sms));
  return 0;
}

// effective_greens2greens! (fermion_measurements.jl:1125-1142): G <- hA½^-1 hB½^-1 G hB½ hA½
static int effective_greens2greens_dev(dqmc_ctx* c, cplx* g) {
  if (!c->fop[F_HAH].idx || !c->fop[F_HAHINV].idx)
    CTX_FAIL(c, "effective_greens2greens: operators DQMC_OP_HOP_HALF_A / DQMC_OP_HOP_HALF_INV_A not set");
  Chain r; r.nsteps = 0;
  push_op(r, c->fop[F_HBH], OP_T); push_op(r, c->fop[F_HAH], OP_T);
  TRY(c, run_chain(c, true, g, r, nullptr, nullptr));
  Chain l; l.nsteps = 0;
  push_op(l, c->fop[F_HBHINV], OP_N); push_op(l, c->fop[F_HAHINV], OP_N);
  TRY(c, run_chain(c, false, g, l, nullptr, nullptr));
  return 0;
}

static int allocate_tdgfs(dqmc_ctx* c) {
  if (c->Gt0) return 0;
  const size_t n = c->n, nn = n * n, K = c->M / c->sm;
  CU(c, cudaMalloc((void**)&c->Gt0, sizeof(cplx) * nn * c->M));
  CU(c, cudaMalloc((void**)&c->G0t, sizeof(cplx) * nn * c->M));
  for (int s = 0; s < 4; ++s) {
    CU(c, cudaMalloc((void**)&c->td_u[s], sizeof(cplx) * nn * K));
    CU(c, cudaMalloc((void**)&c->td_t[s], sizeof(cplx) * nn * K));
    CU(c, cudaMalloc((void**)&c->td_d[s], sizeof(double) * n * K));
  }
  CU(c, cudaMalloc((void**)&c->td_eye, sizeof(cplx) * nn));
  CU(c, cudaMalloc((void**)&c->td_ones, sizeof(double) * n));
  TRY(c, set_identity(c->st, c->td_eye, (int)n, (int)n, c->num_sms));
  TRY(c, fill_ones(c->st, c->td_ones, (int)n));
  return 0;
}

// calc_Bchain_udts! (fermion_measurements.jl:1434-1503) into stack s; entries stored in the reference's final order
// (already reversed for dir = RIGHT).
static int calc_Bchain_udts_dev(dqmc_ctx* c, int s, bool invert, bool left) {
  const int n = c->n, K = c->M / c->sm;
  const size_t nn = (size_t)n * n;
  const bool rightmult = (!left && !invert) || (left && invert);
  const int op = !invert ? (left ? DQMC_B_LEFT : DQMC_B_RIGHT) : (left ? DQMC_B_INV_RIGHT : DQMC_B_INV_LEFT);
  for (int i = 0; i < K; ++i) {
    const int ridx = left ? i : K - 1 - i;                   // index into s.ranges
    const int at = ridx, prev = left ? ridx - 1 : ridx + 1;  // storage positions of this and the previous element
    cplx* U = c->td_u[s] + (size_t)at * nn; cplx* T = c->td_t[s] + (size_t)at * nn; double* D = c->td_d[s] + (size_t)at * n;
    const cplx* src = i == 0 ? c->td_eye : (rightmult ? c->td_t[s] + (size_t)prev * nn : c->td_u[s] + (size_t)prev * nn);
    const double* dprev = i == 0 ? nullptr : c->td_d[s] + (size_t)prev * n;
    CU(c, cudaMemcpyAsync(c->W[0], src, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    const int lo = 1 + ridx * c->sm, hi = (ridx + 1) * c->sm;
    Chain ch; ch.nsteps = 0;
    for (int k = 0; k < c->sm; ++k) {
      const int slice = left ? lo + k : hi - k;
      const bool last = (k == c->sm - 1);
      push_B(c, ch, op, slice);
      if (last || ch.nsteps + 4 > DQMC_MAX_CHAIN) {
        TRY(c, run_chain(c, rightmult, c->W[0], ch, (last && !rightmult) ? dprev : nullptr, (last && !rightmult) ? c->colnorm : nullptr));
        ch.nsteps = 0;
      }
    }
    if (rightmult) {
      if (dprev) TRY(c, ew_combine(c->st, n, EwTerm{c->W[0], 0, dprev, 1, nullptr, 0}, EwTerm{nullptr, 0, nullptr, 0, nullptr, 0}, 1.0, c->W[0], c->num_sms));
      TRY(c, udt_qrcp_dev(c, c->W[3], D));
      CU(c, cudaMemcpyAsync(T, c->W[2], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
      if (i == 0) CU(c, cudaMemcpyAsync(U, c->W[3], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
      else TRY(c, zgemm(c->st, OP_N, OP_N, n, n, n, ONE, c->td_u[s] + (size_t)prev * nn, n, c->W[3], n, ZERO, U, n, c->num_sms));
    } else {
      TRY(c, udt_qrcp_dev(c, U, D));
      if (i == 0) CU(c, cudaMemcpyAsync(T, c->W[2], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
      else TRY(c, zgemm(c->st, OP_N, OP_N, n, n, n, ONE, c->W[2], n, c->td_t[s] + (size_t)prev * nn, n, ZERO, T, n, c->num_sms));
    }
  }
  return 0;
}

// measure_tdgfs! (fermion_measurements.jl:1343-1407) + fill_tdgf! (:1509-1541)
extern "C" int dqmc_measure_tdgfs(dqmc_ctx* c) {
  CU(c, cudaSetDevice(c->p.device));
  NEED_OPS(c);
  if ((c->M / 2) % c->sm != 0 || c->M % 2)
    CTX_FAIL(c, "dqmc_measure_tdgfs: slices/2 = %d must be a multiple of safe_mult = %d (fill_tdgf! starts both directions from "
                "the stabilized slice at beta/2, fermion_measurements.jl:1513-1541)", c->M / 2, c->sm);
  TRY(c, allocate_tdgfs(c));
  const int n = c->n, M = c->M, sm = c->sm, K = M / sm;
  const size_t nn = (size_t)n * n;
  TRY(c, calc_Bchain_udts_dev(c, 0, true, true));     // BT0Inv
  TRY(c, calc_Bchain_udts_dev(c, 1, false, false));   // BBetaT
  TRY(c, calc_Bchain_udts_dev(c, 2, false, true));    // BT0
  TRY(c, calc_Bchain_udts_dev(c, 3, true, false));    // BBetaTInv
  auto U = [&](int s, int i) { return c->td_u[s] + (size_t)i * nn; };
  auto T = [&](int s, int i) { return c->td_t[s] + (size_t)i * nn; };
  auto D = [&](int s, int i) { return c->td_d[s] + (size_t)i * n; };
  for (int i = 0; i < K; ++i) {
    cplx* gt0 = c->Gt0 + (size_t)(i * sm) * nn;
    cplx* g0t = c->G0t + (size_t)(i * sm) * nn;
    if (i > 0) {
      TRY(c, inv_sum_udts_dev(c, U(0, i - 1), D(0, i - 1), T(0, i - 1), U(1, i), D(1, i), T(1, i), gt0));
      TRY(c, inv_sum_udts_dev(c, U(2, i - 1), D(2, i - 1), T(2, i - 1), U(3, i), D(3, i), T(3, i), g0t));
    } else {   // inv_one_plus_udt_scalettar! = the same sum with (1, 1, 1) as first operand
      TRY(c, inv_sum_udts_dev(c, c->td_eye, c->td_ones, c->td_eye, U(1, 0), D(1, 0), T(1, 0), gt0));
      TRY(c, inv_sum_udts_dev(c, c->td_eye, c->td_ones, c->td_eye, U(3, 0), D(3, 0), T(3, 0), g0t));
    }
    TRY(c, effective_greens2greens_dev(c, gt0));
    TRY(c, effective_greens2greens_dev(c, g0t));
  }
  // fill_tdgf!: 0-based tau; reference Mhalf = M/2 + 1
  const int mhalf = M / 2;
  for (int tau = mhalf; tau < M; ++tau) {
    if (tau % sm == 0) continue;
    CU(c, cudaMemcpyAsync(c->Gt0 + (size_t)tau * nn, c->Gt0 + (size_t)(tau - 1) * nn, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    TRY(c, apply_B(c, DQMC_B_LEFT, tau + 1, c->Gt0 + (size_t)tau * nn));
    CU(c, cudaMemcpyAsync(c->G0t + (size_t)tau * nn, c->G0t + (size_t)(tau - 1) * nn, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    TRY(c, apply_B(c, DQMC_B_INV_RIGHT, tau + 1, c->G0t + (size_t)tau * nn));
  }
  for (int tau = mhalf - 1; tau >= 0; --tau) {
    if (tau % sm == 0) continue;
    CU(c, cudaMemcpyAsync(c->Gt0 + (size_t)tau * nn, c->Gt0 + (size_t)(tau + 1) * nn, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    TRY(c, apply_B(c, DQMC_B_INV_LEFT, tau + 2, c->Gt0 + (size_t)tau * nn));
    CU(c, cudaMemcpyAsync(c->G0t + (size_t)tau * nn, c->G0t + (size_t)(tau + 1) * nn, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    TRY(c, apply_B(c, DQMC_B_RIGHT, tau + 2, c->G0t + (size_t)tau * nn));
  }
  for (int tau = 0; tau < M; ++tau)   // minus sign of G(0,tau) (:1402-1404)
    TRY(c, ew_combine(c->st, n, EwTerm{c->G0t + (size_t)tau * nn, 0, nullptr, 0, nullptr, 0}, EwTerm{nullptr, 0, nullptr, 0, nullptr, 0}, -1.0,
                      c->G0t + (size_t)tau * nn, c->num_sms));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int dqmc_get_tdgf(dqmc_ctx* c, int which, int32_t slice, double* out) {
  CU(c, cudaSetDevice(c->p.device));
  if (!c->Gt0) CTX_FAIL(c, "dqmc_get_tdgf: call dqmc_measure_tdgfs first");
  if (slice < 1 || slice > c->M || (which != 0 && which != 1)) CTX_FAIL(c, "dqmc_get_tdgf: bad argument");
  const size_t nn = (size_t)c->n * c->n;
  CU(c, cudaMemcpyAsync(out, (which == 0 ? c->Gt0 : c->G0t) + (size_t)(slice - 1) * nn, sizeof(cplx) * nn, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int dqmc_free_tdgfs(dqmc_ctx* c) {
  CU(c, cudaSetDevice(c->p.device));
  CU(c, cudaStreamSynchronize(c->st));
  void* ptrs[] = {c->Gt0, c->G0t, c->td_eye, c->td_ones, c->td_u[0], c->td_u[1], c->td_u[2], c->td_u[3], c->td_t[0], c->td_t[1],
                  c->td_t[2], c->td_t[3], c->td_d[0], c->td_d[1], c->td_d[2], c->td_d[3]};
  for (void* p : ptrs) if (p) cudaFree(p);
  c->Gt0 = c->G0t = nullptr; c->td_eye = nullptr; c->td_ones = nullptr;
  for (int i = 0; i < 4; ++i) { c->td_u[i] = c->td_t[i] = nullptr; c->td_d[i] = nullptr; }
  return 0;
}

// [Ua Da Ta + Ub Db Tb]^-1 of host operands (test hook for inv_sum_udts_scalettar!, linalg.jl:512-567)
extern "C" int dqmc_inv_sum_udts(dqmc_ctx* c, const double* Ua, const double* Da, const double* Ta, const double* Ub, const double* Db,
                                 const double* Tb, double* res) {
  CU(c, cudaSetDevice(c->p.device));
  const size_t n = c->n, nn = n * n;
  // Ul, Tl, Ur, Tr, Dl, Dr are overwritten by the next stabilization anyway; borrow them as operand buffers
  CU(c, cudaMemcpyAsync(c->Ul, Ua, sizeof(cplx) * nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Tl, Ta, sizeof(cplx) * nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Ur, Ub, sizeof(cplx) * nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Tr, Tb, sizeof(cplx) * nn, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Dl, Da, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->Dr, Db, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
  TRY(c, inv_sum_udts_dev(c, c->Ul, c->Dl, c->Tl, c->Ur, c->Dr, c->Tr, c->W[4]));
  CU(c, cudaMemcpyAsync(res, c->W[4], sizeof(cplx) * nn, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------- bench / test hooks
typedef int (*cublasCreate_t)(void**);
typedef int (*cublasDestroy_t)(void*);
typedef int (*cublasSetStream_t)(void*, cudaStream_t);
typedef int (*cublasZgemm_t)(void*, int, int, int, int, int, const cplx*, const cplx*, int, const cplx*, int, const cplx*, cplx*, int);
typedef int (*cublasDgemm_t)(void*, int, int, int, int, int, const double*, const double*, int, const double*, int, const double*, double*, int);

extern "C" int dqmc_bench_kernel(dqmc_ctx* c, int which, int reps, double* ms_per_launch) {
  CU(c, cudaSetDevice(c->p.device));
  const int n = c->n;
  const size_t nn = (size_t)n * n;
  cudaEvent_t e0, e1;
  CU(c, cudaEventCreate(&e0)); CU(c, cudaEventCreate(&e1));
  const int timing_saved = c->timing;
  c->timing = 0;
  void* blas = nullptr; void* handle = nullptr; cublasZgemm_t zg = nullptr; cublasDgemm_t dg = nullptr; cublasDestroy_t zdestroy = nullptr;
  // FP64 ceilings from cuBLAS (peak probes only, never on the product path): 2 = ZGEMM at the workload's n, 13 = DGEMM and
  // 14 = ZGEMM at 4096^3 on scratch buffers
  const int big = 4096;
  double* scratch = nullptr;
  if (which == 2 || which == 13 || which == 14) {
    blas = dlopen("libcublas.so.12", RTLD_NOW | RTLD_LOCAL);
    if (!blas) blas = dlopen("libcublas.so", RTLD_NOW | RTLD_LOCAL);
    if (!blas) { c->timing = timing_saved; CTX_FAIL(c, "cuBLAS not found (peak probe only): %s", dlerror()); }
    cublasCreate_t zc = (cublasCreate_t)dlsym(blas, "cublasCreate_v2");
    cublasSetStream_t zs = (cublasSetStream_t)dlsym(blas, "cublasSetStream_v2");
    zg = (cublasZgemm_t)dlsym(blas, "cublasZgemm_v2");
    dg = (cublasDgemm_t)dlsym(blas, "cublasDgemm_v2");
    zdestroy = (cublasDestroy_t)dlsym(blas, "cublasDestroy_v2");
    if (!zc || !zs || !zg || !dg || zc(&handle) != 0 || zs(handle, c->st) != 0) { c->timing = timing_saved; CTX_FAIL(c, "cuBLAS init failed"); }
    if (which != 2) {
      const size_t cnt = (size_t)big * big * (which == 14 ? 2 : 1);
      CU(c, cudaMalloc((void**)&scratch, sizeof(double) * 3 * cnt));
      CU(c, cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * cnt, c->st));
    }
  }
  const double done = 1.0, dzero = 0.0;
  if (which == 0 || which == 5 || which == 7) NEED_OPS(c);
  long long pos_saved = 0;
  const long long bound_saved = c->unif_pos_bound;
  const int slice_saved = c->current_slice;
  if (which == 5) {
    CU(c, cudaMemcpyAsync(&pos_saved, c->d_pos, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaMemcpyAsync(c->W[4], c->G, sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->hs_bak, c->hs, sizeof(double) * 3 * c->N * c->M, cudaMemcpyDeviceToDevice, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    c->current_slice = 1;
  }
  int rc = 0;
  for (int it = -1; it < reps && rc == 0; ++it) {     // it == -1: warm-up
    if (it == 0) CU(c, cudaEventRecord(e0, c->st));
    switch (which) {
      case 0: rc = wrap_greens_dev(c, c->W[4], 1, 1, true); break;   // timed like the sweep's wrap of G (half passes + mirror while G is symmetric)
      case 1: rc = zgemm(c->st, OP_N, OP_N, n, n, n, ONE, c->W[0], n, c->W[1], n, ZERO, c->W[2], n, c->num_sms); break;
      case 2: rc = zg(handle, 0, 0, n, n, n, &ONE, c->W[0], n, c->W[1], n, &ZERO, c->W[2], n); break;
      case 13: rc = dg(handle, 0, 0, big, big, big, &done, scratch, big, scratch + (size_t)big * big, big, &dzero, scratch + (size_t)2 * big * big, big); break;
      case 14: rc = zg(handle, 0, 0, big, big, big, &ONE, (cplx*)scratch, big, (cplx*)scratch + (size_t)big * big, big, &ZERO, (cplx*)scratch + (size_t)2 * big * big, big); break;
      case 3:
        rc = cudaMemcpyAsync(c->W[0], c->W[4], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st) != cudaSuccess;
        if (!rc) rc = colnorm2(c->st, c->W[0], n, n, c->colnorm);
        if (!rc) rc = udt_dev(c, c->W[3], c->dabs);
        break;
      case 4: rc = calculate_greens_dev(c); break;
      case 5: {
        CU(c, cudaMemsetAsync(c->d_pos, 0, sizeof(long long), c->st));
        c->unif_pos_bound = 0;
        rc = local_updates_dev(c, 0.5);
        break;
      }
      case 6: rc = cudaMemcpyAsync(c->W[0], c->W[1], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st) != cudaSuccess; break;
      case 7: {
        Chain ch; ch.nsteps = 0;
        for (int k = 0; k < c->sm && ch.nsteps + 4 <= DQMC_MAX_CHAIN; ++k) push_B(c, ch, DQMC_B_LEFT, 1 + k);
        rc = run_chain(c, false, c->W[4], ch, nullptr, c->colnorm);
        break;
      }
      case 8:    // QR only (W[1] <- W[4], factor)
        rc = cudaMemcpyAsync(c->W[1], c->W[4], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st) != cudaSuccess;
        if (!rc) rc = qr_factor(c->st, c->W[1], n, n, c->tau, c->dabs, c->tfac, nullptr, 0, 0, c->num_sms, c->lookahead ? &c->qra : nullptr);
        break;
      case 9: rc = qr_form_q(c->st, c->W[1], n, n, c->tfac, c->W[3], n, c->num_sms); break;
      case 10:   // the serial panel chain alone (no trailing updates): lower bound of the QR critical path
        rc = qr_panels_only(c->st, c->W[1], n, n, c->tau, c->dabs, c->tfac);
        break;
      case 11: rc = trsm_upper(c->st, c->W[1], n, n, c->W[3], n, n, c->trsm_work, nullptr, c->num_sms); break;
      case 12:   // QR with Q^H applied to n right-hand sides (the calculate_greens shape)
        rc = cudaMemcpyAsync(c->W[1], c->W[4], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st) != cudaSuccess;
        if (!rc) rc = qr_factor(c->st, c->W[1], n, n, c->tau, c->dabs, c->tfac, c->W[3], n, n, c->num_sms, c->lookahead ? &c->qra : nullptr);
        break;
      case 15:   // paired QR of a left half with Q^H applied to n/2 right-hand sides (the shape of both paired call sites)
      case 17:   // ... without right-hand sides
        rc = cudaMemcpyAsync(c->W[1], c->W[4], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st) != cudaSuccess;
        if (!rc) rc = qr_factor_paired(c->st, c->W[1], n, n, c->W[3], n, c->dabs, c->tfac, which == 15 ? c->W[1] + nn / 2 : nullptr, n, n / 2,
                                       c->lookahead ? &c->qra : nullptr);
        break;
      case 16:   // the serial paired panel chain alone
        rc = qr_panels_only_paired(c->st, c->W[1], n, n, c->W[3], n, c->dabs, c->tfac);
        break;
      default: rc = -1; snprintf(g_errbuf, sizeof(g_errbuf), "dqmc_bench_kernel: unknown kernel %d", which);
    }
  }
  if (rc == 0) {
    CU(c, cudaEventRecord(e1, c->st));
    CU(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(c, cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / reps;
  }
  if (which == 5) {
    CU(c, cudaMemcpyAsync(c->G, c->W[4], sizeof(cplx) * nn, cudaMemcpyDeviceToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->hs, c->hs_bak, sizeof(double) * 3 * c->N * c->M, cudaMemcpyDeviceToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->d_pos, &pos_saved, sizeof(long long), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    c->current_slice = slice_saved;
    c->unif_pos_bound = bound_saved;
  }
  if (handle && zdestroy) zdestroy(handle);
  if (scratch) cudaFree(scratch);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  c->timing = timing_saved;
  if (rc) { memcpy(c->err, g_errbuf, sizeof(g_errbuf)); return -1; }
  return 0;
}

extern "C" int dqmc_test_zgemm(dqmc_ctx* c, int opA, int opB, int M, int N, int K, const double* alpha, const double* A,
                               int lda, const double* B, int ldb, const double* beta, double* C, int ldc) {
  CU(c, cudaSetDevice(c->p.device));
  const int acols = opA == OP_N ? K : M, bcols = opB == OP_N ? N : K;
  cplx *dA, *dB, *dC;
  CU(c, cudaMalloc((void**)&dA, sizeof(cplx) * (size_t)lda * acols));
  CU(c, cudaMalloc((void**)&dB, sizeof(cplx) * (size_t)ldb * bcols));
  CU(c, cudaMalloc((void**)&dC, sizeof(cplx) * (size_t)ldc * N));
  CU(c, cudaMemcpy(dA, A, sizeof(cplx) * (size_t)lda * acols, cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(dB, B, sizeof(cplx) * (size_t)ldb * bcols, cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(dC, C, sizeof(cplx) * (size_t)ldc * N, cudaMemcpyHostToDevice));
  int rc = zgemm(c->st, opA, opB, M, N, K, cmake(alpha[0], alpha[1]), dA, lda, dB, ldb, cmake(beta[0], beta[1]), dC, ldc, c->num_sms);
  cudaError_t e = cudaStreamSynchronize(c->st);
  if (rc == 0 && e == cudaSuccess) e = cudaMemcpy(C, dC, sizeof(cplx) * (size_t)ldc * N, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  if (rc) { memcpy(c->err, g_errbuf, sizeof(g_errbuf)); return -1; }
  if (e != cudaSuccess) CTX_FAIL(c, "dqmc_test_zgemm: %s", cudaGetErrorString(e));
  return 0;
}

// Test hooks of the half-matrix path.  dqmc_test_qr_paired: the paired QR on host data (n = the context's matrix size): XL and rhs
// are n x n/2 with pair-interleaved rows; on return XL = R_L, rhs = Q^H rhs, V = the n x n explicit reflector blocks, Tfac = the
// n/32 compact-WY factors (32 x 32 each), dabs = n moduli.  lookahead != 0 runs the two-stream driver.
extern "C" int dqmc_test_qr_paired(dqmc_ctx* c, double* XL, double* rhs, double* V, double* Tfac, double* dabs, int32_t lookahead) {
  CU(c, cudaSetDevice(c->p.device));
  const int n = c->n, h = n / 2;
  const size_t half = sizeof(cplx) * (size_t)n * h;
  CU(c, cudaMemcpyAsync(c->W[1], XL, half, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->W[1] + (size_t)n * h, rhs, half, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemsetAsync(c->W[3], 0, 2 * half, c->st));
  TRY(c, qr_factor_paired(c->st, c->W[1], n, n, c->W[3], n, c->dabs, c->tfac, c->W[1] + (size_t)n * h, n, h, lookahead ? &c->qra : nullptr));
  CU(c, cudaMemcpyAsync(XL, c->W[1], half, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(rhs, c->W[1] + (size_t)n * h, half, cudaMemcpyDeviceToHost, c->st));
  if (V) CU(c, cudaMemcpyAsync(V, c->W[3], 2 * half, cudaMemcpyDeviceToHost, c->st));
  if (Tfac) CU(c, cudaMemcpyAsync(Tfac, c->tfac, sizeof(cplx) * (size_t)(n / 32) * QR_NB * QR_NB, cudaMemcpyDeviceToHost, c->st));
  if (dabs) CU(c, cudaMemcpyAsync(dabs, c->dabs, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

// dqmc_test_udt: the sweep's decompose_udt! on a host matrix (sort-once QR, or the paired one if paired != 0; the caller
// guarantees the symmetry then) -> U, D, T.
extern "C" int dqmc_test_udt(dqmc_ctx* c, const double* x, double* U, double* D, double* T, int32_t paired) {
  CU(c, cudaSetDevice(c->p.device));
  const int n = c->n;
  const size_t nn = sizeof(cplx) * (size_t)n * n;
  if (paired && (n % 32 != 0 || n < 64)) CTX_FAIL(c, "dqmc_test_udt: the paired factorization needs n %% 32 == 0, n >= 64");
  CU(c, cudaMemcpyAsync(c->W[0], x, nn, cudaMemcpyHostToDevice, c->st));
  TRY(c, colnorm2(c->st, c->W[0], n, n, c->colnorm));
  const bool po = c->paired_opt, smo = c->sym_model;
  c->paired_opt = paired != 0; if (paired) c->sym_model = true;
  const int rc = udt_dev(c, c->W[4], c->dabs);
  c->paired_opt = po; c->sym_model = smo;
  if (rc) { memcpy(c->err, g_errbuf, sizeof(g_errbuf)); return -1; }
  CU(c, cudaMemcpyAsync(U, c->W[4], nn, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(T, c->W[2], nn, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(D, c->dabs, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return 0;
}

// engine.cu — per-block engine, communication back ends and the C-ABI of libmsplit.so.
// See include/msplit.h for the reference function each entry point replaces.
#include "../../include/msplit.h"
#include "kernels.cuh"

#include <cub/device/device_scan.cuh>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <map>
#include <tuple>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------
// error handling: no exceptions across the ABI; PetscCall-like early return with a message
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
#define MSP_FAIL(msg)                                                                  \
  do {                                                                                 \
    g_err = std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg);     \
    return 1;                                                                          \
  } while (0)
#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) MSP_FAIL(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define RC(call)                                                                       \
  do {                                                                                 \
    int rc_ = (call);                                                                  \
    if (rc_) return rc_;                                                               \
  } while (0)

static int g_num_sms = 148;
static inline int grid_for(long long work_items, int per_sm = 8) {
  long long need = (work_items + MSPK_THREADS - 1) / MSPK_THREADS;
  long long cap = (long long)g_num_sms * per_sm;
  if (cap > MSPK_MAX_PART - 1) cap = (MSPK_MAX_PART - 1) / g_num_sms * g_num_sms;
  if (need < 1) need = 1;
  return (int)std::min(need, cap);
}

// ------------------------------------------------------------------------------------------------
// communication back ends
// ------------------------------------------------------------------------------------------------
struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() {}
  // in-place sum over ranks of n doubles in device memory, result on every rank, stream ordered
  virtual int allreduce_sum(double *dbuf, int n, cudaStream_t st) = 0;
  virtual int barrier(cudaStream_t st) = 0;
};

struct SelfComm : Comm {
  int allreduce_sum(double *, int, cudaStream_t) override { return 0; }
  int barrier(cudaStream_t) override { return 0; }
};

// all blocks in one process, one host thread per block: host-side deterministic reduction
struct LocalShared {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long gen = 0;
  std::vector<std::vector<double>> slot;
  std::vector<double> result;
  explicit LocalShared(int n_) : n(n_), slot(n_), result(0) {}
  void wait_all() {
    std::unique_lock<std::mutex> lk(mu);
    long long g = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct LocalComm : Comm {
  LocalShared *sh;
  std::vector<double> host;
  LocalComm(LocalShared *s, int r) : sh(s) { rank = r; nranks = s->n; }
  int allreduce_sum(double *dbuf, int n, cudaStream_t st) override {
    host.resize(n);
    CK(cudaMemcpyAsync(host.data(), dbuf, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    sh->slot[rank] = host;
    sh->wait_all();
    for (int i = 0; i < n; i++) {
      double s = 0.0;
      for (int r = 0; r < nranks; r++) s += sh->slot[r][i]; // rank order: identical on every rank
      host[i] = s;
    }
    sh->wait_all();
    CK(cudaMemcpyAsync(dbuf, host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return 0;
  }
  int barrier(cudaStream_t st) override {
    CK(cudaStreamSynchronize(st));
    sh->wait_all();
    return 0;
  }
};

// one process per GPU: NCCL (resolved at run time so that a process that already loaded torch's
// bundled libnccl.so.2 shares it)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load() {
    if (h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
    GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce;
  }
};
static NcclApi g_nccl;
struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  double *scratch = nullptr;
  ~NcclComm() override { if (comm) g_nccl.CommDestroy(comm); if (scratch) cudaFree(scratch); }
  int allreduce_sum(double *dbuf, int n, cudaStream_t st) override {
    int rc = g_nccl.AllReduce(dbuf, dbuf, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, st);
    if (rc) MSP_FAIL(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    return 0;
  }
  int barrier(cudaStream_t st) override {
    if (!scratch) { CK(cudaMalloc(&scratch, 64)); CK(cudaMemsetAsync(scratch, 0, 64, st)); }
    return allreduce_sum(scratch, 1, st);
  }
};

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
// receive window, one allocation (exportable with one cudaIpcMemHandle), in doubles:
//   [ lo par0 | lo par1 | hi par0 | hi par1 ]  4 x H   boundary layers (sync: parity = exchange count; async: seq & 1)
//   [ 8 ]                                      two AsyncHdrDev (lo, hi)
//   [ 16 ]                                     CdMailbox (convergence-detection messages)
//   [ G x fslot ]                              TSQR factor mailboxes of the asynchronous global minimisation:
//                                              slot J = { seq, (s+1)^2 factor of block J, seq }
struct Window {
  double *base = nullptr;
  size_t bytes = 0;
  int H = 0, G = 0, fslot = 0;
  double *halo(int side, int par) const { return base + (size_t)(side * 2 + par) * H; }
  AsyncHdrDev *hdr(int side) const { return reinterpret_cast<AsyncHdrDev *>(base + (size_t)4 * H) + side; }
  CdMailbox *mailbox() const { return reinterpret_cast<CdMailbox *>(base + (size_t)4 * H + 8); }
  double *factor(int J) const { return base + (size_t)4 * H + 8 + 16 + (size_t)J * fslot; }
  static int fslot_for(int smax) { return (smax + 1) * (smax + 1) + 2; }
  static size_t size_for(int H, int G, int smax) { return sizeof(double) * ((size_t)4 * H + 8 + 16 + (size_t)G * fslot_for(smax) + 8); }
};
static_assert(sizeof(CdMailbox) == 128, "mailbox layout");

struct msp_engine {
  int device = 0;
  cudaStream_t st = nullptr;
  msp_problem prob{};
  int nb = 0, H = 0, W = 0, off = 0;
  long long ld = 0, ntot = 0;
  int64_t nnz = 0;
  int *rp = nullptr, *ci = nullptr; double *va = nullptr; // strip CSR
  int *ecol = nullptr; double *eval = nullptr;            // ELL
  int *brow = nullptr; int nbrow = 0;
  double *b = nullptr, *rhs = nullptr, *x = nullptr;
  double *halo[2] = {nullptr, nullptr}; // private copies of the neighbours' boundary layers
  bool has_nb[2] = {false, false};
  double *V = nullptr; int nvec = 0;
  double *Wb[2] = {nullptr, nullptr};
  double *S = nullptr, *Slo = nullptr, *Shi = nullptr, *R = nullptr; int smax = 0;
  GmresCtl *ctl = nullptr;
  ReduceWs ws{};
  double *dsc = nullptr; // device scalars [256]
  double *dfac = nullptr; // device staging of the stacked TSQR factors [G x (smax+1)^2]
  double *gram_partial = nullptr; // [45 x MSPK_MAX_PART] partials of the Gram kernel
  bool use_cholqr = true;
  double *hsc = nullptr; // pinned host scalars [256]
  Window win;            // own receive window
  Window peer[2];        // neighbours' windows (peer / IPC mapped); base null if no neighbour
  bool peer_ipc[2] = {false, false};
  Window peer_any[MSP_MAX_BLOCKS]; // every block's window (factor mailboxes of the async global minimisation)
  bool peer_any_ipc[MSP_MAX_BLOCKS] = {};
  int *aint = nullptr;   // device ints: [0..1] last seen async header seq per side
  ProbeDecision *dec = nullptr; // [2]
  int async_sent[2] = {0, 0};
  int factor_sent = 0;
  std::vector<double> fcache; // newest valid factor seen from each block [G x fslot]
  std::vector<int> fcache_seq;
  int par = 0;
  Comm *comm = nullptr;
  bool own_comm = false;
  CdState *cd = nullptr;
  int64_t launches = 0;
  // per-class profiling (CUDA events around each hot-kernel launch, on the launching stream)
  bool prof = false;
  struct ProfRec { int cls; cudaEvent_t a, b; double bytes; };
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void prof_begin(int cls, double bytes) {
    if (!prof) return;
    ProfRec r{cls, get_event(), get_event(), bytes};
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  }
  void prof_end() { if (prof) cudaEventRecord(recs.back().b, st); }
  void prof_collect(msp_result *res) {
    if (!prof) return;
    cudaStreamSynchronize(st);
    double t[4] = {0, 0, 0, 0}, by[4] = {0, 0, 0, 0}; int64_t n[4] = {0, 0, 0, 0};
    for (auto &r : recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      t[r.cls] += ms; by[r.cls] += r.bytes; n[r.cls]++;
      ev_pool.push_back(r.a); ev_pool.push_back(r.b);
    }
    recs.clear();
    res->t_spmv_ms = t[0]; res->t_mdot_ms = t[1]; res->t_maxpy_ms = t[2]; res->t_other_ms = t[3];
    res->b_spmv = by[0]; res->b_mdot = by[1]; res->b_maxpy = by[2]; res->b_other = by[3];
    res->n_spmv = n[0]; res->n_mdot = n[1]; res->n_maxpy = n[2]; res->n_other = n[3];
  }
  // CUDA graphs of whole restart cycles (launch-bound regime: small blocks), keyed by everything baked into the nodes
  typedef std::tuple<int, int, int, const void *, const void *, const void *> CycleKey;
  struct CycleGraph { cudaGraphExec_t exec; int launches; };
  std::map<CycleKey, CycleGraph> cycle_graphs;
  bool use_graphs = true;
  // deterministic turn taking for the emulated asynchronous schedule
  struct msp_group *grp = nullptr;
};

static int64_t stencil_nnz_host(int dim, int nx, int ny, int nz, long long row0, long long nb) {
  // closed form would do; a loop keeps it obviously equal to the kernels' rule
  int64_t nnz = 0;
  if (dim == 2) {
    for (long long r = row0; r < row0 + nb; r++) {
      long long i = r / nx, j = r - i * nx;
      nnz += 1 + (i > 0) + (j > 0) + (j < nx - 1) + (i < ny - 1);
    }
  } else {
    long long pl = (long long)nx * ny;
    for (long long r = row0; r < row0 + nb; r++) {
      long long k = r / pl, rem = r - k * pl, j = rem / nx, i = rem - j * nx;
      nnz += 1 + (k > 0) + (j > 0) + (i > 0) + (i < nx - 1) + (j < ny - 1) + (k < nz - 1);
    }
  }
  return nnz;
}

static int exclusive_scan_inplace(int *cnt_to_rowptr /* nb+1, last = 0 */, int n_plus_1, cudaStream_t st) {
  void *tmp = nullptr; size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt_to_rowptr, cnt_to_rowptr, n_plus_1, st));
  CK(cudaMalloc(&tmp, tb));
  CK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt_to_rowptr, cnt_to_rowptr, n_plus_1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(tmp));
  return 0;
}

// device assembly of a strip into freshly allocated CSR arrays
static int assemble_strip_dev(int dim, int nx, int ny, int nz, long long row0, int nb, cudaStream_t st, int **rp, int **ci,
                              double **va, int64_t *nnz_out) {
  CK(cudaMalloc(rp, sizeof(int) * ((size_t)nb + 1)));
  CK(cudaMemsetAsync(*rp, 0, sizeof(int) * ((size_t)nb + 1), st));
  k_stencil_count<<<grid_for(nb, 16), MSPK_THREADS, 0, st>>>(dim, nx, ny, nz, row0, nb, *rp);
  RC(exclusive_scan_inplace(*rp, nb + 1, st));
  int nnz32 = 0;
  CK(cudaMemcpy(&nnz32, *rp + nb, sizeof(int), cudaMemcpyDeviceToHost));
  *nnz_out = nnz32;
  CK(cudaMalloc(ci, sizeof(int) * (size_t)std::max(nnz32, 1)));
  CK(cudaMalloc(va, sizeof(double) * (size_t)std::max(nnz32, 1)));
  k_stencil_fill<<<grid_for(nb, 16), MSPK_THREADS, 0, st>>>(dim, nx, ny, nz, row0, nb, *rp, *ci, *va);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return 0;
}

static int set_device(int device) {
  CK(cudaSetDevice(device));
  static std::once_flag once;
  std::call_once(once, [&] {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) == cudaSuccess && p.multiProcessorCount > 0) g_num_sms = p.multiProcessorCount;
  });
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel launch helpers (every launch is counted: bench.py reports gpu_launches)
// ------------------------------------------------------------------------------------------------
// resident blocks per SM of a kernel (cached): grids of the grid-stride kernels are SMs x this, i.e. exactly one full
// wave, whatever register count ptxas chose (the 7-point SpMV needs 40 registers: 6 blocks per SM, not 8)
template <typename K>
static int resident_blocks_per_sm(K kernel) {
  static std::mutex mu;
  static std::map<const void *, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find((const void *)kernel);
  if (it != cache.end()) return it->second;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, MSPK_THREADS, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 4; }
  cache[(const void *)kernel] = nb;
  return nb;
}
template <int MODE, bool RESID, bool SCALE, bool NORM>
static void launch_spmv_w(msp_engine *e, const SpmvArgs &a, int ws_slot, GmresCtl *ctl_rw) {
  const long long items = ((long long)a.nb + 1) / 2;
  e->prof_begin(0, 12.0 * (double)e->nnz + 4.0 * (e->nb + 1) + 16.0 * e->nb + (RESID ? 8.0 * e->nb : 0.0));
  if (a.W == 5) {
    auto k = k_spmv_ell<5, MODE, RESID, SCALE, NORM>;
    k<<<grid_for(items, resident_blocks_per_sm(k)), MSPK_THREADS, 0, e->st>>>(a, e->ws, ws_slot, ctl_rw);
  } else if (a.W == 7) {
    auto k = k_spmv_ell<7, MODE, RESID, SCALE, NORM>;
    k<<<grid_for(items, resident_blocks_per_sm(k)), MSPK_THREADS, 0, e->st>>>(a, e->ws, ws_slot, ctl_rw);
  } else {
    auto k = k_spmv_ell<0, MODE, RESID, SCALE, NORM>;
    k<<<grid_for(items, resident_blocks_per_sm(k)), MSPK_THREADS, 0, e->st>>>(a, e->ws, ws_slot, ctl_rw);
  }
  e->prof_end();
  e->launches++;
}

static SpmvArgs spmv_args(msp_engine *e, const double *x, double *y) {
  SpmvArgs a{};
  a.nb = e->nb; a.W = e->W; a.H = e->H; a.ld = e->ld; a.ecol = e->ecol; a.eval = e->eval;
  a.x = x; a.y = y; a.lo = nullptr; a.hi = nullptr; a.b = nullptr; a.ctl = e->ctl; a.guard_it = -1;
  return a;
}

static void launch_mdot(msp_engine *e, int nv, const double *V, long long ldv, const double *w, double *h, double sign, int guard_it,
                        int guard_refine, const double *inv = nullptr) {
  MdotArgs a{};
  a.nb = e->nb; a.nv = nv; a.ld = ldv; a.V = V; a.w = w; a.h = h; a.sign = sign; a.ctl = e->ctl; a.inv = inv;
  a.guard_it = guard_it; a.guard_refine = guard_refine;
  int ngroups = (nv + 7) / 8;
  a.per_group = (nv + ngroups - 1) / ngroups;
  ngroups = (nv + a.per_group - 1) / a.per_group;
  int per_sm = std::max(1, 8 / ngroups);
  e->prof_begin(1, 8.0 * e->nb * (nv + 1));
  if (a.per_group <= 2) {
    dim3 grid(grid_for((long long)e->nb / 16, per_sm), ngroups);
    k_mdot<2, 8><<<grid, MSPK_THREADS, 0, e->st>>>(a, e->ws);
  } else if (a.per_group <= 4) {
    dim3 grid(grid_for((long long)e->nb / 8, per_sm), ngroups);
    k_mdot<4, 4><<<grid, MSPK_THREADS, 0, e->st>>>(a, e->ws);
  } else {
    dim3 grid(grid_for((long long)e->nb / 4, per_sm), ngroups);
    k_mdot<8, 2><<<grid, MSPK_THREADS, 0, e->st>>>(a, e->ws);
  }
  e->prof_end();
  e->launches++;
}

template <int FIN>
static void launch_maxpy(msp_engine *e, int nv, const double *V, long long ldv, const double *coef, double *w, double *norm_out,
                         int guard_it, int guard_refine, int pass, int ws_slot, const double *inv = nullptr) {
  MaxpyArgs a{};
  a.nb = e->nb; a.nv = nv; a.ld = ldv; a.V = V; a.coef = coef; a.w = w; a.norm_out = norm_out; a.ctl = e->ctl; a.inv = inv;
  a.guard_it = guard_it; a.guard_refine = guard_refine; a.pass = pass;
  e->prof_begin(2, 8.0 * e->nb * (nv + 2));
  k_maxpy_norm<FIN><<<grid_for((long long)e->nb / 4, 8), MSPK_THREADS, 0, e->st>>>(a, e->ws, ws_slot);
  e->prof_end();
  e->launches++;
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
static int engine_free(msp_engine *e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  if (e->st) cudaStreamSynchronize(e->st);
  for (auto &kv : e->cycle_graphs) cudaGraphExecDestroy(kv.second.exec);
  for (int J = 0; J < MSP_MAX_BLOCKS; J++)
    if (e->peer_any[J].base && e->peer_any_ipc[J]) cudaIpcCloseMemHandle(e->peer_any[J].base);
  void *ptrs[] = {e->rp, e->ci, e->va, e->ecol, e->eval, e->brow, e->b, e->rhs, e->x, e->halo[0], e->halo[1], e->V, e->Wb[0],
                  e->Wb[1], e->S, e->Slo, e->Shi, e->R, e->ctl, e->ws.partial, e->ws.counter, e->dsc, e->dfac, e->gram_partial, e->win.base, e->cd, e->aint, e->dec};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (e->hsc) cudaFreeHost(e->hsc);
  if (e->own_comm && e->comm) delete e->comm;
  if (e->st) cudaStreamDestroy(e->st);
  delete e;
  return 0;
}

static int engine_create(const msp_problem *p, int device, msp_engine **out) {
  if (!p || !out) MSP_FAIL("null argument");
  if (p->dim != 2 && p->dim != 3) MSP_FAIL("dim must be 2 or 3");
  if (p->nblocks < 1 || p->nblocks > MSP_MAX_BLOCKS || p->block < 0 || p->block >= p->nblocks) MSP_FAIL("bad block / nblocks");
  if (p->max_restart < 1 || p->max_restart > MSP_MAX_RESTART) MSP_FAIL("max_restart out of range (1..64)");
  if (p->s < 0 || p->s > MSP_MAX_S) MSP_FAIL("s out of range (0..32)");
  RC(set_device(device));
  msp_engine *e = new msp_engine();
  e->device = device; e->prob = *p;
  // grid geometry.  2-D (poisson2DMatrix): row Ii = i*n + j, strip = whole grid lines.
  // 3-D (poisson3DMatrix): row = i + j*m + k*m*n, strip = whole z planes.
  int nx, ny, nz;
  long long layers;
  if (p->dim == 2) { nx = p->n; ny = p->m; nz = 1; layers = p->m; e->H = p->n; }
  else { nx = p->m; ny = p->n; nz = p->p; layers = p->p; e->H = p->m * p->n; }
  e->ntot = (long long)nx * ny * nz;
  if (layers % p->nblocks) { delete e; MSP_FAIL("grid lines (2-D) / planes (3-D) must be divisible by the number of blocks"); }
  if (e->ntot / p->nblocks > 2000000000LL) { delete e; MSP_FAIL("block too large for 32-bit indices"); }
  e->nb = (int)(e->ntot / p->nblocks);
  e->off = (int)((long long)e->nb * p->block);
  e->ld = ((long long)e->nb + 63) / 64 * 64;
  e->has_nb[0] = p->block > 0; e->has_nb[1] = p->block < p->nblocks - 1;
  if (cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess) { delete e; MSP_FAIL("stream create failed"); }
  int rc = 0;
  auto fail = [&](int r) { engine_free(e); return r; };
  // ---- matrix: assemble strip (K12), convert to ELL, list boundary rows ----
  rc = assemble_strip_dev(p->dim, nx, ny, nz, (long long)e->off, e->nb, e->st, &e->rp, &e->ci, &e->va, &e->nnz);
  if (rc) return fail(rc);
  e->W = (p->dim == 2) ? 5 : 7;
  if (cudaMalloc(&e->ecol, sizeof(int) * (size_t)e->ld * e->W) != cudaSuccess ||
      cudaMalloc(&e->eval, sizeof(double) * (size_t)e->ld * e->W) != cudaSuccess) { g_err = "out of device memory (ELL)"; return fail(1); }
  k_csr_to_ell<<<grid_for(e->ld, 16), MSPK_THREADS, 0, e->st>>>(e->nb, e->W, e->ld, e->off, e->rp, e->ci, e->va, e->ecol, e->eval);
  {
    int *flag = nullptr;
    if (cudaMalloc(&flag, sizeof(int) * ((size_t)e->nb + 1)) != cudaSuccess) { g_err = "oom"; return fail(1); }
    cudaMemsetAsync(flag, 0, sizeof(int) * ((size_t)e->nb + 1), e->st);
    k_mark_boundary<<<grid_for(e->nb, 16), MSPK_THREADS, 0, e->st>>>(e->nb, e->W, e->ld, e->ecol, flag);
    cudaStreamSynchronize(e->st);
    // boundary rows live in the first and last H rows of a strip: fetch only those flags
    std::vector<int> hf;
    std::vector<int> rows;
    int span = std::min(e->nb, e->H);
    hf.resize(span);
    cudaMemcpy(hf.data(), flag, sizeof(int) * span, cudaMemcpyDeviceToHost);
    for (int i = 0; i < span; i++) if (hf[i]) rows.push_back(i);
    if (e->nb > span) {
      int start = std::max(span, e->nb - span);
      cudaMemcpy(hf.data(), flag + start, sizeof(int) * (e->nb - start), cudaMemcpyDeviceToHost);
      for (int i = 0; i < e->nb - start; i++) if (hf[i]) rows.push_back(start + i);
    }
    cudaFree(flag);
    e->nbrow = (int)rows.size();
    if (e->nbrow) {
      if (cudaMalloc(&e->brow, sizeof(int) * rows.size()) != cudaSuccess) { g_err = "oom"; return fail(1); }
      cudaMemcpy(e->brow, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice);
    }
  }
  if (!p->keep_csr) {
    cudaFree(e->ci); cudaFree(e->va); e->ci = nullptr; e->va = nullptr; // rowptr kept (small)
  }
  // ---- vectors ----
  e->nvec = p->max_restart + 1;
  e->smax = p->s;
  size_t vb = sizeof(double) * (size_t)e->ld;
  bool ok = true;
  auto dalloc = [&](double **ptr, size_t bytes) { if (ok && cudaMalloc(ptr, bytes) != cudaSuccess) ok = false; if (ok) cudaMemsetAsync(*ptr, 0, bytes, e->st); };
  dalloc(&e->b, vb); dalloc(&e->rhs, vb); dalloc(&e->x, vb);
  dalloc(&e->halo[0], sizeof(double) * (size_t)e->H); dalloc(&e->halo[1], sizeof(double) * (size_t)e->H);
  dalloc(&e->V, vb * e->nvec); dalloc(&e->Wb[0], vb); dalloc(&e->Wb[1], vb);
  if (e->smax > 0) {
    dalloc(&e->S, vb * e->smax); dalloc(&e->R, vb * (e->smax + 1));
    dalloc(&e->Slo, sizeof(double) * (size_t)e->H * e->smax); dalloc(&e->Shi, sizeof(double) * (size_t)e->H * e->smax);
  }
  dalloc(&e->ws.partial, sizeof(double) * (size_t)MSPK_MAX_PART * 8 * 24);
  dalloc(&e->dsc, sizeof(double) * 256);
  dalloc(&e->dfac, sizeof(double) * (size_t)p->nblocks * (e->smax + 1) * (e->smax + 1) + 64);
  if (e->smax > 0) dalloc(&e->gram_partial, sizeof(double) * 45 * MSPK_MAX_PART);
  e->use_cholqr = getenv("MSPLIT_NO_CHOLQR") == nullptr;
  e->win.H = e->H; e->win.G = p->nblocks; e->win.fslot = Window::fslot_for(e->smax);
  e->win.bytes = Window::size_for(e->H, p->nblocks, e->smax);
  dalloc(&e->win.base, e->win.bytes);
  if (ok && cudaMalloc(&e->aint, sizeof(int) * 16) != cudaSuccess) ok = false;
  if (ok) cudaMemsetAsync(e->aint, 0, sizeof(int) * 16, e->st);
  if (ok && cudaMalloc(&e->dec, sizeof(ProbeDecision) * 2) != cudaSuccess) ok = false;
  if (ok) cudaMemsetAsync(e->dec, 0, sizeof(ProbeDecision) * 2, e->st);
  e->fcache.assign((size_t)p->nblocks * e->win.fslot, 0.0);
  e->fcache_seq.assign(p->nblocks, 0);
  if (ok && cudaMalloc(&e->ws.counter, sizeof(unsigned) * 64) != cudaSuccess) ok = false;
  if (ok) cudaMemsetAsync(e->ws.counter, 0, sizeof(unsigned) * 64, e->st);
  if (ok && cudaMalloc(&e->ctl, sizeof(GmresCtl)) != cudaSuccess) ok = false;
  if (ok) cudaMemsetAsync(e->ctl, 0, sizeof(GmresCtl), e->st);
  if (ok && cudaMalloc(&e->cd, sizeof(CdState)) != cudaSuccess) ok = false;
  if (ok) cudaMemsetAsync(e->cd, 0, sizeof(CdState), e->st);
  if (ok && cudaMallocHost(&e->hsc, sizeof(double) * 256) != cudaSuccess) ok = false;
  if (!ok) { g_err = "out of device memory (vectors)"; return fail(1); }
  e->comm = new SelfComm(); e->own_comm = true;
  e->use_graphs = getenv("MSPLIT_NO_GRAPHS") == nullptr;
  // ---- b_K = A_K,: * 1 (computeTheRightHandSideWithInitialGuess utils.c:626): halos of ones ----
  {
    k_fill<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, 1.0, e->Wb[0]);
    k_fill<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, 1.0, e->halo[0]);
    k_fill<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, 1.0, e->halo[1]);
    SpmvArgs a = spmv_args(e, e->Wb[0], e->b);
    a.lo = e->has_nb[0] ? e->halo[0] : nullptr; a.hi = e->has_nb[1] ? e->halo[1] : nullptr;
    launch_spmv_w<1, false, false, false>(e, a, 0, nullptr);
    cudaMemsetAsync(e->halo[0], 0, sizeof(double) * e->H, e->st);
    cudaMemsetAsync(e->halo[1], 0, sizeof(double) * e->H, e->st);
    cudaMemsetAsync(e->Wb[0], 0, vb, e->st);
    cudaMemcpyAsync(e->rhs, e->b, vb, cudaMemcpyDeviceToDevice, e->st);
  }
  if (cudaStreamSynchronize(e->st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { g_err = "setup kernels failed"; return fail(1); }
  e->launches = 0;
  *out = e;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// operator surface
// ------------------------------------------------------------------------------------------------
static int op_update_rhs(msp_engine *e) {
  if (e->nbrow == 0) return 0;
  k_update_rhs<<<grid_for(e->nbrow), MSPK_THREADS, 0, e->st>>>(e->nbrow, e->brow, e->nb, e->W, e->H, e->ld, e->ecol, e->eval,
                                                               e->has_nb[0] ? e->halo[0] : nullptr, e->has_nb[1] ? e->halo[1] : nullptr,
                                                               e->b, e->rhs);
  e->launches++;
  return 0;
}

// sum of squares of (rhs - A_KK x) into dsc[slot]; strip variant: (b - A_K,: [halo|x|halo])
static int op_resid_sumsq(msp_engine *e, bool strip, int dsc_slot) {
  SpmvArgs a = spmv_args(e, e->x, e->Wb[1]);
  a.b = strip ? e->b : e->rhs;
  if (strip) {
    a.lo = e->has_nb[0] ? e->halo[0] : nullptr; a.hi = e->has_nb[1] ? e->halo[1] : nullptr;
    launch_spmv_w<1, true, false, true>(e, a, 1, nullptr);
  } else {
    launch_spmv_w<0, true, false, true>(e, a, 1, nullptr);
  }
  CK(cudaMemcpyAsync(e->dsc + dsc_slot, e->ws.partial + 1 * MSPK_MAX_PART + MSPK_MAX_PART - 1, sizeof(double), cudaMemcpyDeviceToDevice, e->st));
  return 0;
}

static int read_scalars(msp_engine *e, int first, int n) {
  CK(cudaMemcpyAsync(e->hsc + first, e->dsc + first, sizeof(double) * n, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}

// inner_solver utils.c:950-970 -> KSPSolve_GMRES.  One host synchronisation per
// restart cycle; inside a cycle every decision is taken on the device.
static int op_inner_solve(msp_engine *e, const msp_ksp_opts *o, bool publish, int *its_out, int *reason_out, double *rnorm_out) {
  if (o->restart < 1 || o->restart > e->prob.max_restart) MSP_FAIL("restart exceeds max_restart of the engine");
  const bool guess_zero = !o->guess_nonzero;
  double *bnorm_sq = nullptr;
  if (!guess_zero && !o->initial_rtol) {
    k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, 0.0, e->ws, 2, e->dsc + 8);
    e->launches++;
    bnorm_sq = e->dsc + 8;
  }
  k_ctl_begin<<<1, 32, 0, e->st>>>(e->ctl, o->restart, o->max_it, o->min_it, o->initial_rtol, guess_zero ? 1 : 0, o->cgs_refine, o->rtol,
                                   o->abstol, o->divtol, bnorm_sq);
  e->launches++;
  int itcount = 0;
  bool first = true;
  struct { int its, it, reason, active; } hc{};
  double *peer_lo = (publish && e->peer[0].base) ? e->peer[0].halo(1, e->par) : nullptr; // lower neighbour's "hi" window
  double *peer_hi = (publish && e->peer[1].base) ? e->peer[1].halo(0, e->par) : nullptr; // upper neighbour's "lo" window
  while (true) {
    const int nsteps = std::min(o->restart, o->max_it - itcount);
    const bool from_rhs = first && guess_zero;
    // everything one restart cycle enqueues: prologue, nsteps Arnoldi steps, solution update, 16-byte status read-back
    auto enqueue_cycle = [&]() -> int {
      // The Krylov basis is stored UN-normalised: vtilde_0 = r, vtilde_(it+1) = orthogonalised A v_it, with
      // v_j = vtilde_j * inv_arr[j]: the SpMV scales its gathered input (bit-identical to a stored normalised vector),
      // MDot scales the reduced value and MAXPY / the solution update fold inv_j into their coefficients.  This
      // removes VecNormalize's write pass (K5) and the scratch vectors from the Arnoldi step.
      // ---- cycle prologue: vtilde_0 = rhs - A x (or rhs), ||r|| and the cycle-begin logic on the device ----
      double *V0 = e->V;
      if (from_rhs) {
        k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, V0);
        k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, 0.0, e->ws, 2, e->dsc + 9);
        k_ctl_cycle_begin_from<<<1, 32, 0, e->st>>>(e->ctl, e->dsc + 9);
        e->launches += 3;
      } else {
        SpmvArgs a = spmv_args(e, e->x, V0);
        a.b = e->rhs;
        launch_spmv_w<0, true, false, true>(e, a, 0, e->ctl);
      }
      double *lhh = reinterpret_cast<double *>(reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, lhh));
      const double *invs = reinterpret_cast<const double *>(reinterpret_cast<const char *>(e->ctl) + offsetof(GmresCtl, inv_arr));
      for (int it = 0; it < nsteps; it++) {
        // w = A v_it (K1), reading vtilde_it scaled on the fly, written straight into the slot of vtilde_(it+1)
        double *w = e->V + (long long)(it + 1) * e->ld;
        SpmvArgs a = spmv_args(e, e->V + (long long)it * e->ld, w);
        a.guard_it = it;
        launch_spmv_w<0, false, true, false>(e, a, 0, nullptr);
        if (o->mgs) {
          // -ksp_gmres_modifiedgramschmidt: it+1 sequential (dot, axpy) pairs; the last axpy closes the step
          for (int j = 0; j <= it; j++) {
            const double *vj = e->V + (long long)j * e->ld;
            launch_mdot(e, 1, vj, e->ld, w, lhh + j, -1.0, it, 0, invs + j);
            if (j < it) launch_maxpy<0>(e, 1, vj, e->ld, lhh + j, w, nullptr, it, 0, 0, 3, invs + j);
            else launch_maxpy<1>(e, 1, vj, e->ld, lhh + j, w, nullptr, it, 0, 0, 0, invs + j);
          }
          continue;
        }
        // classical Gram-Schmidt: lhh = -V^T w (K3); w += V lhh, ||w|| (K4+K5), Hessenberg + test (K6)
        launch_mdot(e, it + 1, e->V, e->ld, w, lhh, -1.0, it, 0, invs);
        launch_maxpy<1>(e, it + 1, e->V, e->ld, lhh, w, nullptr, it, 0, 0, 0, invs);
        if (o->cgs_refine) {
          launch_mdot(e, it + 1, e->V, e->ld, w, lhh, -1.0, it, 1, invs);
          launch_maxpy<1>(e, it + 1, e->V, e->ld, lhh, w, nullptr, it, 1, 1, 0, invs);
        }
      }
      // ---- KSPGMRESBuildSoln + boundary publication ----
      k_build_soln_coef<<<1, 32, 0, e->st>>>(e->ctl);
      UpdateXArgs u{};
      u.nb = e->nb; u.H = e->H; u.ld = e->ld; u.V = e->V; u.x = e->x; u.ctl = e->ctl; u.peer_lo = peer_lo; u.peer_hi = peer_hi;
      k_update_x<<<grid_for(e->nb, 8), MSPK_THREADS, 0, e->st>>>(u);
      e->launches += 2;
      CK(cudaMemcpyAsync(e->hsc + 32, reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, its), 16, cudaMemcpyDeviceToHost, e->st));
      return 0;
    };
    if (e->use_graphs && !e->prof) {
      const msp_engine::CycleKey key(nsteps, o->cgs_refine + 4 * (o->mgs ? 1 : 0), from_rhs ? 1 : 0, (const void *)e->V, (const void *)peer_lo, (const void *)peer_hi);
      auto itg = e->cycle_graphs.find(key);
      if (itg == e->cycle_graphs.end()) {
        const int64_t l0 = e->launches;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_cycle();
        cudaError_t ce = cudaStreamEndCapture(e->st, &graph);
        if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (!rc) MSP_FAIL(std::string("stream capture failed: ") + cudaGetErrorString(ce)); return rc; }
        msp_engine::CycleGraph cg{nullptr, (int)(e->launches - l0)};
        e->launches = l0;
        ce = cudaGraphInstantiate(&cg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) MSP_FAIL(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
        if (e->cycle_graphs.size() > 256) { for (auto &kv : e->cycle_graphs) cudaGraphExecDestroy(kv.second.exec); e->cycle_graphs.clear(); }
        itg = e->cycle_graphs.emplace(key, cg).first;
      }
      CK(cudaGraphLaunch(itg->second.exec, e->st));
      e->launches += itg->second.launches;
    } else {
      RC(enqueue_cycle());
    }
    first = false;
    CK(cudaStreamSynchronize(e->st));
    memcpy(&hc, e->hsc + 32, 16);
    itcount += hc.it;
    if (hc.reason) break;
    if (itcount >= o->max_it) { hc.reason = MSP_DIVERGED_ITS; break; }
  }
  if (its_out) *its_out = hc.its;
  if (reason_out) *reason_out = hc.reason;
  if (rnorm_out) {
    CK(cudaMemcpyAsync(e->hsc + 40, reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, ksp_rnorm), 8, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    *rnorm_out = e->hsc[40];
  }
  return 0;
}

// after the synchronising barrier: copy the freshly received boundary layers into the private halos
static int op_collect_halos(msp_engine *e) {
  for (int side = 0; side < 2; side++)
    if (e->has_nb[side])
      CK(cudaMemcpyAsync(e->halo[side], e->win.halo(side, e->par), sizeof(double) * e->H, cudaMemcpyDeviceToDevice, e->st));
  e->par ^= 1;
  return 0;
}
static int op_publish_boundary(msp_engine *e) {
  double *peer_lo = e->peer[0].base ? e->peer[0].halo(1, e->par) : nullptr;
  double *peer_hi = e->peer[1].base ? e->peer[1].halo(0, e->par) : nullptr;
  if (!peer_lo && !peer_hi) return 0;
  k_publish_boundary<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->nb, e->H, e->x, peer_lo, peer_hi);
  e->launches++;
  return 0;
}

static int op_push_iterate(msp_engine *e, int t) {
  if (t < 0 || t >= e->smax) MSP_FAIL("basis index out of range");
  k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, e->x, e->S + (long long)t * e->ld);
  e->launches++;
  CK(cudaMemcpyAsync(e->Slo + (size_t)t * e->H, e->halo[0], sizeof(double) * e->H, cudaMemcpyDeviceToDevice, e->st));
  CK(cudaMemcpyAsync(e->Shi + (size_t)t * e->H, e->halo[1], sizeof(double) * e->H, cudaMemcpyDeviceToDevice, e->st));
  return 0;
}

// kind: MSP_ALG_*_GLOBAL / SEMI_LOCAL use the strip with stored boundaries, *_LOCAL uses A_KK
static bool kind_is_local(int kind) { return kind == MSP_ALG_SMSM_LOCAL || kind == MSP_ALG_AMAM_LOCAL; }

static int op_spmm(msp_engine *e, int kind, int s, bool diff_basis = true) {
  // basis of successive corrections (same span as the iterates, far better conditioned); the LSQR path keeps the
  // reference's raw basis [x^1 .. x^s]
  if (diff_basis) { k_diff_basis<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, s, e->ld, e->S); e->launches++; }
  if (diff_basis && !kind_is_local(kind)) {
    if (e->has_nb[0]) { k_diff_basis<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, s, e->H, e->Slo); e->launches++; }
    if (e->has_nb[1]) { k_diff_basis<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, s, e->H, e->Shi); e->launches++; }
  }
  SpmmArgs a{};
  a.nb = e->nb; a.W = e->W; a.H = e->H; a.s = s; a.ld = e->ld; a.lds = e->ld; a.ecol = e->ecol; a.eval = e->eval;
  a.S = e->S; a.R = e->R;
  const bool local = kind_is_local(kind);
  a.Slo = (!local && e->has_nb[0]) ? e->Slo : nullptr;
  a.Shi = (!local && e->has_nb[1]) ? e->Shi : nullptr;
  const int g = grid_for(e->nb, 8);
  for (int c0 = 0; c0 < s;) {
    int nc = std::min(8, s - c0);
    // chunk sizes 8,5,4,2,1 cover every s with few passes over the matrix
    int use = nc >= 8 ? 8 : nc >= 5 ? 5 : nc >= 4 ? 4 : nc >= 2 ? 2 : 1;
#define SPMM_CASE(N)                                                                                          \
  case N:                                                                                                     \
    if (local) k_spmm_ell<0, N><<<g, MSPK_THREADS, 0, e->st>>>(a, c0);                                        \
    else k_spmm_ell<1, N><<<g, MSPK_THREADS, 0, e->st>>>(a, c0);                                              \
    break;
    switch (use) { SPMM_CASE(8) SPMM_CASE(5) SPMM_CASE(4) SPMM_CASE(2) SPMM_CASE(1) }
#undef SPMM_CASE
    e->launches++;
    c0 += use;
  }
  return 0;
}

// upper Cholesky factor of a symmetric NC x NC matrix given by its upper triangle (column-major); false on breakdown
static bool chol_upper(int nc, const double *G, double *U) {
  std::fill(U, U + nc * nc, 0.0);
  double dmax = 0.0;
  for (int j = 0; j < nc; j++) dmax = std::max(dmax, G[j * nc + j]);
  for (int j = 0; j < nc; j++) {
    for (int i = 0; i <= j; i++) {
      double t = G[j * nc + i];
      for (int k = 0; k < i; k++) t -= U[i * nc + k] * U[j * nc + k];
      if (i < j) U[j * nc + i] = t / U[i * nc + i];
      else {
        if (!(t > 1e-13 * dmax)) return false; // not safely positive definite at working precision
        U[j * nc + j] = std::sqrt(t);
      }
    }
  }
  return true;
}

template <int NC>
static void launch_gram_nc(msp_engine *e, const double *C, double *out_dev) {
  auto k = k_gram<NC>;
  k<<<grid_for((long long)e->nb / 2, std::min(resident_blocks_per_sm(k), 4)), MSPK_THREADS, 0, e->st>>>(e->nb, e->ld, C, e->gram_partial, e->ws.counter + 40, out_dev);
  e->launches++;
}
template <int NC>
static void launch_trsolve_nc(msp_engine *e, double *C, const double *U_dev) {
  auto k = k_right_trsolve<NC>;
  k<<<grid_for((long long)e->nb / 2, resident_blocks_per_sm(k)), MSPK_THREADS, 0, e->st>>>(e->nb, e->ld, C, U_dev);
  e->launches++;
}
static void launch_gram(msp_engine *e, int nc, const double *C, double *out_dev) {
  switch (nc) {
    case 2: launch_gram_nc<2>(e, C, out_dev); break; case 3: launch_gram_nc<3>(e, C, out_dev); break;
    case 4: launch_gram_nc<4>(e, C, out_dev); break; case 5: launch_gram_nc<5>(e, C, out_dev); break;
    case 6: launch_gram_nc<6>(e, C, out_dev); break; case 7: launch_gram_nc<7>(e, C, out_dev); break;
    case 8: launch_gram_nc<8>(e, C, out_dev); break; default: launch_gram_nc<9>(e, C, out_dev); break;
  }
}
static void launch_trsolve(msp_engine *e, int nc, double *C, const double *U_dev) {
  switch (nc) {
    case 2: launch_trsolve_nc<2>(e, C, U_dev); break; case 3: launch_trsolve_nc<3>(e, C, U_dev); break;
    case 4: launch_trsolve_nc<4>(e, C, U_dev); break; case 5: launch_trsolve_nc<5>(e, C, U_dev); break;
    case 6: launch_trsolve_nc<6>(e, C, U_dev); break; case 7: launch_trsolve_nc<7>(e, C, U_dev); break;
    case 8: launch_trsolve_nc<8>(e, C, U_dev); break; default: launch_trsolve_nc<9>(e, C, U_dev); break;
  }
}

// CGS2 leaf (classical Gram-Schmidt with reorthogonalisation, the Arnoldi kernels K3, K4+K5) on the nc columns at e->R
static int local_qr_cgs2(msp_engine *e, int nc, std::vector<double> &U) {
  U.assign((size_t)nc * nc, 0.0);
  for (int c = 0; c < nc; c++) {
    double *q = e->R + (long long)c * e->ld;
    if (c > 0) {
      launch_mdot(e, c, e->R, e->ld, q, e->dsc + 64, -1.0, -1, 0);
      launch_maxpy<0>(e, c, e->R, e->ld, e->dsc + 64, q, e->dsc + 200, -1, 0, 0, 3);
      launch_mdot(e, c, e->R, e->ld, q, e->dsc + 128, -1.0, -1, 0);
      launch_maxpy<0>(e, c, e->R, e->ld, e->dsc + 128, q, e->dsc + 200, -1, 0, 0, 3);
    } else {
      launch_maxpy<0>(e, 0, e->R, e->ld, e->dsc + 64, q, e->dsc + 200, -1, 0, 0, 3);
    }
    k_scale_by_inv<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->dsc + 200, q);
    e->launches++;
    CK(cudaMemcpyAsync(e->hsc + 64, e->dsc + 64, sizeof(double) * 140, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    for (int j = 0; j < c; j++) U[(size_t)c * nc + j] = -(e->hsc[64 + j] + e->hsc[128 + j]);
    U[(size_t)c * nc + c] = e->hsc[200];
  }
  return 0;
}

// TSQR leaf: the (s+1)x(s+1) upper factor of [R_K | rhs] (column-major, to the host).
//  * s <= 8: CholeskyQR2 — Gram contraction (K9, one pass), Cholesky on the host, C := C U1^{-1} (one pass), Gram again,
//    U = U2 U1: 24 n (s+1) bytes instead of the ~16 n (s+1)(s+3) of Gram-Schmidt, and as accurate as Householder QR
//    while cond([R|rhs]) < ~1e7; a Cholesky breakdown falls back to
//  * CGS2 with the Arnoldi kernels (any s, any conditioning), continuing from whatever basis is in place.
static int op_local_qr(msp_engine *e, int kind, int s, double *u_aug /* host (s+1)^2 */) {
  const int nc = s + 1;
  const double *rhs_src = kind_is_local(kind) ? e->rhs : e->b;
  k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, rhs_src, e->R + (long long)s * e->ld);
  e->launches++;
  std::vector<double> U, U1, U2, G((size_t)nc * nc, 0.0);
  bool have_u1 = false;
  if (e->use_cholqr && nc >= 2 && nc <= 9) {
    U1.assign((size_t)nc * nc, 0.0); U2.assign((size_t)nc * nc, 0.0);
    double *Gdev = e->dfac; // idle between TSQR gathers; (smax+1)^2 doubles fit
    launch_gram(e, nc, e->R, Gdev);
    CK(cudaMemcpyAsync(G.data(), Gdev, sizeof(double) * nc * nc, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    if (chol_upper(nc, G.data(), U1.data())) {
      CK(cudaMemcpyAsync(Gdev, U1.data(), sizeof(double) * nc * nc, cudaMemcpyHostToDevice, e->st));
      launch_trsolve(e, nc, e->R, Gdev);
      have_u1 = true;
      launch_gram(e, nc, e->R, Gdev);
      CK(cudaMemcpyAsync(G.data(), Gdev, sizeof(double) * nc * nc, cudaMemcpyDeviceToHost, e->st));
      CK(cudaStreamSynchronize(e->st));
      if (chol_upper(nc, G.data(), U2.data())) {
        // U = U2 U1
        for (int j = 0; j < nc; j++)
          for (int i = 0; i <= j; i++) {
            double t = 0.0;
            for (int k = i; k <= j; k++) t += U2[(size_t)k * nc + i] * U1[(size_t)j * nc + k];
            u_aug[(size_t)j * nc + i] = t;
          }
        for (int j = 0; j < nc; j++) for (int i = j + 1; i < nc; i++) u_aug[(size_t)j * nc + i] = 0.0;
        return 0;
      }
    }
  }
  RC(local_qr_cgs2(e, nc, U));
  if (have_u1) {
    // the columns in place were C U1^{-1}: overall factor = U_cgs2 U1
    for (int j = 0; j < nc; j++)
      for (int i = 0; i < nc; i++) {
        double t = 0.0;
        for (int k = i; k <= j; k++) t += U[(size_t)k * nc + i] * U1[(size_t)j * nc + k];
        u_aug[(size_t)j * nc + i] = (i <= j) ? t : 0.0;
      }
  } else {
    memcpy(u_aug, U.data(), sizeof(double) * (size_t)nc * nc);
  }
  return 0;
}

static int op_apply_alpha(msp_engine *e, int kind, int s, const double *alpha_host) {
  memcpy(e->hsc + 216, alpha_host, sizeof(double) * s);
  CK(cudaMemcpyAsync(e->dsc + 216, e->hsc + 216, sizeof(double) * s, cudaMemcpyHostToDevice, e->st));
  k_lincomb<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, s, e->ld, e->S, e->dsc + 216, e->x);
  e->launches++;
  if (!kind_is_local(kind)) {
    // the block's copies of the neighbours' boundaries follow x_min = S alpha too (…-semi-local.c:335-338)
    for (int side = 0; side < 2; side++)
      if (e->has_nb[side]) {
        k_lincomb<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, s, e->H, side ? e->Shi : e->Slo, e->dsc + 216, e->halo[side]);
        e->launches++;
      }
  }
  return 0;
}

static int allreduce_host(msp_engine *e, int first, int n);

// KSPSolve_LSQR (PETSc lsqr.c) on the dense column block R_K (n_K x s) distributed over the blocks:
// R v and the vector updates are block-local kernels (K10 lincomb, K4 axpy+norm), R^T u is the MDot kernel (K3);
// with `global` the s-vector R^T u and the two norms of every iteration are summed over blocks (the reference's
// MatMultTranspose_MPIDense / VecNorm_MPI allreduces).  Zero initial guess, initial-residual-norm test
// (outer_solver_norm_equation utils.c:1065-1068); returns alpha, phibar and the iteration count.
static int op_lsqr(msp_engine *e, int kind, int s, bool global, int max_it, double rtol, double abstol, double *alpha, double *rnorm_out,
                   int *its_out) {
  const double *rhs_src = kind_is_local(kind) ? e->rhs : e->b;
  double *U = e->Wb[0], *U1 = e->Wb[1];
  std::vector<double> V(s, 0.0), V1(s, 0.0), W(s, 0.0);
  auto sum_blocks = [&](int first, int n) -> int { return global ? allreduce_host(e, first, n) : read_scalars(e, first, n); };
  auto put_s = [&](const std::vector<double> &v) -> int {
    memcpy(e->hsc + 216, v.data(), sizeof(double) * s);
    CK(cudaMemcpyAsync(e->dsc + 216, e->hsc + 216, sizeof(double) * s, cudaMemcpyHostToDevice, e->st));
    return 0;
  };
  for (int j = 0; j < s; j++) alpha[j] = 0.0;
  k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, rhs_src, U);
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, U, 0.0, e->ws, 2, e->dsc + 4);
  e->launches += 2;
  RC(sum_blocks(4, 1));
  double rnorm = std::sqrt(e->hsc[4]);
  const double rnorm0 = rnorm, ttol = std::max(rtol * rnorm0, abstol);
  int its = 0;
  if (rnorm > 0.0) {
    double beta = rnorm, al;
    k_scale<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, 1.0 / beta, U);
    e->launches++;
    launch_mdot(e, s, e->R, e->ld, U, e->dsc + 64, 1.0, -1, 0);
    RC(sum_blocks(64, s));
    double nv = 0.0;
    for (int j = 0; j < s; j++) { V[j] = e->hsc[64 + j]; nv += V[j] * V[j]; }
    al = std::sqrt(nv);
    if (al > 0.0) for (int j = 0; j < s; j++) V[j] /= al;
    W = V;
    double phibar = beta, rhobar = al;
    int i = 0;
    do {
      RC(put_s(V));
      k_lincomb<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, s, e->ld, e->R, e->dsc + 216, U1); // U1 = R V
      e->launches++;
      e->hsc[5] = -al;
      CK(cudaMemcpyAsync(e->dsc + 5, e->hsc + 5, sizeof(double), cudaMemcpyHostToDevice, e->st));
      launch_maxpy<0>(e, 1, U, e->ld, e->dsc + 5, U1, e->dsc + 6, -1, 0, 0, 3); // U1 -= alpha U, ||U1||
      RC(read_scalars(e, 6, 1));
      e->hsc[6] = e->hsc[6] * e->hsc[6];
      CK(cudaMemcpyAsync(e->dsc + 6, e->hsc + 6, sizeof(double), cudaMemcpyHostToDevice, e->st));
      RC(sum_blocks(6, 1));
      beta = std::sqrt(e->hsc[6]);
      if (beta > 0.0) { k_scale<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, 1.0 / beta, U1); e->launches++; }
      launch_mdot(e, s, e->R, e->ld, U1, e->dsc + 64, 1.0, -1, 0); // V1 = R^T U1
      RC(sum_blocks(64, s));
      nv = 0.0;
      for (int j = 0; j < s; j++) { V1[j] = std::fma(-beta, V[j], e->hsc[64 + j]); nv += V1[j] * V1[j]; }
      al = std::sqrt(nv);
      if (al > 0.0) for (int j = 0; j < s; j++) V1[j] /= al;
      const double rho = std::sqrt(rhobar * rhobar + beta * beta);
      const double c = rhobar / rho, sn = beta / rho, theta = sn * al;
      rhobar = -c * al;
      const double phi = c * phibar;
      phibar = sn * phibar;
      for (int j = 0; j < s; j++) alpha[j] = std::fma(phi / rho, W[j], alpha[j]);
      for (int j = 0; j < s; j++) W[j] = V1[j] + (-theta / rho) * W[j];
      rnorm = phibar;
      its++;
      // KSPConvergedDefault with -ksp_convergence_test default (running_bulk_test_g5k:247)
      bool conv = std::isnan(rnorm) || std::isinf(rnorm) || rnorm <= ttol || rnorm >= 1e4 * rnorm0;
      if (conv) break;
      std::swap(U, U1);
      std::swap(V, V1);
      i++;
    } while (i < max_it);
  }
  if (rnorm_out) *rnorm_out = rnorm;
  if (its_out) *its_out = its;
  return 0;
}

// Normal-equations minimiser (the reference's `outer_solver`, utils.c:972-996: MatTransposeMatMult(R,R), MatMultTranspose(R,b),
// KSPSolve on the s x s system): ONE Gram pass over [R_K | rhs] (K9), one allreduce of the (s+1)^2 Gram entries when the
// least squares is global ("NCCL Gram allreduce"), Cholesky on the host.  Squares the condition number: offered for
// completeness, TSQR stays the default.  ||b - R alpha|| from a second pass (not from b'b - g'alpha: cancellation).
static int op_normal_equations(msp_engine *e, int kind, int s, bool global, double *alpha, double *rnorm_out) {
  const int nc = s + 1;
  if (nc > 9) MSP_FAIL("the normal-equations minimiser supports s <= 8");
  const double *rhs_src = kind_is_local(kind) ? e->rhs : e->b;
  double *bcol = e->R + (long long)s * e->ld;
  k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, rhs_src, bcol);
  e->launches++;
  launch_gram(e, nc, e->R, e->dfac);
  if (global) RC(e->comm->allreduce_sum(e->dfac, nc * nc, e->st));
  std::vector<double> G((size_t)nc * nc), U((size_t)s * s), Gs((size_t)s * s);
  CK(cudaMemcpyAsync(G.data(), e->dfac, sizeof(double) * nc * nc, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  for (int j = 0; j < s; j++) for (int i = 0; i <= j; i++) Gs[(size_t)j * s + i] = G[(size_t)j * nc + i];
  if (!chol_upper(s, Gs.data(), U.data())) MSP_FAIL("normal equations: the Gram matrix is not positive definite at working precision (use the TSQR minimiser)");
  // U^T U alpha = g, g = R^T b = last column of the augmented Gram
  std::vector<double> y(s);
  for (int i = 0; i < s; i++) {
    double t = G[(size_t)s * nc + i];
    for (int k = 0; k < i; k++) t -= U[(size_t)i * s + k] * y[k];
    y[i] = t / U[(size_t)i * s + i];
  }
  for (int i = s - 1; i >= 0; i--) {
    double t = y[i];
    for (int k = i + 1; k < s; k++) t -= U[(size_t)k * s + i] * alpha[k];
    alpha[i] = t / U[(size_t)i * s + i];
  }
  if (rnorm_out) {
    // r = rhs - R alpha, in place in the rhs column
    for (int j = 0; j < s; j++) e->hsc[216 + j] = -alpha[j];
    CK(cudaMemcpyAsync(e->dsc + 216, e->hsc + 216, sizeof(double) * s, cudaMemcpyHostToDevice, e->st));
    launch_maxpy<0>(e, s, e->R, e->ld, e->dsc + 216, bcol, e->dsc + 6, -1, 0, 0, 3);
    RC(read_scalars(e, 6, 1));
    e->hsc[6] = e->hsc[6] * e->hsc[6];
    if (global) {
      CK(cudaMemcpyAsync(e->dsc + 6, e->hsc + 6, sizeof(double), cudaMemcpyHostToDevice, e->st));
      RC(allreduce_host(e, 6, 1));
    }
    *rnorm_out = std::sqrt(e->hsc[6]);
  }
  return 0;
}

// small dense least squares on the host: stack nfac upper factors [U_k | c_k; 0 rho_k] and solve by
// Householder QR.  This is the root of the TSQR tree (s <= 32: a few kflop).
static int tsqr_combine(int s, int nfac, const double *uall, double *alpha, double *resnorm) {
  const int nc = s + 1, rows = nfac * nc;
  std::vector<double> A((size_t)rows * nc, 0.0); // column-major rows x nc
  for (int f = 0; f < nfac; f++)
    for (int c = 0; c < nc; c++)
      for (int r = 0; r <= c; r++) A[(size_t)c * rows + f * nc + r] = uall[(size_t)f * nc * nc + (size_t)c * nc + r];
  std::vector<double> diag(nc, 0.0);
  for (int k = 0; k < nc; k++) {
    double *a = &A[(size_t)k * rows];
    double nrm = 0.0;
    for (int r = k; r < rows; r++) nrm += a[r] * a[r];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) { diag[k] = 0.0; continue; }
    double beta = (a[k] >= 0.0) ? -nrm : nrm;
    a[k] -= beta;
    double vtv = 0.0;
    for (int r = k; r < rows; r++) vtv += a[r] * a[r];
    for (int j = k + 1; j < nc; j++) {
      double *aj = &A[(size_t)j * rows];
      double d = 0.0;
      for (int r = k; r < rows; r++) d += a[r] * aj[r];
      d = 2.0 * d / vtv;
      for (int r = k; r < rows; r++) aj[r] -= d * a[r];
    }
    diag[k] = beta;
  }
  // back substitution on the leading s x s block against column s
  const double *cvec = &A[(size_t)s * rows];
  double dmax = 0.0;
  for (int k = 0; k < s; k++) dmax = std::max(dmax, std::fabs(diag[k]));
  for (int k = s - 1; k >= 0; k--) {
    double t = cvec[k];
    for (int j = k + 1; j < s; j++) t -= A[(size_t)j * rows + k] * alpha[j];
    // numerically dependent basis vector (iterates identical to rounding): drop it
    alpha[k] = (std::fabs(diag[k]) > 1e-14 * dmax) ? t / diag[k] : 0.0;
  }
  if (resnorm) *resnorm = std::fabs(diag[s]);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// the drivers' outer loops (one engine = one block; comm provides barrier / allreduce)
// ------------------------------------------------------------------------------------------------
static int exchange_sync(msp_engine *e) {
  // the boundary layers were already stored into the neighbours' windows by k_update_x / k_publish_boundary
  // (class 3 of the profile = barrier + collection of the received layers; bytes = what crossed NVLink into this block)
  e->prof_begin(3, 8.0 * e->H * ((e->has_nb[0] ? 1 : 0) + (e->has_nb[1] ? 1 : 0)));
  int rc = e->comm->barrier(e->st);
  if (!rc) rc = op_collect_halos(e);
  e->prof_end();
  return rc;
}

static int allreduce_host(msp_engine *e, int first, int n) {
  // sum dsc[first..first+n) over blocks and bring it to hsc
  RC(e->comm->allreduce_sum(e->dsc + first, n, e->st));
  return read_scalars(e, first, n);
}

static int engine_solve_sync(msp_engine *e, const msp_solve_opts *o, msp_result *res) {
  const int G = e->prob.nblocks, s = o->s;
  const int alg = o->alg;
  const double atol = 1e-100; // hard-coded absolute_tolerance (…-global.c:34)
  msp_ksp_opts in = o->inner;
  in.initial_rtol = 1; in.guess_nonzero = 1; // inner_solver utils.c:956-957
  if (alg != MSP_ALG_SM && (s < 1 || s > e->smax)) MSP_FAIL("s exceeds the engine's basis storage");
  const int max_outer = o->max_outer > 0 ? o->max_outer : 1000000;
  memset(res, 0, sizeof(*res));
  // global_norm_0 = computeFinalResidualNorm(x = 0) before the loop (…multisplitting.c:162) = ||b||; computed from b so
  // that a call continuing from a previous iterate keeps the same reference norm
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->b, 0.0, e->ws, 2, e->dsc + 0);
  e->launches++;
  RC(allreduce_host(e, 0, 1));
  res->norm0 = std::sqrt(e->hsc[0]);
  const double thr_global = std::max(atol, o->rtol * res->norm0);
  const double thr_local = std::max(atol, (o->rtol / std::sqrt((double)G)) * 1.0 * res->norm0);
  RC(e->comm->barrier(e->st)); // PetscBarrier before MPI_Wtime
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  CK(cudaEventRecord(ev0, e->st));
  const int64_t launches0 = e->launches;
  e->prof = o->profile != 0;
  bool done = false;
  int sticky = 0;
  const bool lsqr = o->outer_type == 1;
  const int lsqr_max_it = o->outer_max_it > 0 ? o->outer_max_it : 100;
  const double lsqr_rtol = o->outer_rtol > 0 ? o->outer_rtol : 1e-15, lsqr_abstol = o->outer_abstol > 0 ? o->outer_abstol : 1e-100;
  typedef std::chrono::steady_clock clk;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  std::vector<double> uaug((size_t)(s + 1) * (s + 1)), alpha(std::max(s, 1));
  if (alg == MSP_ALG_SM) RC(op_update_rhs(e)); // …multisplitting.c:164
  while (!done && res->outer_its < max_outer) {
    if (alg == MSP_ALG_SM) {
      int its = 0, reason = 0;
      auto t0 = clk::now();
      RC(op_inner_solve(e, &in, true, &its, &reason, nullptr));
      auto t1 = clk::now();
      res->stage_inner_s += secs(t0, t1);
      res->inner_its_total += its;
      RC(exchange_sync(e));
      RC(op_update_rhs(e));
      RC(op_resid_sumsq(e, false, 0));
      RC(allreduce_host(e, 0, 1));
      const double norm = std::sqrt(e->hsc[0]);
      res->last_norm = norm;
      if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = norm;
      if (norm <= thr_global) done = true;
      res->outer_its++;
      res->stage_outer_s += secs(t1, clk::now());
      continue;
    }
    auto t_outer0 = clk::now();
    double inner_this = 0.0;
    for (int t = 0; t < s; t++) {
      RC(op_update_rhs(e));
      int its = 0, reason = 0;
      auto t0 = clk::now();
      RC(op_inner_solve(e, &in, true, &its, &reason, nullptr));
      inner_this += secs(t0, clk::now());
      res->inner_its_total += its;
      RC(exchange_sync(e));
      RC(op_push_iterate(e, t));
    }
    if (lsqr || o->outer_type == 2) {
      // the reference's minimisers, literally: LSQR on R = A S with the raw basis (utils.c:1061-1103), or the normal
      // equations on the Gram matrix (utils.c:972-996; basis of successive corrections to keep it solvable)
      int lits = 0;
      double norm = 0.0;
      RC(op_spmm(e, alg, s, !lsqr));
      if (alg == MSP_ALG_SMSM_LOCAL) RC(op_update_rhs(e));
      double ln = 0.0;
      if (alg == MSP_ALG_SMSM_SEMI_LOCAL) { RC(op_resid_sumsq(e, false, 1)); RC(read_scalars(e, 1, 1)); ln = std::sqrt(e->hsc[1]); }
      if (lsqr) RC(op_lsqr(e, alg, s, alg == MSP_ALG_SMSM_GLOBAL, lsqr_max_it, lsqr_rtol, lsqr_abstol, alpha.data(), &norm, &lits));
      else RC(op_normal_equations(e, alg, s, alg == MSP_ALG_SMSM_GLOBAL, alpha.data(), &norm));
      res->outer_solver_its += lits;
      RC(op_apply_alpha(e, alg, s, alpha.data()));
      if (alg == MSP_ALG_SMSM_GLOBAL) {
        res->last_norm = norm; // KSPGetResidualNorm(outer_ksp) = phibar (…-global.c:343)
        if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = norm;
        if (norm <= thr_global) done = true;
      } else {
        if (alg == MSP_ALG_SMSM_LOCAL) { RC(op_resid_sumsq(e, false, 1)); RC(read_scalars(e, 1, 1)); ln = std::sqrt(e->hsc[1]); }
        if (ln <= thr_local) sticky = 1;
        e->hsc[2] = sticky; e->hsc[3] = ln * ln;
        CK(cudaMemcpyAsync(e->dsc + 2, e->hsc + 2, 16, cudaMemcpyHostToDevice, e->st));
        RC(allreduce_host(e, 2, 2));
        res->last_norm = ln;
        if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = ln;
        if ((int)std::lround(e->hsc[2]) == G) done = true;
      }
    } else if (alg == MSP_ALG_SMSM_GLOBAL) {
      RC(op_spmm(e, alg, s));
      RC(op_local_qr(e, alg, s, uaug.data()));
      // TSQR: gather every block's factor (zero-padded allreduce = allgather), identical small solve everywhere
      const int nn = (s + 1) * (s + 1);
      std::vector<double> all((size_t)G * nn, 0.0);
      if (G > 1) {
        memcpy(all.data() + (size_t)e->prob.block * nn, uaug.data(), sizeof(double) * nn);
        CK(cudaMemcpyAsync(e->dfac, all.data(), sizeof(double) * (size_t)G * nn, cudaMemcpyHostToDevice, e->st));
        RC(e->comm->allreduce_sum(e->dfac, G * nn, e->st));
        CK(cudaMemcpyAsync(all.data(), e->dfac, sizeof(double) * (size_t)G * nn, cudaMemcpyDeviceToHost, e->st));
        CK(cudaStreamSynchronize(e->st));
      } else {
        all = uaug;
      }
      double norm = 0.0;
      RC(tsqr_combine(s, G, all.data(), alpha.data(), &norm));
      RC(op_apply_alpha(e, alg, s, alpha.data()));
      res->last_norm = norm;
      if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = norm;
      if (norm <= thr_global) done = true;
    } else if (alg == MSP_ALG_SMSM_SEMI_LOCAL) {
      RC(op_spmm(e, alg, s));
      RC(op_local_qr(e, alg, s, uaug.data()));
      RC(tsqr_combine(s, 1, uaug.data(), alpha.data(), nullptr));
      RC(op_resid_sumsq(e, false, 1)); // pre-minimisation x_K against the stale rhs_K (…-semi-local.c:326)
      RC(read_scalars(e, 1, 1));
      const double ln = std::sqrt(e->hsc[1]);
      if (ln <= thr_local) sticky = 1;
      RC(op_apply_alpha(e, alg, s, alpha.data()));
      e->hsc[2] = sticky; e->hsc[3] = ln * ln;
      CK(cudaMemcpyAsync(e->dsc + 2, e->hsc + 2, 16, cudaMemcpyHostToDevice, e->st));
      RC(allreduce_host(e, 2, 2));
      res->last_norm = ln;
      if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = ln;
      if ((int)std::lround(e->hsc[2]) == G) done = true; // comm_sync_convergence_detection comm.c:235-250
    } else if (alg == MSP_ALG_SMSM_LOCAL) {
      RC(op_spmm(e, alg, s));
      RC(op_update_rhs(e));
      RC(op_local_qr(e, alg, s, uaug.data()));
      RC(tsqr_combine(s, 1, uaug.data(), alpha.data(), nullptr));
      RC(op_apply_alpha(e, alg, s, alpha.data()));
      RC(op_resid_sumsq(e, false, 1));
      RC(read_scalars(e, 1, 1));
      const double ln = std::sqrt(e->hsc[1]);
      if (ln <= thr_local) sticky = 1;
      e->hsc[2] = sticky; e->hsc[3] = ln * ln;
      CK(cudaMemcpyAsync(e->dsc + 2, e->hsc + 2, 16, cudaMemcpyHostToDevice, e->st));
      RC(allreduce_host(e, 2, 2));
      res->last_norm = ln;
      if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = ln;
      if ((int)std::lround(e->hsc[2]) == G) done = true;
    } else {
      MSP_FAIL("algorithm not handled by the synchronous driver");
    }
    res->outer_its++;
    res->stage_inner_s += inner_this;
    res->stage_outer_s += secs(t_outer0, clk::now()) - inner_this;
  }
  RC(e->comm->barrier(e->st));
  CK(cudaEventRecord(ev1, e->st));
  CK(cudaEventSynchronize(ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ev0, ev1));
  res->elapsed_s = ms * 1e-3;
  res->kernel_launches = e->launches - launches0;
  CK(cudaEventDestroy(ev0)); CK(cudaEventDestroy(ev1));
  e->prof_collect(res);
  e->prof = false;
  // closing exchange + true residual + error (comm_sync_send_and_receive_final comm.c:199, utils.c:575, :1045)
  RC(op_publish_boundary(e));
  RC(exchange_sync(e));
  RC(op_resid_sumsq(e, true, 0));
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->x, 1.0, e->ws, 2, e->dsc + 1);
  e->launches++;
  RC(allreduce_host(e, 0, 2));
  res->final_residual = std::sqrt(e->hsc[0]);
  res->error = std::sqrt(e->hsc[1]);
  return 0;
}

// gmres_solution.c:50-85
static int engine_gmres(msp_engine *e, const msp_ksp_opts *o, msp_result *res) {
  memset(res, 0, sizeof(*res));
  if (e->prob.nblocks != 1) MSP_FAIL("stand-alone GMRES runs on a single block");
  CK(cudaMemsetAsync(e->x, 0, sizeof(double) * e->ld, e->st));
  CK(cudaMemcpyAsync(e->rhs, e->b, sizeof(double) * e->ld, cudaMemcpyDeviceToDevice, e->st));
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->b, 0.0, e->ws, 2, e->dsc + 0);
  RC(read_scalars(e, 0, 1));
  res->norm0 = std::sqrt(e->hsc[0]);
  msp_ksp_opts in = *o;
  in.guess_nonzero = 0;
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  CK(cudaEventRecord(ev0, e->st));
  const int64_t l0 = e->launches;
  RC(op_inner_solve(e, &in, false, &res->gmres_its, &res->gmres_reason, &res->gmres_rnorm));
  CK(cudaEventRecord(ev1, e->st));
  CK(cudaEventSynchronize(ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ev0, ev1));
  res->elapsed_s = ms * 1e-3;
  res->kernel_launches = e->launches - l0;
  CK(cudaEventDestroy(ev0)); CK(cudaEventDestroy(ev1));
  res->outer_its = res->gmres_its;
  res->last_norm = res->gmres_rnorm;
  RC(op_resid_sumsq(e, true, 0));
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->x, 1.0, e->ws, 2, e->dsc + 1);
  RC(read_scalars(e, 0, 2));
  res->final_residual = std::sqrt(e->hsc[0]);
  res->error = std::sqrt(e->hsc[1]);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// group: all blocks in one process
// ------------------------------------------------------------------------------------------------
struct msp_group {
  int G = 0;
  std::vector<msp_engine *> eng;
  LocalShared *sh = nullptr;
};

static int group_wire(msp_group *g) {
  for (int k = 0; k < g->G; k++) {
    msp_engine *e = g->eng[k];
    if (e->own_comm && e->comm) delete e->comm;
    e->comm = new LocalComm(g->sh, k); e->own_comm = true;
    e->grp = g;
    for (int side = 0; side < 2; side++) {
      int nbk = side == 0 ? k - 1 : k + 1;
      if (nbk < 0 || nbk >= g->G) continue;
      msp_engine *p = g->eng[nbk];
      if (p->device != e->device) {
        cudaSetDevice(e->device);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, e->device, p->device);
        if (!can) MSP_FAIL("peer access between the two GPUs is not available");
        cudaError_t er = cudaDeviceEnablePeerAccess(p->device, 0);
        if (er != cudaSuccess && er != cudaErrorPeerAccessAlreadyEnabled) MSP_FAIL("cudaDeviceEnablePeerAccess failed");
        cudaGetLastError();
      }
      e->peer[side] = p->win; e->peer_ipc[side] = false;
    }
    for (int J = 0; J < g->G; J++) {
      if (J == k) continue;
      msp_engine *p = g->eng[J];
      if (p->device != e->device) {
        cudaSetDevice(e->device);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, e->device, p->device);
        if (!can) MSP_FAIL("peer access between the two GPUs is not available");
        cudaError_t er = cudaDeviceEnablePeerAccess(p->device, 0);
        if (er != cudaSuccess && er != cudaErrorPeerAccessAlreadyEnabled) MSP_FAIL("cudaDeviceEnablePeerAccess failed");
        cudaGetLastError();
      }
      e->peer_any[J] = p->win; e->peer_any_ipc[J] = false;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int msp_version(void) { return MSP_VERSION; }
const char *msp_last_error(void) { return g_err.c_str(); }
int msp_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

int64_t msp_poisson2d_nnz(int m, int n, int block, int nblocks) {
  long long nb = ((long long)m * n) / nblocks;
  return stencil_nnz_host(2, n, m, 1, nb * block, nb);
}
int64_t msp_poisson3d_nnz(int nx, int ny, int nz, int block, int nblocks) {
  long long nb = ((long long)nx * ny * nz) / nblocks;
  return stencil_nnz_host(3, nx, ny, nz, nb * block, nb);
}

static int assemble_to_host(int device, int dim, int nx, int ny, int nz, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val) {
  if (!rowptr || !colidx || !val) MSP_FAIL("null output array");
  if (nblocks < 1 || block < 0 || block >= nblocks) MSP_FAIL("bad block / nblocks");
  RC(set_device(device));
  long long ntot = (long long)nx * ny * nz;
  int nb = (int)(ntot / nblocks);
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  int *rp = nullptr, *ci = nullptr; double *va = nullptr; int64_t nnz = 0;
  int rc = assemble_strip_dev(dim, nx, ny, nz, (long long)nb * block, nb, st, &rp, &ci, &va, &nnz);
  if (!rc) {
    cudaMemcpy(rowptr, rp, sizeof(int) * ((size_t)nb + 1), cudaMemcpyDeviceToHost);
    cudaMemcpy(colidx, ci, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost);
    if (cudaMemcpy(val, va, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost) != cudaSuccess) { g_err = "copy back failed"; rc = 1; }
  }
  cudaFree(rp); cudaFree(ci); cudaFree(va);
  cudaStreamDestroy(st);
  return rc;
}
int msp_assemble_poisson2d(int device, int m, int n, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val) {
  return assemble_to_host(device, 2, n, m, 1, block, nblocks, rowptr, colidx, val);
}
int msp_assemble_poisson2d_complete(int device, int m, int n, int32_t *rowptr, int32_t *colidx, double *val) {
  if (m != n) MSP_FAIL("poisson2DMatrix_complete assumes a square mesh (utils.c:390)");
  return assemble_to_host(device, 2, n, m, 1, 0, 1, rowptr, colidx, val);
}
int msp_assemble_poisson3d(int device, int nx, int ny, int nz, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val) {
  return assemble_to_host(device, 3, nx, ny, nz, block, nblocks, rowptr, colidx, val);
}
int msp_dimension_related(int nprocs, int npb, int rank, int m, int n, int *njacobi_blocks, int *rank_jacobi_block,
                          int *proc_local_rank, int *n_mesh_points, int *jacobi_block_size) {
  if (npb <= 0 || nprocs <= 0) MSP_FAIL("bad process counts");
  *njacobi_blocks = nprocs / npb;
  *rank_jacobi_block = rank / npb;
  *proc_local_rank = rank % npb;
  *n_mesh_points = m * n;
  *jacobi_block_size = (*n_mesh_points) / (*njacobi_blocks);
  return 0;
}

int msp_create(const msp_problem *prob, int device, msp_engine **out) { return engine_create(prob, device, out); }
int msp_destroy(msp_engine *e) { return engine_free(e); }
int msp_rows(const msp_engine *e) { return e ? e->nb : -1; }
int msp_halo_size(const msp_engine *e) { return e ? e->H : -1; }

static int sub_extract(msp_engine *e, int which, int32_t *orp_h, int32_t *oci_h, double *ova_h, int64_t *nnz_out) {
  if (!e->ci) MSP_FAIL("engine was created without keep_csr");
  cudaSetDevice(e->device);
  if (which == MSP_MAT_STRIP) {
    *nnz_out = e->nnz;
    if (orp_h) {
      CK(cudaMemcpy(orp_h, e->rp, sizeof(int) * ((size_t)e->nb + 1), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(oci_h, e->ci, sizeof(int) * (size_t)e->nnz, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ova_h, e->va, sizeof(double) * (size_t)e->nnz, cudaMemcpyDeviceToHost));
    }
    return 0;
  }
  const int inside = (which == MSP_MAT_DIAG) ? 1 : 0;
  const int shift = (which == MSP_MAT_DIAG) ? e->off : 0;
  int *orp = nullptr, *oci = nullptr; double *ova = nullptr;
  CK(cudaMalloc(&orp, sizeof(int) * ((size_t)e->nb + 1)));
  CK(cudaMemsetAsync(orp, 0, sizeof(int) * ((size_t)e->nb + 1), e->st));
  k_sub_count<<<grid_for(e->nb, 16), MSPK_THREADS, 0, e->st>>>(e->nb, e->rp, e->ci, e->off, e->off + e->nb, inside, orp);
  RC(exclusive_scan_inplace(orp, e->nb + 1, e->st));
  int nnz32 = 0;
  CK(cudaMemcpy(&nnz32, orp + e->nb, sizeof(int), cudaMemcpyDeviceToHost));
  *nnz_out = nnz32;
  if (orp_h) {
    CK(cudaMalloc(&oci, sizeof(int) * (size_t)std::max(nnz32, 1)));
    CK(cudaMalloc(&ova, sizeof(double) * (size_t)std::max(nnz32, 1)));
    k_sub_fill<<<grid_for(e->nb, 16), MSPK_THREADS, 0, e->st>>>(e->nb, e->rp, e->ci, e->va, e->off, e->off + e->nb, inside, shift, orp, oci, ova);
    CK(cudaStreamSynchronize(e->st));
    CK(cudaMemcpy(orp_h, orp, sizeof(int) * ((size_t)e->nb + 1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(oci_h, oci, sizeof(int) * (size_t)nnz32, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ova_h, ova, sizeof(double) * (size_t)nnz32, cudaMemcpyDeviceToHost));
    cudaFree(oci); cudaFree(ova);
  }
  cudaFree(orp);
  return 0;
}
int64_t msp_mat_nnz(msp_engine *e, int which) {
  int64_t nnz = -1;
  if (!e || sub_extract(e, which, nullptr, nullptr, nullptr, &nnz)) return -1;
  return nnz;
}
int msp_get_csr(msp_engine *e, int which, int32_t *rowptr, int32_t *colidx, double *val) {
  if (!e || !rowptr || !colidx || !val) MSP_FAIL("null argument");
  int64_t nnz;
  return sub_extract(e, which, rowptr, colidx, val, &nnz);
}

#define VEC_SETTER(NAME, FIELD, LEN)                                                               \
  int NAME(msp_engine *e, const double *h) {                                                       \
    if (!e || !h) MSP_FAIL("null argument");                                                       \
    cudaSetDevice(e->device);                                                                      \
    CK(cudaMemcpyAsync(e->FIELD, h, sizeof(double) * (size_t)(LEN), cudaMemcpyHostToDevice, e->st)); \
    CK(cudaStreamSynchronize(e->st));                                                              \
    return 0;                                                                                      \
  }
#define VEC_GETTER(NAME, FIELD, LEN)                                                               \
  int NAME(msp_engine *e, double *h) {                                                             \
    if (!e || !h) MSP_FAIL("null argument");                                                       \
    cudaSetDevice(e->device);                                                                      \
    CK(cudaMemcpyAsync(h, e->FIELD, sizeof(double) * (size_t)(LEN), cudaMemcpyDeviceToHost, e->st)); \
    CK(cudaStreamSynchronize(e->st));                                                              \
    return 0;                                                                                      \
  }
VEC_SETTER(msp_set_b, b, e->nb)
VEC_GETTER(msp_get_b, b, e->nb)
VEC_SETTER(msp_set_x, x, e->nb)
VEC_GETTER(msp_get_x, x, e->nb)
VEC_GETTER(msp_get_rhs, rhs, e->nb)
int msp_set_halo(msp_engine *e, int side, const double *h) {
  if (!e || !h || side < 0 || side > 1) MSP_FAIL("bad argument");
  cudaSetDevice(e->device);
  CK(cudaMemcpyAsync(e->halo[side], h, sizeof(double) * (size_t)e->H, cudaMemcpyHostToDevice, e->st));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_get_halo(msp_engine *e, int side, double *h) {
  if (!e || !h || side < 0 || side > 1) MSP_FAIL("bad argument");
  cudaSetDevice(e->device);
  CK(cudaMemcpyAsync(h, e->halo[side], sizeof(double) * (size_t)e->H, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}

int msp_update_local_rhs(msp_engine *e) {
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  RC(op_update_rhs(e));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_inner_solve(msp_engine *e, const msp_ksp_opts *o, int *its, int *reason, double *rnorm) {
  if (!e || !o) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  msp_ksp_opts in = *o;
  in.initial_rtol = 1; in.guess_nonzero = 1; // utils.c:956-957
  return op_inner_solve(e, &in, false, its, reason, rnorm);
}
int msp_local_residual_norm(msp_engine *e, double *nrm) {
  if (!e || !nrm) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  RC(op_resid_sumsq(e, false, 0));
  RC(read_scalars(e, 0, 1));
  *nrm = std::sqrt(e->hsc[0]);
  return 0;
}
int msp_block_residual_norm(msp_engine *e, double *nrm) {
  if (!e || !nrm) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  RC(op_resid_sumsq(e, true, 0));
  RC(read_scalars(e, 0, 1));
  *nrm = std::sqrt(e->hsc[0]);
  return 0;
}
int msp_error_norm_sq(msp_engine *e, double *sq) {
  if (!e || !sq) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->x, 1.0, e->ws, 2, e->dsc + 1);
  RC(read_scalars(e, 1, 1));
  *sq = e->hsc[1];
  return 0;
}
int msp_push_iterate(msp_engine *e, int t) {
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  RC(op_push_iterate(e, t));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_spmm_AS(msp_engine *e, int kind) {
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  RC(op_spmm(e, kind, e->smax));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_minimize_local_qr(msp_engine *e, int kind, double *u_aug) {
  if (!e || !u_aug) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  return op_local_qr(e, kind, e->smax, u_aug);
}
int msp_apply_alpha(msp_engine *e, int kind, const double *alpha) {
  if (!e || !alpha) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  RC(op_apply_alpha(e, kind, e->smax, alpha));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_tsqr_combine(int s, int nfac, const double *u_aug_all, double *alpha, double *resnorm) {
  if (s < 1 || s > MSP_MAX_S || nfac < 1 || !u_aug_all || !alpha) MSP_FAIL("bad argument");
  return tsqr_combine(s, nfac, u_aug_all, alpha, resnorm);
}

// ---- raw kernels on host data (parity tests) ----
int msp_op_spmv(msp_engine *e, int which, const double *x, const double *halo_lo, const double *halo_hi, double *y) {
  if (!e || !x || !y) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  // staging: the first two boundary buffers of the own receive window (idle outside a solve)
  double *dlo = halo_lo ? e->win.halo(0, 0) : nullptr, *dhi = halo_hi ? e->win.halo(1, 0) : nullptr;
  CK(cudaMemcpyAsync(e->Wb[0], x, sizeof(double) * e->nb, cudaMemcpyHostToDevice, e->st));
  if (halo_lo) CK(cudaMemcpyAsync(dlo, halo_lo, sizeof(double) * e->H, cudaMemcpyHostToDevice, e->st));
  if (halo_hi) CK(cudaMemcpyAsync(dhi, halo_hi, sizeof(double) * e->H, cudaMemcpyHostToDevice, e->st));
  SpmvArgs a = spmv_args(e, e->Wb[0], e->Wb[1]);
  if (which == MSP_MAT_DIAG) launch_spmv_w<0, false, false, false>(e, a, 0, nullptr);
  else if (which == MSP_MAT_STRIP) { a.lo = dlo; a.hi = dhi; launch_spmv_w<1, false, false, false>(e, a, 0, nullptr); }
  else MSP_FAIL("which must be STRIP or DIAG");
  CK(cudaMemcpyAsync(y, e->Wb[1], sizeof(double) * e->nb, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_op_mdot(msp_engine *e, int nv, const double *V, const double *w, double *h) {
  if (!e || !V || !w || !h || nv < 1 || nv > e->nvec) MSP_FAIL("bad argument");
  cudaSetDevice(e->device);
  for (int j = 0; j < nv; j++) CK(cudaMemcpyAsync(e->V + (long long)j * e->ld, V + (size_t)j * e->nb, sizeof(double) * e->nb, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemcpyAsync(e->Wb[0], w, sizeof(double) * e->nb, cudaMemcpyHostToDevice, e->st));
  launch_mdot(e, nv, e->V, e->ld, e->Wb[0], e->dsc + 64, 1.0, -1, 0);
  CK(cudaMemcpyAsync(h, e->dsc + 64, sizeof(double) * nv, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_op_maxpy(msp_engine *e, int nv, const double *V, const double *coef, double *w, double *norm) {
  if (!e || !V || !w || !coef || nv < 1 || nv > e->nvec) MSP_FAIL("bad argument");
  cudaSetDevice(e->device);
  for (int j = 0; j < nv; j++) CK(cudaMemcpyAsync(e->V + (long long)j * e->ld, V + (size_t)j * e->nb, sizeof(double) * e->nb, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemcpyAsync(e->Wb[0], w, sizeof(double) * e->nb, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemcpyAsync(e->dsc + 64, coef, sizeof(double) * nv, cudaMemcpyHostToDevice, e->st));
  launch_maxpy<0>(e, nv, e->V, e->ld, e->dsc + 64, e->Wb[0], e->dsc + 200, -1, 0, 0, 3);
  CK(cudaMemcpyAsync(w, e->Wb[0], sizeof(double) * e->nb, cudaMemcpyDeviceToHost, e->st));
  if (norm) CK(cudaMemcpyAsync(norm, e->dsc + 200, sizeof(double), cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}

int msp_bench_kernel(msp_engine *e, int op, int nv, int iters, int flush_l2, double *ms_avg) {
  if (!e || !ms_avg || iters < 1) MSP_FAIL("bad argument");
  cudaSetDevice(e->device);
  if ((op == 1 || op == 2) && (nv < 1 || nv > e->nvec)) MSP_FAIL("nv out of range");
  if (op == 3 && (nv < 1 || nv > e->smax)) MSP_FAIL("s out of range");
  if (op == 6 && (nv < 2 || nv > 9 || nv > e->smax + 1)) MSP_FAIL("gram: 2 <= columns <= min(9, s+1)");
  double *flush = nullptr;
  const size_t flush_bytes = (size_t)256 << 20;
  if (flush_l2) CK(cudaMalloc(&flush, flush_bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // give the control block sane values for the guarded / scaled variants
  msp_ksp_opts o{30, 1000000, 1e-30, 1e-300, 1e300, 1, 1, 0, 0, 0};
  k_ctl_begin<<<1, 32, 0, e->st>>>(e->ctl, o.restart, o.max_it, 0, 1, 0, 0, o.rtol, o.abstol, o.divtol, nullptr);
  k_fill<<<1, 32, 0, e->st>>>(nv > 0 ? nv : 1, 1e-3, e->dsc + 64);
  double total = 0.0;
  for (int i = -3; i < iters; i++) {
    if (flush) CK(cudaMemsetAsync(flush, i & 0xff, flush_bytes, e->st));
    CK(cudaEventRecord(e0, e->st));
    switch (op) {
      case 0: { SpmvArgs a = spmv_args(e, e->x, e->Wb[1]); launch_spmv_w<0, false, false, false>(e, a, 0, nullptr); break; }
      case 1: launch_mdot(e, nv, e->V, e->ld, e->Wb[0], e->dsc + 64, -1.0, -1, 0); break;
      case 2: launch_maxpy<0>(e, nv, e->V, e->ld, e->dsc + 64, e->Wb[0], e->dsc + 200, -1, 0, 0, 3); break;
      case 3: RC(op_spmm(e, MSP_ALG_SMSM_GLOBAL, nv)); break;
      case 4: k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, e->Wb[0], e->Wb[1]); break;
      case 5: { SpmvArgs a = spmv_args(e, e->Wb[0], e->Wb[1]); launch_spmv_w<0, false, true, false>(e, a, 0, nullptr); break; }
      case 6: launch_gram(e, nv, e->R, e->dfac); break;
      default: MSP_FAIL("unknown op");
    }
    CK(cudaEventRecord(e1, e->st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (i >= 0) total += ms;
  }
  CK(cudaGetLastError());
  *ms_avg = total / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (flush) cudaFree(flush);
  return 0;
}

int msp_gmres_solve(msp_engine *e, const msp_ksp_opts *o, msp_result *res) {
  if (!e || !o || !res) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  return engine_gmres(e, o, res);
}

// ---- group ----
int msp_group_create(const msp_problem *prob, int nblocks, const int *devices, msp_group **out) {
  if (!prob || !out || nblocks < 1 || nblocks > MSP_MAX_BLOCKS) MSP_FAIL("bad argument");
  msp_group *g = new msp_group();
  g->G = nblocks;
  g->sh = new LocalShared(nblocks);
  for (int k = 0; k < nblocks; k++) {
    msp_problem p = *prob;
    p.block = k; p.nblocks = nblocks;
    msp_engine *e = nullptr;
    int rc = engine_create(&p, devices ? devices[k] : 0, &e);
    if (rc) { for (auto *x : g->eng) engine_free(x); delete g->sh; delete g; return rc; }
    g->eng.push_back(e);
  }
  int rc = group_wire(g);
  if (rc) { for (auto *x : g->eng) engine_free(x); delete g->sh; delete g; return rc; }
  *out = g;
  return 0;
}
int msp_group_destroy(msp_group *g) {
  if (!g) return 0;
  for (auto *e : g->eng) engine_free(e);
  delete g->sh;
  delete g;
  return 0;
}
msp_engine *msp_group_engine(msp_group *g, int k) { return (g && k >= 0 && k < g->G) ? g->eng[k] : nullptr; }

int engine_solve_async_group(msp_group *g, const msp_solve_opts *o, msp_result *res);
static int engine_solve_async(msp_engine *e, const msp_solve_opts *o, msp_result *res);

int msp_group_solve(msp_group *g, const msp_solve_opts *o, msp_result *res) {
  if (!g || !o || !res) MSP_FAIL("null argument");
  if (o->alg == MSP_ALG_GMRES) MSP_FAIL("use msp_gmres_solve for the stand-alone GMRES");
  if (o->alg >= MSP_ALG_AM) return engine_solve_async_group(g, o, res);
  std::vector<int> rcs(g->G, 0);
  std::vector<std::string> errs(g->G);
  std::vector<std::thread> th;
  for (int k = 0; k < g->G; k++)
    th.emplace_back([&, k] {
      cudaSetDevice(g->eng[k]->device);
      rcs[k] = engine_solve_sync(g->eng[k], o, &res[k]);
      if (rcs[k]) errs[k] = g_err;
    });
  for (auto &t : th) t.join();
  for (int k = 0; k < g->G; k++) if (rcs[k]) { g_err = errs[k]; return rcs[k]; }
  return 0;
}

// ---- one process per GPU ----
int msp_comm_unique_id(void *id128) {
  if (!id128) MSP_FAIL("null argument");
  if (!g_nccl.load()) MSP_FAIL("libnccl.so.2 not found");
  ncclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc) MSP_FAIL("ncclGetUniqueId failed");
  memcpy(id128, &id, 128);
  return 0;
}
int msp_comm_init(msp_engine *e, const void *id128, int rank, int nranks) {
  if (!e || !id128) MSP_FAIL("null argument");
  if (nranks != e->prob.nblocks || rank != e->prob.block) MSP_FAIL("rank / nranks must equal block / nblocks");
  if (!g_nccl.load()) MSP_FAIL("libnccl.so.2 not found");
  cudaSetDevice(e->device);
  NcclComm *c = new NcclComm();
  c->rank = rank; c->nranks = nranks;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  int rc = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (rc) { delete c; MSP_FAIL(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?")); }
  if (e->own_comm && e->comm) delete e->comm;
  e->comm = c; e->own_comm = true;
  return 0;
}
int msp_comm_export(msp_engine *e, void *handle64) {
  if (!e || !handle64) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->win.base));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
  return 0;
}
static int connect_block(msp_engine *e, int J, const void *handle64) {
  if (J < 0 || J >= e->prob.nblocks || J == e->prob.block) MSP_FAIL("bad block index");
  cudaSetDevice(e->device);
  if (!e->peer_any[J].base) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->peer_any[J] = e->win; // same geometry on every block
    e->peer_any[J].base = (double *)p;
    e->peer_any_ipc[J] = true;
  }
  if (J == e->prob.block - 1) e->peer[0] = e->peer_any[J];
  if (J == e->prob.block + 1) e->peer[1] = e->peer_any[J];
  return 0;
}
int msp_comm_connect(msp_engine *e, int side, const void *handle64) {
  if (!e || !handle64 || side < 0 || side > 1) MSP_FAIL("bad argument");
  if (!e->has_nb[side]) MSP_FAIL("no neighbour on that side");
  return connect_block(e, side == 0 ? e->prob.block - 1 : e->prob.block + 1, handle64);
}
int msp_comm_connect_block(msp_engine *e, int block, const void *handle64) {
  if (!e || !handle64) MSP_FAIL("bad argument");
  return connect_block(e, block, handle64);
}
int msp_solve(msp_engine *e, const msp_solve_opts *o, msp_result *res) {
  if (!e || !o || !res) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  if (o->alg == MSP_ALG_GMRES) return engine_gmres(e, &o->inner, res);
  for (int side = 0; side < 2; side++)
    if (e->has_nb[side] && !e->peer[side].base) MSP_FAIL("neighbour window not connected (msp_comm_connect)");
  if (o->alg >= MSP_ALG_AM) return engine_solve_async(e, o, res);
  return engine_solve_sync(e, o, res);
}

int msp_conv_detect_step(msp_engine *e, int under_threshold, int *state, int *phase_tag) {
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  k_cd_step<<<1, 32, 0, e->st>>>(e->cd, under_threshold, nullptr, 0.0);
  e->launches++;
  int hs[2];
  CK(cudaMemcpyAsync(hs, e->cd, sizeof(int) * 2, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  if (state) *state = hs[0];
  if (phase_tag) *phase_tag = hs[1];
  return 0;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// asynchronous variants (…multisplitting_prime.c:321-393, …-minimization-{global,semi-local,local}_prime.c)
// ------------------------------------------------------------------------------------------------
struct AsyncRun {
  int iters = 0;         // number_of_iterations
  int inner_outer = 0;   // number_of_inner_times_outer_iterations
  int state = 0;
  double thr_local = 0;
  double last_norm = 0;
  std::vector<double> uaug, alpha, all;
};

static int slot_of_side(const msp_engine *e, int side) { return e->has_nb[0] ? side : 0; }

// comm_async_probe_and_receive_prime comm.c:455-529 for both neighbours
static int async_probe(msp_engine *e) {
  for (int side = 0; side < 2; side++) {
    if (!e->has_nb[side]) continue;
    k_async_probe<<<1, 32, 0, e->st>>>(e->win.hdr(side), e->cd, slot_of_side(e, side), e->aint + side, e->dec + side);
    k_async_copy<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->dec + side, e->H, e->win.halo(side, 0), e->win.halo(side, 1), e->halo[side]);
    e->launches += 2;
  }
  return 0;
}
// comm_async_test_and_send_prime comm.c:531-554: P2P store of the boundary layers + header release
static int async_publish(msp_engine *e, int iter) {
  for (int side = 0; side < 2; side++) {
    if (!e->peer[side].base) continue;
    const int q = ++e->async_sent[side];
    // my first layer goes to the lower neighbour's "hi" window, my last layer to the upper neighbour's "lo" window
    double *dst = e->peer[side].halo(1 - side, q & 1);
    k_publish_boundary<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->nb, e->H, e->x, side == 0 ? dst : nullptr, side == 1 ? dst : nullptr);
    k_async_release<<<1, 32, 0, e->st>>>(e->peer[side].hdr(1 - side), e->cd, iter, q);
    e->launches += 2;
  }
  return 0;
}

static int async_begin(msp_engine *e, const msp_solve_opts *o, msp_result *res, AsyncRun *run) {
  const int G = e->prob.nblocks, s = o->s;
  memset(res, 0, sizeof(*res));
  if (o->outer_type != 0) MSP_FAIL("the LSQR and normal-equations minimisers are available for the synchronous variants only");
  if (o->alg != MSP_ALG_AM && (s < 1 || s > e->smax)) MSP_FAIL("s exceeds the engine's basis storage");
  for (int side = 0; side < 2; side++)
    if (e->has_nb[side] && !e->peer[side].base) MSP_FAIL("neighbour window not connected");
  k_cd_init<<<1, 32, 0, e->st>>>(e->cd, e->prob.block, G, e->win.mailbox(), e->peer[0].base ? e->peer[0].mailbox() : nullptr,
                                 e->peer[1].base ? e->peer[1].mailbox() : nullptr, e->win.hdr(0), e->win.hdr(1), e->aint);
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->b, 0.0, e->ws, 2, e->dsc + 0);
  e->launches += 2;
  RC(allreduce_host(e, 0, 1));
  res->norm0 = std::sqrt(e->hsc[0]);
  run->thr_local = std::max(1e-100, (o->rtol / std::sqrt((double)G)) * 1.0 * res->norm0);
  run->uaug.assign((size_t)(s + 1) * (s + 1), 0.0);
  run->alpha.assign(std::max(s, 1), 0.0);
  std::fill(e->fcache_seq.begin(), e->fcache_seq.end(), 0);
  RC(op_update_rhs(e)); // …multisplitting_prime.c:315
  return 0;
}

// one pass of the do { } while (state != FINISHED) body of this block
static int async_step(msp_engine *e, const msp_solve_opts *o, msp_result *res, AsyncRun *run) {
  const int alg = o->alg, s = o->s, G = e->prob.nblocks;
  msp_ksp_opts in = o->inner;
  in.initial_rtol = 1; in.guess_nonzero = 1;
  int its = 0, reason = 0;
  if (alg == MSP_ALG_AM) {
    RC(async_probe(e));
    RC(op_update_rhs(e));
    RC(op_inner_solve(e, &in, false, &its, &reason, nullptr));
    res->inner_its_total += its;
    RC(async_publish(e, run->iters));
    RC(op_resid_sumsq(e, false, 1));
  } else {
    for (int t = 0; t < s; t++) {
      RC(async_probe(e));
      RC(op_update_rhs(e));
      RC(op_inner_solve(e, &in, false, &its, &reason, nullptr));
      res->inner_its_total += its;
      RC(async_publish(e, run->inner_outer));
      RC(async_probe(e));
      RC(op_push_iterate(e, t));
      run->inner_outer++;
    }
    if (alg == MSP_ALG_AMAM_GLOBAL) {
      RC(op_spmm(e, alg, s));
      RC(op_local_qr(e, alg, s, run->uaug.data()));
      const int nn = (s + 1) * (s + 1), fs = e->win.fslot;
      // publish my TSQR factor to every block (replaces the R-slab Isend, comm.c:330-347), newest wins
      if (G > 1) {
        std::vector<double> slot(fs, 0.0);
        const double q = (double)(++e->factor_sent);
        slot[0] = q; slot[fs - 1] = q;
        memcpy(slot.data() + 1, run->uaug.data(), sizeof(double) * nn);
        for (int J = 0; J < G; J++)
          if (J != e->prob.block && e->peer_any[J].base)
            CK(cudaMemcpyAsync(e->peer_any[J].factor(e->prob.block), slot.data(), sizeof(double) * fs, cudaMemcpyHostToDevice, e->st));
        CK(cudaStreamSynchronize(e->st));
        // newest factors the others have published so far (comm_async_probe_and_receive_min comm.c:288-328)
        std::vector<double> mine((size_t)G * fs);
        CK(cudaMemcpyAsync(mine.data(), e->win.factor(0), sizeof(double) * (size_t)G * fs, cudaMemcpyDeviceToHost, e->st));
        CK(cudaStreamSynchronize(e->st));
        for (int J = 0; J < G; J++) {
          if (J == e->prob.block) continue;
          const double *sl = mine.data() + (size_t)J * fs;
          if (sl[0] > 0 && sl[0] == sl[fs - 1] && (int)sl[0] != e->fcache_seq[J]) {
            memcpy(e->fcache.data() + (size_t)J * fs, sl, sizeof(double) * fs);
            e->fcache_seq[J] = (int)sl[0];
          }
        }
      }
      run->all.clear();
      int nfac = 0;
      for (int J = 0; J < G; J++) {
        const double *f = nullptr;
        if (J == e->prob.block) f = run->uaug.data();
        else if (e->fcache_seq[J] > 0) f = e->fcache.data() + (size_t)J * fs + 1;
        if (f) { run->all.insert(run->all.end(), f, f + nn); nfac++; }
      }
      RC(tsqr_combine(s, nfac, run->all.data(), run->alpha.data(), nullptr));
      RC(op_apply_alpha(e, alg, s, run->alpha.data()));
      RC(op_resid_sumsq(e, true, 1)); // ||b_K - A_K,: x_min|| (…-global_prime.c:436-437)
    } else if (alg == MSP_ALG_AMAM_SEMI_LOCAL) {
      RC(op_spmm(e, alg, s));
      RC(op_local_qr(e, alg, s, run->uaug.data()));
      RC(tsqr_combine(s, 1, run->uaug.data(), run->alpha.data(), nullptr));
      RC(op_apply_alpha(e, alg, s, run->alpha.data()));
      RC(op_resid_sumsq(e, true, 1));
    } else {
      RC(op_spmm(e, alg, s));
      RC(op_update_rhs(e));
      RC(op_local_qr(e, alg, s, run->uaug.data()));
      RC(tsqr_combine(s, 1, run->uaug.data(), run->alpha.data(), nullptr));
      RC(op_apply_alpha(e, alg, s, run->alpha.data()));
      RC(op_resid_sumsq(e, false, 1));
    }
  }
  // root of the block: UnderThreshold + convergence detection state machine on the device
  k_cd_step<<<1, 32, 0, e->st>>>(e->cd, 0, e->dsc + 1, run->thr_local);
  e->launches++;
  CK(cudaMemcpyAsync(e->hsc + 48, e->cd, 8, cudaMemcpyDeviceToHost, e->st));
  RC(read_scalars(e, 1, 1));
  int hs[2];
  memcpy(hs, e->hsc + 48, 8);
  run->state = hs[0];
  run->last_norm = std::sqrt(e->hsc[1]);
  run->iters++;
  res->last_norm = run->last_norm;
  if (o->record_history && res->hist_len < 4096) res->hist[res->hist_len++] = run->last_norm;
  return 0;
}

static int async_finish(msp_engine *e, msp_result *res, AsyncRun *run) {
  res->outer_its = run->iters;
  // closing synchronous exchange + true residual + error (…multisplitting_prime.c:404-420)
  RC(op_publish_boundary(e));
  RC(exchange_sync(e));
  RC(op_resid_sumsq(e, true, 0));
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->x, 1.0, e->ws, 2, e->dsc + 1);
  e->launches++;
  RC(allreduce_host(e, 0, 2));
  res->final_residual = std::sqrt(e->hsc[0]);
  res->error = std::sqrt(e->hsc[1]);
  return 0;
}

// free-running: what each process (or each block thread) executes
static int engine_solve_async(msp_engine *e, const msp_solve_opts *o, msp_result *res) {
  AsyncRun run;
  RC(async_begin(e, o, res, &run));
  RC(e->comm->barrier(e->st));
  const int max_outer = o->max_outer > 0 ? o->max_outer : 1000000;
  const int64_t l0 = e->launches;
  auto t0 = std::chrono::steady_clock::now();
  while (run.state != 3 && run.iters < max_outer) RC(async_step(e, o, res, &run));
  CK(cudaStreamSynchronize(e->st));
  res->elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  res->kernel_launches = e->launches - l0;
  RC(e->comm->barrier(e->st));
  return async_finish(e, res, &run);
}

int engine_solve_async_group(msp_group *g, const msp_solve_opts *o, msp_result *res) {
  const int G = g->G;
  bool scheduled = false;
  for (int k = 0; k < G; k++) scheduled |= (o->period[k] > 0);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  if (!scheduled) {
    std::vector<std::thread> th;
    for (int k = 0; k < G; k++)
      th.emplace_back([&, k] {
        cudaSetDevice(g->eng[k]->device);
        rcs[k] = engine_solve_async(g->eng[k], o, &res[k]);
        if (rcs[k]) errs[k] = g_err;
      });
    for (auto &t : th) t.join();
    for (int k = 0; k < G; k++) if (rcs[k]) { g_err = errs[k]; return rcs[k]; }
    return 0;
  }
  // deterministic schedule (tests): block K runs one step at tick t iff t % period[K] == 0, blocks in index order;
  // collective phases (norm0, closing exchange) still need one thread per block
  std::vector<AsyncRun> runs(G);
  auto par_all = [&](auto fn) {
    std::vector<std::thread> th;
    for (int k = 0; k < G; k++)
      th.emplace_back([&, k] { cudaSetDevice(g->eng[k]->device); rcs[k] = fn(k); if (rcs[k]) errs[k] = g_err; });
    for (auto &t : th) t.join();
    for (int k = 0; k < G; k++) if (rcs[k]) { g_err = errs[k]; return rcs[k]; }
    return 0;
  };
  RC(par_all([&](int k) { int rc = async_begin(g->eng[k], o, &res[k], &runs[k]); if (!rc) rc = g->eng[k]->comm->barrier(g->eng[k]->st); return rc; }));
  const long long max_ticks = (long long)(o->max_outer > 0 ? o->max_outer : 1000000) * 64;
  int nfin = 0;
  for (long long tick = 0; nfin < G && tick < max_ticks; tick++) {
    for (int k = 0; k < G; k++) {
      const int per = o->period[k] > 0 ? o->period[k] : 1;
      if (tick % per || runs[k].state == 3) continue;
      cudaSetDevice(g->eng[k]->device);
      RC(async_step(g->eng[k], o, &res[k], &runs[k]));
      CK(cudaStreamSynchronize(g->eng[k]->st));
      if (runs[k].state == 3) nfin++;
    }
  }
  if (nfin < G) MSP_FAIL("asynchronous schedule cap reached before every block finished");
  return par_all([&](int k) { return async_finish(g->eng[k], &res[k], &runs[k]); });
}

// kernels.cuh — hand-written sm_100a kernels of the multisplitting solve path.
//
// All arithmetic is fp64, indices int32 (PETSc build of the reference: --with-precision=double,
// 32-bit indices, config/petsc/arch-linux-mpich-g5k-opt.py:46).  Every kernel is HBM-bandwidth
// bound (<= 0.25 flop/B); the design rules are therefore: 128-bit coalesced loads,
// many independent loads in flight per thread, grids sized in multiples of the SM count, one pass
// over each operand, reductions finished in-kernel by the last block (fixed order => deterministic).
//
// The file is compiled with --fmad=false; fused multiply-adds are written explicitly so that the
// summation order and rounding are the same as the CPU oracle's (oracle/msplit_oracle.c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MSPK_MAXK 64     // max restart
#define MSPK_THREADS 256
#define MSPK_MAX_PART 2048
#ifndef MSPK_CDIA_MINB5
#define MSPK_CDIA_MINB5 4   // resident blocks per SM the hot coded-DIA SpMV is compiled for (64 registers; measured: 5 blocks = 48 registers spills and loses 20 %, uncapped 98+ registers loses 25 %)
#define MSPK_CDIA_MINB7 4
#endif

// grid of a grid-stride kernel: enough blocks for the work, at most `per_sm` per SM and fewer than MSPK_MAX_PART (one
// reduction partial per block).  Host and device share it: the persistent restart-cycle kernel (cycle_coop.cuh) forms its
// reduction partials over the SAME virtual grids as the one-kernel-per-phase path, which is what makes the two bit-identical.
__host__ __device__ inline int msp_grid_for(long long work_items, int per_sm, int num_sms) {
  long long need = (work_items + MSPK_THREADS - 1) / MSPK_THREADS;
  long long cap = (long long)num_sms * per_sm;
  if (cap > MSPK_MAX_PART - 1) cap = (MSPK_MAX_PART - 1) / num_sms * num_sms;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
// launch geometry of VecMDot over nv vectors: y-groups of <= gmax vectors, kernel variant and x-grid per group size
struct MdotGeom { int per_group, ngroups, gx, variant; }; // variant: 24 / 16 / 8 / 4 / 2 = NVMAX of k_mdot
__host__ __device__ inline MdotGeom mdot_geometry(int nb, int nv, int gmax, int num_sms) {
  MdotGeom g;
  g.ngroups = (nv + gmax - 1) / gmax;
  g.per_group = (nv + g.ngroups - 1) / g.ngroups;
  g.ngroups = (nv + g.per_group - 1) / g.per_group;
  int per_sm = 8 / g.ngroups;
  if (per_sm < 1) per_sm = 1;
  if (g.per_group > 16) { g.variant = 24; g.gx = msp_grid_for((long long)nb / 2, 1, num_sms); }
  else if (g.per_group > 8) { g.variant = 16; g.gx = msp_grid_for((long long)nb / 2, per_sm < 2 ? per_sm : 2, num_sms); }
  else if (g.per_group <= 2) { g.variant = 2; g.gx = msp_grid_for((long long)nb / 16, per_sm, num_sms); }
  else if (g.per_group <= 4) { g.variant = 4; g.gx = msp_grid_for((long long)nb / 8, per_sm, num_sms); }
  else { g.variant = 8; g.gx = msp_grid_for((long long)nb / 4, per_sm, num_sms); }
  return g;
}

// ------------------------------------------------------------------------------------------------
// device-resident GMRES control block (KSP_GMRES of PETSc: HH, cc/ss rotations, GRS, convergence
// context).  Kernels read `active`/`it` to turn into no-ops once the cycle has ended, so that a whole
// restart cycle is enqueued without any host round trip ("conv_detection logic as device-side flags").
// ------------------------------------------------------------------------------------------------
struct GmresCtl {
  // options
  int restart, max_it, min_it, initial_rtol, guess_zero, cgs_refine;
  double rtol, abstol, divtol, bnorm;
  // state
  int its;         // ksp->its
  int it;          // index inside the current cycle
  int reason;      // KSPConvergedReason
  int active;      // 1 while the current cycle iterates
  int hapend;
  int refine;      // second CGS pass requested (REFINE_IFNEEDED)
  int first_cycle; // ksp->rnorm == -1 marker
  int pad0;
  long long its_total; // sum of KSPGetIterationNumber over the inner solves since the driver cleared it (read once per solve)
  double res;      // current recurrence residual
  double ksp_rnorm;
  double gm_rnorm0; // residual at the start of the cycle
  double rnorm0, ttol; // KSPConvergedDefault context
  double inv;      // 1/norm of the newest basis vector
  double inv_arr[MSPK_MAXK + 2]; // 1/||w_j||: the basis is stored un-normalised, v_j = vtilde_j * inv_arr[j] on the fly
  double tt;       // last norm
  double grs[MSPK_MAXK + 2], cc[MSPK_MAXK + 2], ss[MSPK_MAXK + 2];
  double lhh[MSPK_MAXK + 2];  // MDot result of the current step (pass 0 / pass 1)
  double nrs[MSPK_MAXK + 2];
  double hh[(MSPK_MAXK + 2) * (MSPK_MAXK + 1)]; // column-major, ld = MSPK_MAXK + 2
};

#define MSPK_NCOUNTER 128      // reduction tickets: 0..7 norm slots, 8..8+63 MDot groups (<= 64 groups), 120 Gram kernel
#define MSPK_GRAM_COUNTER 120
struct ReduceWs {       // workspace of the last-block-done reductions
  double *partial;      // [MSPK_MAX_PART * 192]
  unsigned int *counter; // [MSPK_NCOUNTER]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  Fixed order => deterministic.
__device__ __forceinline__ double block_sum(double v, double *sm /* >= 32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = (lane < (blockDim.x >> 5)) ? sm[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// 128-bit streaming loads (read-once data: bypass L1 allocation)
__device__ __forceinline__ double2 ld_stream2(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_stream_i2(const int *p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------------
// KSPConvergedDefault (iterativ.c) on the device
// ------------------------------------------------------------------------------------------------
__device__ inline int ctl_converged(GmresCtl *c, int n, double rnorm) {
  if (n == 0) {
    if (!c->guess_zero && !c->initial_rtol) {
      double snorm = c->bnorm;
      if (snorm == 0.0) snorm = rnorm;
      c->rnorm0 = snorm;
    } else {
      c->rnorm0 = rnorm;
    }
    c->ttol = fmax(c->rtol * c->rnorm0, c->abstol);
  }
  if (n <= c->min_it) return 0;
  if (isnan(rnorm) || isinf(rnorm)) return -9;
  if (rnorm <= c->ttol) return (rnorm < c->abstol) ? 3 : 2;
  if (rnorm >= c->divtol * c->rnorm0) return -4;
  return 0;
}

// start of a restart cycle (KSPGMRESCycle prologue): res = ||r||, breakdown sanity
// check, GRS(0) = res, convergence test with (ksp->its, res)
__device__ inline void ctl_cycle_begin(GmresCtl *c, double res) {
  c->it = 0;
  c->hapend = 0;
  c->refine = 0;
  // a cycle enqueued ahead of time (deferred status reads) after the solve has already ended: nothing to do.  KSPSolve_GMRES
  // leaves its loop on any reason, and on itcount >= max_it; the host-driven path never gets here in that state.
  if (!c->first_cycle && (c->reason != 0 || c->its >= c->max_it)) { c->active = 0; return; }
  c->tt = res;
  c->inv = (res > 0.0) ? 1.0 / res : 0.0;
  c->inv_arr[0] = c->inv;
  if (isnan(res) || isinf(res)) { c->reason = -9; c->active = 0; return; }
  if (!c->first_cycle && c->ksp_rnorm > 0.0 && fabs(res - c->ksp_rnorm) > 0.1 * c->gm_rnorm0) {
    c->reason = -5; c->active = 0; return;
  }
  c->first_cycle = 0;
  c->grs[0] = res;
  c->gm_rnorm0 = res;
  c->ksp_rnorm = res;
  c->res = res;
  if (res == 0.0) { c->reason = 3; c->active = 0; return; }
  c->reason = ctl_converged(c, c->its, res);
  c->active = (c->reason == 0 && c->its < c->max_it) ? 1 : 0;
}

// after orthogonalisation of step `it`: tt = ||v_{it+1}||, Hessenberg/Givens update
// (KSPGMRESUpdateHessenberg), its++, convergence test, happy breakdown
__device__ inline void ctl_step_end(GmresCtl *c, double tt) {
  const int it = c->it;
  const int ld = MSPK_MAXK + 2;
  double *hh = &c->hh[(size_t)it * ld];
  c->tt = tt;
  c->inv = (tt > 0.0) ? 1.0 / tt : 0.0;
  c->inv_arr[it + 1] = c->inv;
  if (isnan(tt) || isinf(tt)) { c->reason = -9; c->active = 0; return; }
  hh[it + 1] = tt;
  double hapbnd = fabs(tt / c->grs[it]);
  if (hapbnd > 1e-30) hapbnd = 1e-30;
  if (tt < hapbnd) c->hapend = 1;
  for (int j = 1; j <= it; j++) {
    double t = hh[j - 1];
    hh[j - 1] = c->cc[j - 1] * t + c->ss[j - 1] * hh[j];
    hh[j] = c->cc[j - 1] * hh[j] - c->ss[j - 1] * t;
  }
  double res = c->res;
  if (!c->hapend) {
    double t2 = sqrt(hh[it] * hh[it] + hh[it + 1] * hh[it + 1]);
    if (t2 == 0.0) {
      c->reason = -2;
    } else {
      c->cc[it] = hh[it] / t2;
      c->ss[it] = hh[it + 1] / t2;
      c->grs[it + 1] = -(c->ss[it] * c->grs[it]);
      c->grs[it] = c->cc[it] * c->grs[it];
      hh[it] = c->cc[it] * hh[it] + c->ss[it] * hh[it + 1];
      res = fabs(c->grs[it + 1]);
    }
  } else {
    res = 0.0;
  }
  c->it = it + 1;
  c->its += 1;
  c->ksp_rnorm = res;
  c->res = res;
  if (!c->reason) c->reason = ctl_converged(c, c->its, res);
  if (c->hapend && !c->reason) c->reason = -5;
  c->active = (c->reason == 0 && c->it < c->restart && c->its < c->max_it) ? 1 : 0;
}

// end of one classical Gram-Schmidt pass of step `it` (borthog2.c): hh[j] -= lhh[j] (lhh holds -<w, v_j>), decide on the
// second pass (REFINE_IFNEEDED / REFINE_ALWAYS), otherwise close the step with tt = ||w||
__device__ inline void ctl_cgs_pass_end(GmresCtl *c, double tt, int pass) {
  const int it = c->it;
  double *hh = &c->hh[(size_t)it * (MSPK_MAXK + 2)];
  double hn = 0.0;
  for (int j = 0; j <= it; j++) {
    if (pass == 0) hh[j] = 0.0;
    hh[j] -= c->lhh[j];
    hn = fma(c->lhh[j], c->lhh[j], hn);
  }
  bool more = false;
  if (pass == 0) {
    if (c->cgs_refine == 2) more = true;
    else if (c->cgs_refine == 1) more = (tt < sqrt(hn));
  }
  c->refine = more ? 1 : 0;
  if (!more) ctl_step_end(c, tt);
}

// ------------------------------------------------------------------------------------------------
// K12  CSR assembly of the Poisson strips (poisson2DMatrix utils.c:247-293, poisson3DMatrix :30-121)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int stencil_row(int dim, int nx, int ny, int nz, long long row, int *cols, double *vals) {
  // dim 2: nx = n_grid_columns (fastest), ny = n_grid_lines.  dim 3: nx lines (fastest), ny columns, nz depth.
  int c = 0;
  if (dim == 2) {
    long long i = row / nx, j = row - i * nx;
    if (i > 0) { cols[c] = (int)(row - nx); vals[c++] = -1.0; }
    if (j > 0) { cols[c] = (int)(row - 1); vals[c++] = -1.0; }
    cols[c] = (int)row; vals[c++] = 4.0;
    if (j < nx - 1) { cols[c] = (int)(row + 1); vals[c++] = -1.0; }
    if (i < ny - 1) { cols[c] = (int)(row + nx); vals[c++] = -1.0; }
  } else {
    long long pl = (long long)nx * ny;
    long long k = row / pl, rem = row - k * pl, j = rem / nx, i = rem - j * nx;
    if (k > 0) { cols[c] = (int)(row - pl); vals[c++] = -1.0; }
    if (j > 0) { cols[c] = (int)(row - nx); vals[c++] = -1.0; }
    if (i > 0) { cols[c] = (int)(row - 1); vals[c++] = -1.0; }
    cols[c] = (int)row; vals[c++] = 6.0;
    if (i < nx - 1) { cols[c] = (int)(row + 1); vals[c++] = -1.0; }
    if (j < ny - 1) { cols[c] = (int)(row + nx); vals[c++] = -1.0; }
    if (k < nz - 1) { cols[c] = (int)(row + pl); vals[c++] = -1.0; }
  }
  return c;
}

__global__ void k_stencil_count(int dim, int nx, int ny, int nz, long long row0, int nb, int *cnt) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    int cols[7]; double vals[7];
    cnt[r] = stencil_row(dim, nx, ny, nz, row0 + r, cols, vals);
  }
}

__global__ void k_stencil_fill(int dim, int nx, int ny, int nz, long long row0, int nb, const int *__restrict__ rowptr,
                               int *__restrict__ colidx, double *__restrict__ val) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    int cols[7]; double vals[7];
    int c = stencil_row(dim, nx, ny, nz, row0 + r, cols, vals);
    int p = rowptr[r];
    for (int k = 0; k < c; k++) { colidx[p + k] = cols[k]; val[p + k] = vals[k]; }
  }
}

// divideSubDomainIntoBlockMatrices utils.c:450-478: count / fill the entries of a column window
__global__ void k_sub_count(int nb, const int *__restrict__ rowptr, const int *__restrict__ colidx, int lo, int hi, int inside,
                            int *cnt) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    int c = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      int in = (colidx[k] >= lo && colidx[k] < hi);
      c += (in == inside);
    }
    cnt[r] = c;
  }
}
__global__ void k_sub_fill(int nb, const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val,
                           int lo, int hi, int inside, int shift, const int *__restrict__ orp, int *__restrict__ oci,
                           double *__restrict__ ova) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    int p = orp[r];
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      int in = (colidx[k] >= lo && colidx[k] < hi);
      if (in == inside) { oci[p] = colidx[k] - shift; ova[p++] = val[k]; }
    }
  }
}

// strip CSR (global columns) -> slot-major ELL with block-local signed columns: c = global - off.
// c in [0, nb) own rows, c < 0 lower neighbour's boundary, c >= nb upper neighbour's boundary.
// Padding: value 0, column = own row (fma(0, x, s) == s exactly => bit-identical to the CSR chain).
__global__ void k_csr_to_ell(int nb, int W, long long ld, int off, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                             const double *__restrict__ val, int *__restrict__ ecol, double *__restrict__ eval) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < ld; r += (long long)gridDim.x * blockDim.x) {
    int p = 0, q = 0;
    if (r < nb) { p = rowptr[r]; q = rowptr[r + 1]; }
    for (int k = 0; k < W; k++) {
      if (p + k < q) { ecol[k * ld + r] = colidx[p + k] - off; eval[k * ld + r] = val[p + k]; }
      else { ecol[k * ld + r] = (r < nb) ? (int)r : 0; eval[k * ld + r] = 0.0; }
    }
  }
}

// rows that own at least one off-block entry (the boundary rows of A_KJ)
__global__ void k_mark_boundary(int nb, int W, long long ld, const int *__restrict__ ecol, int *flag) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    int f = 0;
    for (int k = 0; k < W; k++) { int c = ecol[k * ld + r]; f |= (c < 0 || c >= nb); }
    flag[r] = f;
  }
}

// ------------------------------------------------------------------------------------------------
// Diagonal (DIA) view of a strip whose entries sit on at most 8 distinct diagonals (col - row), which is what the
// 5-/7-point stencil strips are.  Found by inspecting the assembled CSR, not assumed: k_mark_offsets sets one bit per
// occurring offset, the host sorts them ascending (= sorted-column order of every row), k_csr_to_dia scatters the
// values into ND slot-major arrays (missing entries 0).  The hot SpMV then needs NO index loads: 8 ND + 16 bytes per
// row instead of 12 W + 16.  fma(0, x, s) == s keeps the result bit-identical to the CSR chain.
// ------------------------------------------------------------------------------------------------
__global__ void k_mark_offsets(int nb, int off, int H, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                               unsigned int *bitmap /* 2H+1 bits */, int *out_of_range) {
  // every thread remembers the offsets it has already published (a stencil strip has 5 or 7 in total), so the
  // bitmap sees a handful of atomics per thread instead of one per non-zero
  long long seen[8];
  int nseen = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x)
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      const long long d = (long long)colidx[k] - (off + r);
      bool known = false;
      for (int j = 0; j < nseen; j++) known |= (seen[j] == d);
      if (known) continue;
      if (nseen < 8) seen[nseen++] = d;
      if (d < -H || d > H) { *out_of_range = 1; continue; }
      const unsigned bit = (unsigned)(d + H);
      atomicOr(bitmap + (bit >> 5), 1u << (bit & 31));
    }
}
struct DiaOffsets { int nd; int off[8]; };
__global__ void k_csr_to_dia(int nb, long long ld, int off, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                             const double *__restrict__ val, DiaOffsets d, double *__restrict__ dval /* zero-initialised */) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x)
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      const int dd = colidx[k] - (off + (int)r);
      for (int j = 0; j < d.nd; j++)
        if (d.off[j] == dd) { dval[j * ld + r] = val[k]; break; }
    }
}


// ------------------------------------------------------------------------------------------------
// Coded DIA view.  In a constant-coefficient stencil strip every diagonal holds ONE value wherever it is present
// (-1, or 4 / 6 on the main diagonal) and nothing at the domain faces.  k_dia_probe_const checks exactly that on the
// DIA arrays (bit patterns, per diagonal: all non-zero entries identical); if it holds, k_dia_to_mask packs the strip
// into one presence byte per row and the hot SpMV streams 1 + 16 bytes per row instead of 8 ND + 16.  The value used
// for entry (r, k) is `bit ? const[k] : 0.0`, i.e. exactly the double the DIA array held, so the fma chain and its
// result are unchanged bit for bit.  Any other matrix keeps the plain DIA (or ELL) view.
// ------------------------------------------------------------------------------------------------
#define MSPK_DIA_UNSET 0x7ff8dead0badc0deULL   // a NaN payload no assembled matrix carries
__global__ void k_dia_probe_const(int nb, long long ld, int nd, const double *__restrict__ dval,
                                  unsigned long long *slot /* [8], MSPK_DIA_UNSET */, int *nonconst) {
  unsigned long long mine[8];
#pragma unroll
  for (int k = 0; k < 8; k++) mine[k] = MSPK_DIA_UNSET;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x)
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k >= nd) break;
      const unsigned long long v = (unsigned long long)__double_as_longlong(dval[k * ld + r]);
      if (v == 0ULL || v == mine[k]) continue;
      if (mine[k] == MSPK_DIA_UNSET) {   // first non-zero this thread sees on diagonal k: agree on it grid-wide
        const unsigned long long old = atomicCAS(slot + k, MSPK_DIA_UNSET, v);
        mine[k] = (old == MSPK_DIA_UNSET) ? v : old;
        if (mine[k] == v) continue;
      }
      *nonconst = 1;
    }
}
__global__ void k_dia_to_mask(int nb, long long ld, int nd, const double *__restrict__ dval,
                              const unsigned long long *__restrict__ slot, unsigned char *__restrict__ mask) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    unsigned m = 0;
    for (int k = 0; k < nd; k++)
      if ((unsigned long long)__double_as_longlong(dval[k * ld + r]) == slot[k] && slot[k] != 0ULL) m |= 1u << k;
    mask[r] = (unsigned char)m;
  }
}

// ------------------------------------------------------------------------------------------------
// K1/K2  ELL SpMV  y = A x | y = b - A x,  optional deferred normalisation of the input
// (v = w * inv written to vout, K5) and optional fused ||y||^2 (cycle prologue).
// Two rows per thread: 128-bit loads of the values, 64-bit of the indices, all coalesced
// (slot-major).  x is gathered through L1/L2: each x entry is read from HBM once.
//   MODE 0: A_KK only (halo columns contribute 0)      MODE 1: strip, halos from lo/hi
// ------------------------------------------------------------------------------------------------
struct SpmvArgs {
  int nb, W, H;
  long long ld;
  const int *ecol;
  const double *eval;
  const double *dval;   // DIA values (slot-major), or null
  DiaOffsets dia;
  const unsigned char *dmask; // coded DIA: one presence byte per row (bit k = diagonal k holds its constant), or null
  int stencil;          // coded DIA: offsets are (.., -D, -1, 0, +1, +D, ..) with every far offset a multiple of 4
  double dconst[8];     // coded DIA: the constant of every diagonal
  const double *x;      // input (own rows)
  const double *lo, *hi; // neighbour boundaries (MODE 1), may be null
  const double *b;      // RESID: y = b - A x
  double *y;
  const GmresCtl *ctl;  // SCALE: input is the un-normalised basis vector, inv = ctl->inv_arr[max(guard_it, 0)] ; guards
  int guard_it;         // run only if ctl->active && ctl->it == guard_it  (-1: always)
};

// COH: the vector being read was written earlier in the SAME kernel by other thread blocks (persistent restart-cycle
// kernel): read it through L2 (ld.cg) instead of the non-coherent path; the value loaded is the same double.
template <int MODE, bool COH = false>
__device__ __forceinline__ double gather_x(const SpmvArgs &a, int c, double inv, bool scale) {
  if ((unsigned)c < (unsigned)a.nb) {
    double v = COH ? __ldcg(a.x + c) : __ldg(a.x + c);
    return scale ? v * inv : v;
  }
  if (MODE == 0) return 0.0;
  if (c < 0) return a.lo ? __ldg(a.lo + (c + a.H)) : 0.0;
  return a.hi ? __ldg(a.hi + (c - a.nb)) : 0.0;
}

// fused ||y||^2 of the cycle-prologue SpMV: block partials, the last block to arrive sums them in index order
// (deterministic) and opens the restart cycle on the device
__device__ __forceinline__ void spmv_norm_epilogue(double nrm, ReduceWs ws, int ws_slot, GmresCtl *ctl_rw) {
  __shared__ double sm[32];
  __shared__ bool last;
  double bs = block_sum(nrm, sm);
  if (threadIdx.x == 0) {
    ws.partial[ws_slot * MSPK_MAX_PART + blockIdx.x] = bs;
    __threadfence();
    unsigned t = atomicAdd(ws.counter + ws_slot, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    double v = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(ws.partial + ws_slot * MSPK_MAX_PART + i);
    double tot = block_sum(v, sm);
    if (threadIdx.x == 0) {
      ws.counter[ws_slot] = 0;
      ws.partial[ws_slot * MSPK_MAX_PART + MSPK_MAX_PART - 1] = tot; // also left readable for the host
      if (ctl_rw) ctl_cycle_begin(ctl_rw, sqrt(tot));
    }
  }
}

template <int W_T, int MODE, bool RESID, bool SCALE, bool NORM>
__global__ void __launch_bounds__(MSPK_THREADS) k_spmv_ell(SpmvArgs a, ReduceWs ws, int ws_slot, GmresCtl *ctl_rw) {
  if (a.guard_it >= 0) {
    if (!a.ctl->active || a.ctl->it != a.guard_it) return;
  }
  const int W = (W_T > 0) ? W_T : a.W;
  const double inv = SCALE ? a.ctl->inv_arr[a.guard_it > 0 ? a.guard_it : 0] : 1.0;
  double nrm = 0.0;
  const long long npairs = (a.nb + 1) >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    const long long r = p * 2;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < ((W_T > 0) ? W_T : 8); k++) {
      if (W_T == 0 && k >= W) break;
      const double2 v = ld_stream2(a.eval + k * a.ld + r);
      const int2 c = ld_stream_i2(a.ecol + k * a.ld + r);
      const double x0 = gather_x<MODE>(a, c.x, inv, SCALE);
      const double x1 = gather_x<MODE>(a, c.y, inv, SCALE);
      s0 = fma(v.x, x0, s0);
      s1 = fma(v.y, x1, s1);
    }
    if (RESID) {
      s0 = a.b[r] - s0;
      if (r + 1 < a.nb) s1 = a.b[r + 1] - s1;
    }
    if (r + 1 < a.nb) {
      *reinterpret_cast<double2 *>(a.y + r) = make_double2(s0, s1);
      if (NORM) nrm = fma(s0, s0, fma(s1, s1, nrm));
    } else {
      a.y[r] = s0;
      if (NORM) nrm = fma(s0, s0, nrm);
    }
  }
  if (NORM) spmv_norm_epilogue(nrm, ws, ws_slot, ctl_rw);
}

// DIA SpMV: same contract, template flags and epilogue as k_spmv_ell; columns are r + off[k], no index stream.
template <int ND_T, int MODE, bool RESID, bool SCALE, bool NORM>
__global__ void __launch_bounds__(MSPK_THREADS) k_spmv_dia(SpmvArgs a, ReduceWs ws, int ws_slot, GmresCtl *ctl_rw) {
  if (a.guard_it >= 0) {
    if (!a.ctl->active || a.ctl->it != a.guard_it) return;
  }
  const int ND = (ND_T > 0) ? ND_T : a.dia.nd;
  const double inv = SCALE ? a.ctl->inv_arr[a.guard_it > 0 ? a.guard_it : 0] : 1.0;
  double nrm = 0.0;
  const long long npairs = (a.nb + 1) >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    const long long r = p * 2;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < ((ND_T > 0) ? ND_T : 8); k++) {
      if (ND_T == 0 && k >= ND) break;
      const double2 v = ld_stream2(a.dval + k * a.ld + r);
      const int c = (int)r + a.dia.off[k];
      const double x0 = gather_x<MODE>(a, c, inv, SCALE);
      // the second row of the last pair does not exist when nb is odd: no gather (its column may lie past the halo)
      const double x1 = (r + 1 < a.nb) ? gather_x<MODE>(a, c + 1, inv, SCALE) : 0.0;
      s0 = fma(v.x, x0, s0);
      s1 = fma(v.y, x1, s1);
    }
    if (RESID) {
      s0 = a.b[r] - s0;
      if (r + 1 < a.nb) s1 = a.b[r + 1] - s1;
    }
    if (r + 1 < a.nb) {
      *reinterpret_cast<double2 *>(a.y + r) = make_double2(s0, s1);
      if (NORM) nrm = fma(s0, s0, fma(s1, s1, nrm));
    } else {
      a.y[r] = s0;
      if (NORM) nrm = fma(s0, s0, nrm);
    }
  }
  if (NORM) spmv_norm_epilogue(nrm, ws, ws_slot, ctl_rw);
}

// Coded-DIA SpMV, general form: same contract, template flags and epilogue as k_spmv_ell / k_spmv_dia.  Four rows per
// thread: one 32-bit load brings their presence bytes, x is read with the widest aligned load the diagonal's offset
// allows (256-bit when off % 4 == 0, 2 x 128-bit when even, 64/128/64 when odd; all L1-cached), y leaves with one
// 256-bit store.  Algorithmic traffic 17 bytes per row.  L1-wavefront-bound (ncu: 88 %, 0.214 ms at 67 M rows) because
// of the odd offsets and of its per-diagonal branches; stencil-shaped strips take k_spmv_cdia_stencil below.
__device__ __forceinline__ void ld4_cached(const double *p, double (&v)[4]) {
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld4_l2(const double *p, double (&v)[4]) { // coherent at L2 (see gather_x, COH)
  asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st4(double *p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
// x[c .. c+3] for the rows r .. r+3 of one thread (nvalid of them exist); entries of missing rows are 0
template <int MODE>
__device__ __forceinline__ void gather_x4(const SpmvArgs &a, int c, int nvalid, bool x_al32, double inv, bool scale, double (&xv)[4]) {
  if (nvalid == 4 && c >= 0 && c + 3 < a.nb) { // all four columns are own rows
    if ((c & 3) == 0 && x_al32) {
      ld4_cached(a.x + c, xv);
    } else if ((c & 1) == 0 && x_al32) {
      const double2 p = __ldg(reinterpret_cast<const double2 *>(a.x + c));
      const double2 q = __ldg(reinterpret_cast<const double2 *>(a.x + c + 2));
      xv[0] = p.x; xv[1] = p.y; xv[2] = q.x; xv[3] = q.y;
    } else if (x_al32) {
      xv[0] = __ldg(a.x + c);
      const double2 p = __ldg(reinterpret_cast<const double2 *>(a.x + c + 1));
      xv[1] = p.x; xv[2] = p.y;
      xv[3] = __ldg(a.x + c + 3);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) xv[i] = __ldg(a.x + c + i);
    }
    if (scale) {
#pragma unroll
      for (int i = 0; i < 4; i++) xv[i] = xv[i] * inv;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) xv[i] = (i < nvalid) ? gather_x<MODE>(a, c + i, inv, scale) : 0.0;
  }
}
template <int ND_T, int MODE, bool RESID, bool SCALE, bool NORM>
__global__ void __launch_bounds__(MSPK_THREADS) k_spmv_cdia(SpmvArgs a, ReduceWs ws, int ws_slot, GmresCtl *ctl_rw) {
  if (a.guard_it >= 0) {
    if (!a.ctl->active || a.ctl->it != a.guard_it) return;
  }
  const int ND = (ND_T > 0) ? ND_T : a.dia.nd;
  const double inv = SCALE ? a.ctl->inv_arr[a.guard_it > 0 ? a.guard_it : 0] : 1.0;
  const bool x_al32 = ((reinterpret_cast<uintptr_t>(a.x) & 31) == 0);
  const bool y_al32 = ((reinterpret_cast<uintptr_t>(a.y) & 31) == 0);
  double nrm = 0.0;
  const long long nquads = ((long long)a.nb + 3) >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquads; q += (long long)gridDim.x * blockDim.x) {
    const long long r = q * 4;
    const unsigned m = __ldg(reinterpret_cast<const unsigned *>(a.dmask) + q); // rows >= nb: byte 0
    const int nvalid = (a.nb - r < 4) ? (int)(a.nb - r) : 4;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < ((ND_T > 0) ? ND_T : 8); k++) {
      if (ND_T == 0 && k >= ND) break;
      const double ck = a.dconst[k];
      double xv[4];
      gather_x4<MODE>(a, (int)r + a.dia.off[k], nvalid, x_al32, inv, SCALE, xv);
#pragma unroll
      for (int i = 0; i < 4; i++) s[i] = fma(((m >> (8 * i + k)) & 1u) ? ck : 0.0, xv[i], s[i]);
    }
    if (nvalid == 4) {
      if (RESID) {
        double bv[4];
        if ((reinterpret_cast<uintptr_t>(a.b) & 31) == 0) ld4_cached(a.b + r, bv);
        else {
#pragma unroll
          for (int i = 0; i < 4; i++) bv[i] = a.b[r + i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) s[i] = bv[i] - s[i];
      }
      if (y_al32) st4(a.y + r, s);
      else {
#pragma unroll
        for (int i = 0; i < 4; i++) a.y[r + i] = s[i];
      }
      // same association as the two-rows-per-thread kernels: (s0, s1) then (s2, s3) of consecutive row pairs
      if (NORM) { nrm = fma(s[0], s[0], fma(s[1], s[1], nrm)); nrm = fma(s[2], s[2], fma(s[3], s[3], nrm)); }
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (i < nvalid) {
          if (RESID) s[i] = a.b[r + i] - s[i];
          a.y[r + i] = s[i];
          if (NORM) nrm = fma(s[i], s[i], nrm);
        }
    }
  }
  if (NORM) spmv_norm_epilogue(nrm, ws, ws_slot, ctl_rw);
}

// Coded-DIA SpMV for stencil-shaped strips: ND = 5 / 7 diagonals at offsets (.., -D, -1, 0, +1, +D, ..) with every far
// offset a multiple of 4 (checked on the host: e->dia_stencil).  Four rows per thread.  All loads of a trip leave back
// to back before anything depends on them: the thread's own x quad and one aligned quad per far diagonal as 256-bit
// loads, the presence bytes, and on the warp's edge lanes x[r-1] / x[r+4]; the other lanes take those two from their
// neighbours by shuffle.  Then the fma chains run in diagonal order — the same doubles in the same order as the CSR
// chain.  Per 128 rows: (ND - 2) x 8 L1 wavefronts of loads + 8 of stores.
//   Two warp-uniform specialisations keep the instruction count down (the first version issued ~280 instructions per
//   trip for 7 diagonals and ncu showed 61 % issue utilisation next to 50 % DRAM):
//   * INTERIOR — every column any lane touches is an own row of a complete quad: no clamping, no patching;
//     otherwise addresses are clamped into the block and the quads whose columns fall outside it (or into a
//     neighbour's boundary layer) are patched through gather_x;
//   * all presence bytes of the warp full (no domain face among its 128 rows): plain fma chain, no selects.
template <int ND, int MODE, bool RESID, bool SCALE, bool NORM, bool INTERIOR, bool COH = false>
__device__ __forceinline__ void cdia_stencil_trip(const SpmvArgs &a, long long q, int nvalid, int lane, double inv, int last4, double &nrm) {
  constexpr int C = ND / 2;   // main diagonal; C - 1 / C + 1 are the -1 / +1 neighbours
  constexpr int NF = ND - 3;  // far diagonals
  constexpr unsigned FULL = (ND == 5) ? 0x1f1f1f1fu : 0x7f7f7f7fu;
  const long long r = q * 4;
  // ---- all loads first
  double own[4], far[NF][4], bv[4], xm1, xp4;
  int cf[NF];
  const int rc = INTERIOR ? (int)r : ((r < last4) ? (int)r : last4);
  if (COH) ld4_l2(a.x + rc, own); else ld4_cached(a.x + rc, own);
#pragma unroll
  for (int f = 0; f < NF; f++) {
    cf[f] = (int)r + a.dia.off[(f < C - 1) ? f : f + 3];
    if (COH) ld4_l2(a.x + (INTERIOR ? cf[f] : min(max(cf[f], 0), last4)), far[f]);
    else ld4_cached(a.x + (INTERIOR ? cf[f] : min(max(cf[f], 0), last4)), far[f]);
  }
  if (RESID) ld4_cached(a.b + rc, bv);
  const unsigned m = (INTERIOR || nvalid) ? __ldg(reinterpret_cast<const unsigned *>(a.dmask) + q) : 0u;
  double edge = 0.0;
  if (INTERIOR) { // one predicated load (never a branch: it must leave with the others, not after the first use of `own`)
    const double *pe = a.x + r + ((lane == 0) ? -1 : 4);
    if (COH) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.cg.f64 %0, [%1];\n\t}" : "+d"(edge) : "l"(pe), "r"((int)(lane == 0 || lane == 31)) : "memory");
    else asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.nc.f64 %0, [%1];\n\t}" : "+d"(edge) : "l"(pe), "r"((int)(lane == 0 || lane == 31)));
  }
  // ---- own quad and its two outer neighbours
  if (SCALE) {
#pragma unroll
    for (int i = 0; i < 4; i++) own[i] = own[i] * inv;
    if (INTERIOR) edge = edge * inv;
  }
  if (!INTERIOR && nvalid != 4) {
#pragma unroll
    for (int i = 0; i < 4; i++) own[i] = (i < nvalid) ? gather_x<MODE, COH>(a, (int)r + i, inv, SCALE) : 0.0;
  }
  {
    const double up = __shfl_up_sync(0xffffffffu, own[3], 1), dn = __shfl_down_sync(0xffffffffu, own[0], 1);
    xm1 = (INTERIOR && lane == 0) ? edge : up;
    xp4 = (INTERIOR && lane == 31) ? edge : dn;
  }
  if (!INTERIOR) {
    if (nvalid && lane == 0) xm1 = gather_x<MODE, COH>(a, (int)r - 1, inv, SCALE);
    if (nvalid == 4) {
      if (lane == 31 || r + 4 >= a.nb) xp4 = gather_x<MODE, COH>(a, (int)r + 4, inv, SCALE); // the next lane holds no row
    } else xp4 = 0.0; // only row r + 3 would use it
  }
  // ---- far diagonals
#pragma unroll
  for (int f = 0; f < NF; f++) {
    if (SCALE) {
#pragma unroll
      for (int i = 0; i < 4; i++) far[f][i] = far[f][i] * inv;
    }
    if (!INTERIOR && (nvalid != 4 || cf[f] < 0 || cf[f] + 3 >= a.nb)) {
#pragma unroll
      for (int i = 0; i < 4; i++) far[f][i] = (i < nvalid) ? gather_x<MODE, COH>(a, cf[f] + i, inv, SCALE) : 0.0;
    }
  }
  // ---- fma chains in diagonal (= sorted column) order
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const bool full = INTERIOR && __all_sync(0xffffffffu, m == FULL);
#pragma unroll
  for (int k = 0; k < ND; k++) {
    const double ck = a.dconst[k];
    double xv[4];
    if (k == C) { xv[0] = own[0]; xv[1] = own[1]; xv[2] = own[2]; xv[3] = own[3]; }
    else if (k == C - 1) { xv[0] = xm1; xv[1] = own[0]; xv[2] = own[1]; xv[3] = own[2]; }
    else if (k == C + 1) { xv[0] = own[1]; xv[1] = own[2]; xv[2] = own[3]; xv[3] = xp4; }
    else {
#pragma unroll
      for (int i = 0; i < 4; i++) xv[i] = far[(k < C) ? k : k - 3][i];
    }
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; i++) s[i] = fma(ck, xv[i], s[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) s[i] = fma(((m >> (8 * i + k)) & 1u) ? ck : 0.0, xv[i], s[i]);
    }
  }
  if (INTERIOR || nvalid == 4) {
    if (RESID) {
#pragma unroll
      for (int i = 0; i < 4; i++) s[i] = bv[i] - s[i];
    }
    st4(a.y + r, s);
    if (NORM) { nrm = fma(s[0], s[0], fma(s[1], s[1], nrm)); nrm = fma(s[2], s[2], fma(s[3], s[3], nrm)); }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (i < nvalid) {
        if (RESID) s[i] = a.b[r + i] - s[i];
        a.y[r + i] = s[i];
        if (NORM) nrm = fma(s[i], s[i], nrm);
      }
  }
}
template <int ND, int MODE, bool RESID, bool SCALE, bool NORM>
__global__ void __launch_bounds__(MSPK_THREADS, (RESID || MODE != 0) ? 1 : (ND == 5 ? MSPK_CDIA_MINB5 : MSPK_CDIA_MINB7)) k_spmv_cdia_stencil(SpmvArgs a, ReduceWs ws, int ws_slot, GmresCtl *ctl_rw) {
  if (a.guard_it >= 0) {
    if (!a.ctl->active || a.ctl->it != a.guard_it) return;
  }
  const double inv = SCALE ? a.ctl->inv_arr[a.guard_it > 0 ? a.guard_it : 0] : 1.0;
  const int lane = threadIdx.x & 31;
  const int last4 = (a.nb - 4) & ~3; // last aligned quad that lies inside the block (nb >= 4)
  // rows whose every column (own quad +- 1, far quads) is an own row of the block
  const long long r_lo = -(long long)a.dia.off[0], r_hi = (long long)a.nb - 4 - a.dia.off[ND - 1];
  double nrm = 0.0;
  const long long nquads = ((long long)a.nb + 3) >> 2;
  // the trip count is uniform over a warp (the lanes exchange x entries by shuffle); lanes past the end idle.
  // (Tried and dropped: one contiguous row range per block instead of the grid stride, hoping for L1 hits on the +-nx
  // diagonals — 0.33 ms instead of 0.19 ms at 67 M rows: the +-D quads then miss L2 as well.)
  for (long long q0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x - lane); q0 < nquads; q0 += (long long)gridDim.x * blockDim.x) {
    const long long q = q0 + lane;
    const int nvalid = (q < nquads) ? (int)((a.nb - q * 4 < 4) ? (a.nb - q * 4) : 4) : 0;
    // warp-uniform without a vote: first and last lane of the warp decide (rows are consecutive across the lanes)
    const bool interior = (q0 * 4 >= r_lo) && ((q0 + 31) * 4 <= r_hi);
    if (interior) cdia_stencil_trip<ND, MODE, RESID, SCALE, NORM, true>(a, q, 4, lane, inv, last4, nrm);
    else cdia_stencil_trip<ND, MODE, RESID, SCALE, NORM, false>(a, q, nvalid, lane, inv, last4, nrm);
  }
  if (NORM) spmv_norm_epilogue(nrm, ws, ws_slot, ctl_rw);
}

// ------------------------------------------------------------------------------------------------
// K3  VecMDot: h[j] = <w, v_j>, j = 0..nv-1.  grid = (bx, ngroups): each y-group owns <= 8 vectors and
// streams them once together with w (8 n (nv + ngroups) bytes).  128-bit loads, 2x unrolled
// (up to 18 independent 16-B loads in flight per thread); last block reduces the partials in fixed order.
// ------------------------------------------------------------------------------------------------
struct MdotArgs {
  int nb, nv, per_group;
  long long ld;        // distance between consecutive basis vectors
  const double *V;     // basis, vector j at V + j*ld
  const double *w;
  double *h;           // result (device); written as sign * dot
  const double *inv;   // per-vector scale of the stored (un-normalised) basis, or null
  double sign;         // -1: PETSc's lhh = -<w, v>
  const GmresCtl *ctl;
  int guard_it, guard_refine; // guard_refine: run only if ctl->refine (second CGS pass)
};

// one warp sums the block partials of one vector: lane-strided over the blocks in index order, then the shuffle tree
__device__ __forceinline__ double mdot_sum_partials(const double *row, int nblocks, int lane) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = lane;
  for (; i + 96 < nblocks; i += 128) {
    s0 += __ldcg(row + i); s1 += __ldcg(row + i + 32); s2 += __ldcg(row + i + 64); s3 += __ldcg(row + i + 96);
  }
  for (; i < nblocks; i += 32) s0 += __ldcg(row + i);
  return warp_sum((s0 + s1) + (s2 + s3));
}

template <int NVMAX, int U>
__global__ void __launch_bounds__(MSPK_THREADS, NVMAX > 16 ? 1 : 2) k_mdot(MdotArgs a, ReduceWs ws) {
  if (a.guard_it >= 0) {
    if (!a.ctl->active || a.ctl->it != a.guard_it) return;
    if (a.guard_refine && !a.ctl->refine) return;
  }
  const int g = blockIdx.y;
  const int v0 = g * a.per_group;
  const int nv = min(a.per_group, a.nv - v0);
  const double *Vg = a.V + (long long)v0 * a.ld;
  double acc[NVMAX];
#pragma unroll
  for (int v = 0; v < NVMAX; v++) acc[v] = 0.0;
  const long long npairs = a.nb >> 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // U row-pairs per thread per trip: U * (NVMAX + 1) independent 16-byte loads in flight
  for (; p + (U - 1) * stride < npairs; p += U * stride) {
    double2 w[U], x[U][NVMAX];
#pragma unroll
    for (int u = 0; u < U; u++) w[u] = ld_stream2(a.w + 2 * (p + u * stride));
#pragma unroll
    for (int v = 0; v < NVMAX; v++)
      if (v < nv) {
#pragma unroll
        for (int u = 0; u < U; u++) x[u][v] = ld_stream2(Vg + v * a.ld + 2 * (p + u * stride));
      }
#pragma unroll
    for (int v = 0; v < NVMAX; v++)
      if (v < nv) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          acc[v] = fma(x[u][v].x, w[u].x, acc[v]);
          acc[v] = fma(x[u][v].y, w[u].y, acc[v]);
        }
      }
  }
  for (; p < npairs; p += stride) {
    const double2 w0 = ld_stream2(a.w + 2 * p);
#pragma unroll
    for (int v = 0; v < NVMAX; v++)
      if (v < nv) {
        const double2 x0 = ld_stream2(Vg + v * a.ld + 2 * p);
        acc[v] = fma(x0.x, w0.x, acc[v]);
        acc[v] = fma(x0.y, w0.y, acc[v]);
      }
  }
  if ((a.nb & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double wl = a.w[a.nb - 1];
#pragma unroll
    for (int v = 0; v < NVMAX; v++)
      if (v < nv) acc[v] = fma(Vg[v * a.ld + a.nb - 1], wl, acc[v]);
  }
  // ---- block partials: warp sums of all vectors, ONE barrier, thread v adds the 8 warp sums of vector v in warp order
  __shared__ double sm[(MSPK_THREADS / 32) * NVMAX];
  __shared__ bool last;
  const int slot = 8 + g; // reduce slots 8.. are MDot groups
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // partial sums: workspace rows 64 + g * NVMAX + v (192 rows in all; at most 64 + 24 + 23)
#pragma unroll
  for (int v = 0; v < NVMAX; v++) {
    if (v < nv) {
      const double ws_ = warp_sum(acc[v]);
      if (lane == 0) sm[wid * NVMAX + v] = ws_;
    }
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    double bs = 0.0;
#pragma unroll
    for (int w = 0; w < MSPK_THREADS / 32; w++) bs += sm[w * NVMAX + threadIdx.x];
    ws.partial[(64 + g * NVMAX + threadIdx.x) * (long long)MSPK_MAX_PART + blockIdx.x] = bs;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned t = atomicAdd(ws.counter + slot, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    // ---- last block to arrive: one warp per vector (all vectors in parallel), lane-strided sum in block order, then the
    // shuffle tree — a fixed order, hence deterministic.  (The first version walked the vectors one after the other with two
    // block barriers each: ~1.2 us per vector, 11 % of a 170 us launch at 8.4 M rows and 20 vectors.)
    __threadfence();
    for (int v = wid; v < nv; v += MSPK_THREADS / 32) {
      const double tot = mdot_sum_partials(ws.partial + (64 + g * NVMAX + v) * (long long)MSPK_MAX_PART, (int)gridDim.x, lane);
      // <w, v_j> = inv_j <w, vtilde_j>: the scale of the un-normalised basis vector is applied to the reduced value
      if (lane == 0) a.h[v0 + v] = a.sign * (a.inv ? tot * a.inv[v0 + v] : tot);
    }
    if (threadIdx.x == 0) ws.counter[slot] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K4+K5  VecMAXPY fused with VecNorm:  w += sum_j coef[j] v_j ;  ||w||^2 reduced in the same pass.
// Last block finishes the norm and runs the Hessenberg/Givens update + convergence test on the
// device (K6), so no host round trip is needed inside a restart cycle.
//   FIN 0: only store the norm       FIN 1: CGS pass epilogue (ctl_step_end / refinement decision)
//   FIN 2: store the SUM OF SQUARES (a Jacobi block spread over several GPUs sums it over its ranks, then k_step_end)
// ------------------------------------------------------------------------------------------------
struct MaxpyArgs {
  int nb, nv;
  long long ld;
  const double *V;
  const double *coef;  // device, nv coefficients
  const double *inv;   // per-vector scale of the stored (un-normalised) basis, or null
  double *w;
  double *norm_out;    // device scalar (may be null)
  GmresCtl *ctl;
  int guard_it, guard_refine;
  int pass;            // CGS pass index (0 or 1)
};

template <int FIN>
__global__ void __launch_bounds__(MSPK_THREADS) k_maxpy_norm(MaxpyArgs a, ReduceWs ws, int ws_slot) {
  if (a.guard_it >= 0) {
    if (!a.ctl->active || a.ctl->it != a.guard_it) return;
    if (a.guard_refine && !a.ctl->refine) return;
  }
  __shared__ double cf[MSPK_MAXK + 2];
  // w += sum_j coef_j v_j with v_j = inv_j vtilde_j: the scale is folded into the coefficient
  for (int j = threadIdx.x; j < a.nv; j += blockDim.x) cf[j] = a.inv ? a.coef[j] * a.inv[j] : a.coef[j];
  __syncthreads();
  double nrm = 0.0;
  const long long npairs = a.nb >> 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // two row-pairs per thread per trip, 8 basis vectors per chunk: 16 independent 16-byte loads in flight
  for (; p + stride < npairs; p += 2 * stride) {
    double2 t0 = __ldcs(reinterpret_cast<const double2 *>(a.w + 2 * p));
    double2 t1 = __ldcs(reinterpret_cast<const double2 *>(a.w + 2 * (p + stride)));
    int j = 0;
    for (; j + 8 <= a.nv; j += 8) {
      double2 x0[8], x1[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        x0[u] = ld_stream2(a.V + (long long)(j + u) * a.ld + 2 * p);
        x1[u] = ld_stream2(a.V + (long long)(j + u) * a.ld + 2 * (p + stride));
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double c = cf[j + u];
        t0.x = fma(c, x0[u].x, t0.x); t0.y = fma(c, x0[u].y, t0.y);
        t1.x = fma(c, x1[u].x, t1.x); t1.y = fma(c, x1[u].y, t1.y);
      }
    }
    if (j < a.nv) {
      double2 x0[8], x1[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (j + u < a.nv) {
          x0[u] = ld_stream2(a.V + (long long)(j + u) * a.ld + 2 * p);
          x1[u] = ld_stream2(a.V + (long long)(j + u) * a.ld + 2 * (p + stride));
        }
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (j + u < a.nv) {
          const double c = cf[j + u];
            t0.x = fma(c, x0[u].x, t0.x); t0.y = fma(c, x0[u].y, t0.y);
          t1.x = fma(c, x1[u].x, t1.x); t1.y = fma(c, x1[u].y, t1.y);
        }
    }
    __stcs(reinterpret_cast<double2 *>(a.w + 2 * p), t0);
    __stcs(reinterpret_cast<double2 *>(a.w + 2 * (p + stride)), t1);
    nrm = fma(t0.x, t0.x, fma(t0.y, t0.y, nrm));
    nrm = fma(t1.x, t1.x, fma(t1.y, t1.y, nrm));
  }
  for (; p < npairs; p += stride) {
    double2 t = *reinterpret_cast<const double2 *>(a.w + 2 * p);
    for (int j = 0; j < a.nv; j++) {
      const double2 x = ld_stream2(a.V + (long long)j * a.ld + 2 * p);
      t.x = fma(cf[j], x.x, t.x); t.y = fma(cf[j], x.y, t.y);
    }
    *reinterpret_cast<double2 *>(a.w + 2 * p) = t;
    nrm = fma(t.x, t.x, fma(t.y, t.y, nrm));
  }
  if ((a.nb & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double t = a.w[a.nb - 1];
    for (int j = 0; j < a.nv; j++) t = fma(cf[j], a.V[(long long)j * a.ld + a.nb - 1], t);
    a.w[a.nb - 1] = t;
    nrm = fma(t, t, nrm);
  }
  __shared__ double sm[32];
  __shared__ bool last;
  double bs = block_sum(nrm, sm);
  if (threadIdx.x == 0) {
    ws.partial[ws_slot * MSPK_MAX_PART + blockIdx.x] = bs;
    __threadfence();
    unsigned t = atomicAdd(ws.counter + ws_slot, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    double v = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(ws.partial + ws_slot * MSPK_MAX_PART + i);
    double tot = block_sum(v, sm);
    if (threadIdx.x == 0) {
      ws.counter[ws_slot] = 0;
      const double tt = sqrt(tot);
      if (a.norm_out) *a.norm_out = (FIN == 2) ? tot : tt; // FIN 2: the sum of squares, to be summed over the ranks of a block
      if (FIN == 1) ctl_cgs_pass_end(a.ctl, tt, a.pass);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K6  KSPGMRESBuildSoln: back substitution on the device (one thread), then
//     x += sum_j nrs[j] v_j fused with the boundary exchange: the first / last boundary layer of the
//     new iterate is stored straight into the neighbours' receive windows (P2P stores over NVLink when
//     the neighbour lives on another GPU) — replaces comm_sync_send_and_receive comm.c:126-141.
// ------------------------------------------------------------------------------------------------
__device__ inline void ctl_build_soln(GmresCtl *c) {
  const int ld = MSPK_MAXK + 2;
  const int k1 = c->it - 1;
  if (k1 < 0) return;
  c->its_total += c->it;
  bool bad = false;
  if (c->hh[(size_t)k1 * ld + k1] != 0.0) c->nrs[k1] = c->grs[k1] / c->hh[(size_t)k1 * ld + k1];
  else bad = true;
  for (int ii = 1; ii <= k1 && !bad; ii++) {
    const int k = k1 - ii;
    double t = c->grs[k];
    for (int j = k + 1; j <= k1; j++) t = t - c->hh[(size_t)j * ld + k] * c->nrs[j];
    if (c->hh[(size_t)k * ld + k] == 0.0) { bad = true; break; }
    c->nrs[k] = t / c->hh[(size_t)k * ld + k];
  }
  if (bad) { c->reason = -5; c->it = 0; /* no update */ }
}
__global__ void k_build_soln_coef(GmresCtl *c) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl_build_soln(c);
}

struct UpdateXArgs {
  int nb, H;
  long long ld;
  const double *V;
  double *x;
  const GmresCtl *ctl;     // nv = ctl->it, coefficients ctl->nrs
  double *peer_lo;         // lower neighbour's "upper boundary" receive window (or null)
  double *peer_hi;         // upper neighbour's "lower boundary" receive window (or null)
};

__global__ void __launch_bounds__(MSPK_THREADS) k_update_x(UpdateXArgs a) {
  const int nv = a.ctl->it;
  __shared__ double cf[MSPK_MAXK + 2];
  for (int j = threadIdx.x; j < nv; j += blockDim.x) cf[j] = a.ctl->nrs[j] * a.ctl->inv_arr[j];
  __syncthreads();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < a.nb; r += (long long)gridDim.x * blockDim.x) {
    double xv = a.x[r];
    if (nv > 0) {
      // PETSc: TEMP = 0; TEMP += sum nrs_j v_j; x += TEMP
      double t = 0.0;
      int j = 0;
      for (; j + 8 <= nv; j += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldg(a.V + (long long)(j + u) * a.ld + r);
#pragma unroll
        for (int u = 0; u < 8; u++) t = fma(cf[j + u], v[u], t);
      }
      for (; j < nv; j++) t = fma(cf[j], __ldg(a.V + (long long)j * a.ld + r), t);
      xv = xv + t;
      a.x[r] = xv;
    }
    if (a.peer_lo && r < a.H) a.peer_lo[r] = xv;
    if (a.peer_hi && r >= a.nb - a.H) a.peer_hi[r - (a.nb - a.H)] = xv;
  }
}

// synchronous exchange without a global barrier: after its boundary layers are stored, a block publishes the exchange
// sequence number into each neighbour's window (release: the layers are visible before the number); the neighbour's
// stream waits on that word (cuStreamWaitValue64, no kernel spins) before it collects the layers.  Only blocks K-1 / K+1
// ever wait for block K — replaces the MPI_Sendrecv pairing of comm.c:135 (and round 1's NCCL allreduce barrier).
__global__ void k_signal_neighbours(unsigned long long *flag_lo, unsigned long long *flag_hi, unsigned long long seq) {
  if (threadIdx.x || blockIdx.x) return;
  __threadfence_system();
  if (flag_lo) *reinterpret_cast<volatile unsigned long long *>(flag_lo) = seq;
  if (flag_hi) *reinterpret_cast<volatile unsigned long long *>(flag_hi) = seq;
  __threadfence_system();
}

// ---- a Jacobi block spread over several GPUs (the reference's npb > 1): the pieces around the block-wide reductions ----
// after the sum of squares of w has been summed over the ranks of the block: close the Gram-Schmidt pass of step guard_it
__global__ void k_step_end(GmresCtl *c, const double *sumsq, int pass, int guard_it, int guard_refine) {
  if (threadIdx.x || blockIdx.x) return;
  if (!c->active || c->it != guard_it) return;
  if (guard_refine && !c->refine) return;
  ctl_cgs_pass_end(c, sqrt(*sumsq), pass);
}
// first / last boundary layer of a vector into the intra-block neighbours' windows, scaled by *scale (the Krylov basis is
// stored un-normalised; the neighbour's SpMV does not scale what it reads from its halo).  guard_it as in the SpMV.
__global__ void k_publish_boundary_scaled(int nb, int H, const double *__restrict__ v, const double *scale, const GmresCtl *ctl, int guard_it,
                                          double *peer_lo, double *peer_hi) {
  if (guard_it >= 0 && (!ctl->active || ctl->it != guard_it)) return;
  const double sc = scale ? *scale : 1.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x) {
    if (peer_lo) peer_lo[i] = v[i] * sc;
    if (peer_hi) peer_hi[i] = v[nb - H + i] * sc;
  }
}

// publish only the boundary layers (used after the minimisation rewrote x, and by the closing exchange)
__global__ void k_publish_boundary(int nb, int H, const double *__restrict__ x, double *peer_lo, double *peer_hi) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x) {
    if (peer_lo) peer_lo[i] = x[i];
    if (peer_hi) peer_hi[i] = x[nb - H + i];
  }
}

// ------------------------------------------------------------------------------------------------
// K2  updateLocalRHS utils.c:943-948: rhs_K = b_K - A_KJ x_J, touching only the boundary rows
// ------------------------------------------------------------------------------------------------
__global__ void k_update_rhs(int nbrow, const int *__restrict__ brow, int nb, int W, int H, long long ld,
                             const int *__restrict__ ecol, const double *__restrict__ eval, const double *__restrict__ lo,
                             const double *__restrict__ hi, const double *__restrict__ b, double *__restrict__ rhs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbrow; i += gridDim.x * blockDim.x) {
    const int r = brow[i];
    double s = 0.0;
    for (int k = 0; k < W; k++) {
      const int c = ecol[k * ld + r];
      if (c < 0) s = fma(eval[k * ld + r], lo ? lo[c + H] : 0.0, s);
      else if (c >= nb) s = fma(eval[k * ld + r], hi ? hi[c - nb] : 0.0, s);
    }
    rhs[r] = b[r] - s;
  }
}

// ------------------------------------------------------------------------------------------------
// K8  R = A S tall-skinny SpMM: the matrix is streamed once for up to 8 columns of S.
//   MODE 0: A_KK S_K (local variant)        MODE 1: strip with the stored neighbour boundaries of
//   every iterate (global / semi-local): Slo/Shi hold s boundary layers each.
// ------------------------------------------------------------------------------------------------
struct SpmmArgs {
  int nb, W, H, s;
  long long ld, lds;      // ELL leading dim; distance between columns of S and of R
  const int *ecol;
  const double *eval;
  const double *S;        // s columns
  const double *Slo, *Shi; // s boundary layers of H values each (MODE 1)
  double *R;
};

// one row per thread (a two-rows-per-thread variant with 128-bit matrix loads measured 40 % slower: the 2 x NC x W
// scattered gathers per thread, not the matrix stream, bound this kernel)
template <int MODE, int NC>
__global__ void __launch_bounds__(MSPK_THREADS) k_spmm_ell(SpmmArgs a, int c0) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < a.nb; r += (long long)gridDim.x * blockDim.x) {
    double acc[NC];
#pragma unroll
    for (int t = 0; t < NC; t++) acc[t] = 0.0;
    for (int k = 0; k < a.W; k++) {
      const double v = a.eval[k * a.ld + r];
      const int c = a.ecol[k * a.ld + r];
      if ((unsigned)c < (unsigned)a.nb) {
#pragma unroll
        for (int t = 0; t < NC; t++) acc[t] = fma(v, __ldg(a.S + (long long)(c0 + t) * a.lds + c), acc[t]);
      } else if (MODE == 1) {
        const double *base = (c < 0) ? a.Slo : a.Shi;
        const int idx = (c < 0) ? c + a.H : c - a.nb;
        if (base) {
#pragma unroll
          for (int t = 0; t < NC; t++) acc[t] = fma(v, base[(long long)(c0 + t) * a.H + idx], acc[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NC; t++) a.R[(long long)(c0 + t) * a.lds + r] = acc[t];
  }
}

// ------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_copy(long long n, const double *__restrict__ src, double *__restrict__ dst) {
  const long long np = n >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < np; p += (long long)gridDim.x * blockDim.x)
    *reinterpret_cast<double2 *>(dst + 2 * p) = ld_stream2(src + 2 * p);
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = src[n - 1];
}
__global__ void k_fill(long long n, double v, double *dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}
// w *= (*inv_norm_src > 0 ? 1 / *norm : 0)   (VecNormalize tail used by the QR of the minimisation)
__global__ void k_scale_by_inv(long long n, const double *norm, double *w) {
  const double nv = *norm;
  const double inv = nv > 0.0 ? 1.0 / nv : 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) w[i] *= inv;
}
// w *= a  (VecScale)
__global__ void k_scale(long long n, double a, double *w) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) w[i] *= a;
}
// x[r] = sum_t alpha[t] S[t][r]   (MatMult(S, alpha, x) utils.c:1075), plus the stored boundaries
__global__ void k_lincomb(int nb, int s, long long lds, const double *__restrict__ S, const double *__restrict__ alpha, double *__restrict__ x) {
  __shared__ double al[64];
  for (int j = threadIdx.x; j < s; j += blockDim.x) al[j] = alpha[j];
  __syncthreads();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nb; r += (long long)gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int j = 0; j < s; j++) t = fma(S[(long long)j * lds + r], al[j], t);
    x[r] = t;
  }
}
// ------------------------------------------------------------------------------------------------
// K9  Gram contraction G = C^T C of the tall-skinny block C = [R_K | rhs] (NC = s+1 <= 9 columns): ONE pass over C
// (8 n NC bytes) for all NC(NC+1)/2 entries, accumulated in registers, last block reduces in fixed order.
// Intensity = (NC+1)/8 flop/B <= 1.25: far left of the fp64 ridge, so the kernel is HBM-bound on CUDA cores and
// fp64 tensor-core MMA (DMMA) has nothing to win here.
// ------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(MSPK_THREADS) k_gram(int nb, long long ld, const double *__restrict__ C, double *partial, unsigned int *counter,
                                                       double *out /* NC*NC, column-major, upper triangle filled */) {
  constexpr int NT = NC * (NC + 1) / 2;
  double acc[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) acc[t] = 0.0;
  const long long npairs = nb >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    double2 v[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) v[c] = ld_stream2(C + c * ld + 2 * p);
    int t = 0;
#pragma unroll
    for (int j = 0; j < NC; j++)
#pragma unroll
      for (int i = 0; i <= j; i++, t++) acc[t] = fma(v[i].y, v[j].y, fma(v[i].x, v[j].x, acc[t]));
  }
  if ((nb & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int j = 0; j < NC; j++)
#pragma unroll
      for (int i = 0; i <= j; i++, t++) acc[t] = fma(C[i * ld + nb - 1], C[j * ld + nb - 1], acc[t]);
  }
  __shared__ double sm[32];
  __shared__ bool last;
#pragma unroll
  for (int t = 0; t < NT; t++) {
    double bs = block_sum(acc[t], sm);
    if (threadIdx.x == 0) partial[(long long)t * MSPK_MAX_PART + blockIdx.x] = bs;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned q = atomicAdd(counter, 1u);
    last = (q == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    int t = 0;
    for (int j = 0; j < NC; j++)
      for (int i = 0; i <= j; i++, t++) {
        double v = 0.0;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) v += __ldcg(partial + (long long)t * MSPK_MAX_PART + b);
        double tot = block_sum(v, sm);
        if (threadIdx.x == 0) out[j * NC + i] = tot;
      }
    if (threadIdx.x == 0) *counter = 0;
  }
}

// C := C U^{-1} for an upper-triangular U (NC x NC, column-major in `U`): one read+write pass, two rows per thread
// (128-bit loads/stores), reciprocal pivots precomputed in shared memory.
template <int NC>
__global__ void __launch_bounds__(MSPK_THREADS) k_right_trsolve(int nb, long long ld, double *__restrict__ C, const double *__restrict__ U) {
  __shared__ double u[NC * NC], rinv[NC];
  for (int i = threadIdx.x; i < NC * NC; i += blockDim.x) u[i] = U[i];
  for (int i = threadIdx.x; i < NC; i += blockDim.x) rinv[i] = 1.0 / U[i * NC + i];
  __syncthreads();
  const long long npairs = nb >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    double2 y[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) y[c] = ld_stream2(C + c * ld + 2 * p);
#pragma unroll
    for (int j = 0; j < NC; j++) {
      double2 t = y[j];
#pragma unroll
      for (int i = 0; i < j; i++) { t.x = fma(-y[i].x, u[j * NC + i], t.x); t.y = fma(-y[i].y, u[j * NC + i], t.y); }
      y[j] = make_double2(t.x * rinv[j], t.y * rinv[j]);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) __stcs(reinterpret_cast<double2 *>(C + c * ld + 2 * p), y[c]);
  }
  if ((nb & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long r = nb - 1;
    double y[NC];
    for (int c = 0; c < NC; c++) y[c] = C[c * ld + r];
    for (int j = 0; j < NC; j++) {
      double t = y[j];
      for (int i = 0; i < j; i++) t = fma(-y[i], u[j * NC + i], t);
      y[j] = t * rinv[j];
    }
    for (int c = 0; c < NC; c++) C[c * ld + r] = y[c];
  }
}

// ------------------------------------------------------------------------------------------------
// Wide bases (s + 1 > 9 columns: the reference's s = 10 and s = 20 runs): the same CholeskyQR2, blocked in panels of
// <= 8 columns.  k_gram_panel contracts two panels (<= 8 x 8 inner products, 64 accumulators in registers, one pass over
// the <= 16 columns); k_apply_upper multiplies by the inverse Cholesky factor panel by panel, in place, last panel first
// (output panel j only needs the input panels <= j).  Bytes for 21 columns: 3 diagonal + 3 off-diagonal panel pairs =
// 576 n per Gram, 576 n per application, against ~8 000 n for Gram-Schmidt with reorthogonalisation column by column.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MSPK_THREADS) k_gram_panel(int nb, long long ld, const double *__restrict__ A, int na, const double *__restrict__ B, int nbc,
                                                             double *partial /* [64 x MSPK_MAX_PART] */, unsigned int *counter,
                                                             double *out, int ldo /* out[(b) * ldo + a] = <A_a, B_b> */) {
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b] = 0.0;
  const long long npairs = nb >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    double2 va[8], vb[8];
#pragma unroll
    for (int a = 0; a < 8; a++) va[a] = (a < na) ? ld_stream2(A + a * ld + 2 * p) : make_double2(0.0, 0.0);
#pragma unroll
    for (int b = 0; b < 8; b++) vb[b] = (b < nbc) ? ld_stream2(B + b * ld + 2 * p) : make_double2(0.0, 0.0);
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int b = 0; b < 8; b++) acc[a][b] = fma(va[a].y, vb[b].y, fma(va[a].x, vb[b].x, acc[a][b]));
  }
  if ((nb & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int a = 0; a < na; a++)
      for (int b = 0; b < nbc; b++) acc[a][b] = fma(A[a * ld + nb - 1], B[b * ld + nb - 1], acc[a][b]);
  }
  __shared__ double sm[32];
  __shared__ bool last;
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) {
      if (a < na && b < nbc) { // uniform over the block
        double bs = block_sum(acc[a][b], sm);
        if (threadIdx.x == 0) partial[(long long)(a * 8 + b) * MSPK_MAX_PART + blockIdx.x] = bs;
      }
    }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned q = atomicAdd(counter, 1u);
    last = (q == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    for (int a = 0; a < na; a++)
      for (int b = 0; b < nbc; b++) {
        double v = 0.0;
        for (int k = threadIdx.x; k < gridDim.x; k += blockDim.x) v += __ldcg(partial + (long long)(a * 8 + b) * MSPK_MAX_PART + k);
        double tot = block_sum(v, sm);
        if (threadIdx.x == 0) out[(long long)b * ldo + a] = tot;
      }
    if (threadIdx.x == 0) *counter = 0;
  }
}

// C[:, j0 .. j0+wj) := sum_{i < j0+wj} C[:, i] * T[i, j] for an upper-triangular T (nc x nc, column-major): in place,
// called for the LAST panel first.  T lives in shared memory (nc <= 33: 8.7 KB).
__global__ void __launch_bounds__(MSPK_THREADS) k_apply_upper(int nb, long long ld, double *__restrict__ C, int nc, int j0, int wj, const double *__restrict__ T) {
  extern __shared__ double t_sm[];
  for (int i = threadIdx.x; i < nc * nc; i += blockDim.x) t_sm[i] = T[i];
  __syncthreads();
  const long long npairs = nb >> 1;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    double2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = make_double2(0.0, 0.0);
    for (int i0 = 0; i0 < j0 + wj; i0 += 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (i0 + u < j0 + wj) ? ld_stream2(C + (long long)(i0 + u) * ld + 2 * p) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int c = 0; c < 8; c++)
          if (c < wj && i0 + u <= j0 + c) {
            const double t = t_sm[(j0 + c) * nc + i0 + u];
            acc[c].x = fma(v[u].x, t, acc[c].x); acc[c].y = fma(v[u].y, t, acc[c].y);
          }
    }
#pragma unroll
    for (int c = 0; c < 8; c++)
      if (c < wj) __stcs(reinterpret_cast<double2 *>(C + (long long)(j0 + c) * ld + 2 * p), acc[c]);
  }
  if ((nb & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long r = nb - 1;
    double acc[8];
    for (int c = 0; c < wj; c++) {
      double t = 0.0;
      for (int i = 0; i <= j0 + c; i++) t = fma(C[(long long)i * ld + r], t_sm[(j0 + c) * nc + i], t);
      acc[c] = t;
    }
    for (int c = 0; c < wj; c++) C[(long long)(j0 + c) * ld + r] = acc[c];
  }
}

// basis change of the minimisation: [x^1 .. x^s] -> [x^1, x^2-x^1, .., x^s-x^(s-1)] in place (same span, much better
// conditioned least-squares problem).  One pass, each thread owns a row.
__global__ void k_diff_basis(long long rows, int s, long long lds, double *__restrict__ S) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    double prev = S[r];
    for (int t = 1; t < s; t++) {
      const double cur = S[(long long)t * lds + r];
      S[(long long)t * lds + r] = cur - prev;
      prev = cur;
    }
  }
}
// sum of (x - 1)^2 and generic sum of squares through the norm slot
__global__ void __launch_bounds__(MSPK_THREADS) k_sumsq(long long n, const double *__restrict__ x, double shift, ReduceWs ws, int ws_slot, double *out) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = x[i] - shift;
    acc = fma(d, d, acc);
  }
  __shared__ double sm[32];
  __shared__ bool last;
  double bs = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    ws.partial[ws_slot * MSPK_MAX_PART + blockIdx.x] = bs;
    __threadfence();
    unsigned t = atomicAdd(ws.counter + ws_slot, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    double v = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(ws.partial + ws_slot * MSPK_MAX_PART + i);
    double tot = block_sum(v, sm);
    if (threadIdx.x == 0) { ws.counter[ws_slot] = 0; *out = tot; }
  }
}

__global__ void k_ctl_begin(GmresCtl *c, int restart, int max_it, int min_it, int initial_rtol, int guess_zero, int cgs_refine,
                            double rtol, double abstol, double divtol, const double *bnorm_sq) {
  if (threadIdx.x || blockIdx.x) return;
  c->restart = restart; c->max_it = max_it; c->min_it = min_it; c->initial_rtol = initial_rtol;
  c->guess_zero = guess_zero; c->cgs_refine = cgs_refine; c->rtol = rtol; c->abstol = abstol; c->divtol = divtol;
  c->bnorm = bnorm_sq ? sqrt(*bnorm_sq) : 0.0;
  c->its = 0; c->it = 0; c->reason = 0; c->active = 0; c->hapend = 0; c->refine = 0; c->first_cycle = 1;
  c->res = 0.0; c->ksp_rnorm = -1.0; c->gm_rnorm0 = 0.0; c->rnorm0 = 0.0; c->ttol = 0.0; c->inv = 0.0; c->tt = 0.0;
}
// cycle prologue when the residual vector is the right-hand side itself (zero initial guess)
__global__ void k_ctl_cycle_begin_from(GmresCtl *c, const double *sumsq) {
  if (threadIdx.x || blockIdx.x) return;
  ctl_cycle_begin(c, sqrt(*sumsq));
}

// ------------------------------------------------------------------------------------------------
// asynchronous convergence detection on the device (conv_detection_prime.c).
// One 256-byte control line per block; peers write their messages into `inbox` with system-scope
// stores (P2P over NVLink when on another GPU).  A mailbox keeps the last message only, which is
// what the reference's drain-the-queue receive handlers implement.
// ------------------------------------------------------------------------------------------------
struct CdMsg { int seq, a, b, pad; };       // seq increments on every send; consumed when seq != seen
struct CdMailbox { CdMsg m[2][4]; };         // [neighbour slot][PARTIAL_CV, VERIFICATION, RESPONSE, VERDICT]
struct CdState {
  int state, phase_tag, under, pp_begin, pp_end, local_cv, elected, partial_cv_sent, response_sent;
  int nb_not_recvd, nb_neighbors, me;
  int neighbors[2], recvd_pcv[2], responses[2], newer_dep[2], last_iter[2];
  int seen[2][4];       // last consumed seq per inbox cell
  int sent[2][4];       // last seq written to each neighbour cell
  int state_seen;       // state broadcast at the end of the previous outer iteration
  CdMailbox *inbox;     // own mailbox (written by peers)
  CdMailbox *outbox[2]; // neighbours' mailboxes (peer mapped); cell index = my slot in their table
  int my_slot_at[2];
  // legacy counter-based detector (conv_detection.c): see k_cd_legacy_step
  int lg_not_lcv, lg_count, lg_pre, lg_slcv, lg_global, lg_dest;
  int lg_prev_s[2], lg_prev_c[2];
};

// seqlock cell: seq odd = being written.  Writer: seq = 2q+1, payload, seq = 2q+2.  Reader: consume when seq is even,
// new, and unchanged after the payload has been read (otherwise the message is picked up at the next step, which is
// what a later MPI_Iprobe would do in the reference).
__device__ inline void cd_send(CdState *s, int nslot, int type, int a, int b) {
  CdMailbox *mb = s->outbox[nslot];
  if (!mb) return;
  CdMsg *m = &mb->m[s->my_slot_at[nslot]][type];
  const int q = ++s->sent[nslot][type];
  *reinterpret_cast<volatile int *>(&m->seq) = 2 * q - 1;
  __threadfence_system();
  *reinterpret_cast<volatile int *>(&m->a) = a;
  *reinterpret_cast<volatile int *>(&m->b) = b;
  __threadfence_system();
  *reinterpret_cast<volatile int *>(&m->seq) = 2 * q;
  __threadfence_system();
}
__device__ inline bool cd_recv(CdState *s, int nslot, int type, int *a, int *b) {
  volatile CdMsg *m = &s->inbox->m[nslot][type];
  const int q = m->seq;
  if ((q & 1) || q == s->seen[nslot][type]) return false;
  __threadfence_system();
  const int va = m->a, vb = m->b;
  __threadfence_system();
  if (m->seq != q) return false;
  *a = va; *b = vb;
  s->seen[nslot][type] = q;
  return true;
}
__device__ inline void cd_reinit_pp(CdState *s) { s->pp_begin = 0; s->pp_end = 0; for (int i = 0; i < 2; i++) s->newer_dep[i] = 0; }
__device__ inline void cd_init_state(CdState *s) {
  s->nb_not_recvd = s->nb_neighbors;
  for (int i = 0; i < 2; i++) s->recvd_pcv[i] = 0;
  s->elected = 0; s->local_cv = 0; s->partial_cv_sent = 0;
  cd_reinit_pp(s);
  s->state = 0;
}
__device__ inline void cd_init_verif(CdState *s) {
  cd_reinit_pp(s);
  s->phase_tag += 1;
  for (int i = 0; i < 2; i++) s->responses[i] = 0;
  s->response_sent = 0;
}
__device__ inline int cd_all_newer(const CdState *s) { for (int i = 0; i < s->nb_neighbors; i++) if (!s->newer_dep[i]) return 0; return 1; }
__device__ inline int cd_count(const CdState *s, int v) { int c = 0; for (int i = 0; i < s->nb_neighbors; i++) c += (s->responses[i] == v); return c; }

// receive_data_dependency conv_detection_prime.c:603-633
__device__ inline int cd_data_arrival(CdState *s, int nslot, int tag, int iter) {
  if (s->last_iter[nslot] < iter && (s->state_seen != 2 || tag == s->phase_tag)) {
    s->last_iter[nslot] = iter; s->newer_dep[nslot] = 1; return 1;
  }
  return 0;
}

__global__ void k_cd_step(CdState *s, int under, const double *local_norm_sq, double thr) {
  if (threadIdx.x || blockIdx.x) return;
  if (local_norm_sq) under = (sqrt(*local_norm_sq) <= thr) ? 1 : 0;
  s->under = under;
  const int NN = s->nb_neighbors;
  if (s->state == 0) {
    if (!s->under) cd_reinit_pp(s);
    else if (!s->pp_begin) s->pp_begin = 1;
    else if (s->pp_end) {
      s->local_cv = 1;
      if (s->nb_not_recvd == 0) {
        s->elected = 1; cd_init_verif(s);
        for (int i = 0; i < NN; i++) cd_send(s, i, 1, s->phase_tag, 0);
        s->state = 2;
      } else if (s->nb_not_recvd == 1) {
        for (int i = 0; i < NN; i++) if (!s->recvd_pcv[i]) { cd_send(s, i, 0, s->phase_tag, 0); break; }
        s->partial_cv_sent = 1; s->state = 1;
      }
    } else if (cd_all_newer(s)) s->pp_end = 1;
  } else if (s->state == 1) {
    // conv_detection_prime.c:84 compares a pointer with PETSC_FALSE: dead branch, replicated
  } else if (s->state == 2) {
    if (s->elected) {
      const int neg = cd_count(s, -1) > 0;
      if (!s->local_cv || neg) {
        s->phase_tag += 1;
        for (int i = 0; i < NN; i++) cd_send(s, i, 3, s->phase_tag, -1);
        cd_init_state(s);
      } else if (s->pp_end) {
        if (cd_count(s, 0) == 0) {
          if (cd_count(s, -1) == 0) { for (int i = 0; i < NN; i++) cd_send(s, i, 3, s->phase_tag, +1); s->state = 3; }
          else { s->phase_tag += 1; for (int i = 0; i < NN; i++) cd_send(s, i, 3, s->phase_tag, -1); cd_init_state(s); }
        }
      } else if (cd_all_newer(s)) s->pp_end = 1;
    } else if (!s->response_sent) {
      const int neg = cd_count(s, -1) > 0;
      if (!s->local_cv || neg) {
        for (int i = 0; i < NN; i++) if (!s->recvd_pcv[i]) { cd_send(s, i, 2, s->phase_tag, -1); break; }
        s->response_sent = 1;
      } else if (s->pp_end) {
        if (cd_count(s, 0) == 1) {
          int ask = -1;
          for (int i = 0; i < NN; i++) if (s->responses[i] == 0) { ask = i; break; }
          const int v = (cd_count(s, +1) == NN - 1) ? +1 : -1;
          cd_send(s, ask, 2, s->phase_tag, v);
          s->response_sent = 1;
        }
      } else if (cd_all_newer(s)) s->pp_end = 1;
    }
  }
  int a, b;
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 0, &a, &b)) {       // receive_partial_CV :314-370
    if (a == s->phase_tag) {
      s->recvd_pcv[sl] = 1; s->nb_not_recvd -= 1;
      const int src = s->neighbors[sl];
      const int leader = src > s->me ? src : s->me;
      if (s->nb_not_recvd == 0 && s->partial_cv_sent && leader == s->me) {
        s->elected = 1; cd_init_verif(s);
        for (int i = 0; i < NN; i++) cd_send(s, i, 1, s->phase_tag, 0);
        s->state = 2;
      }
    }
  }
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 1, &a, &b)) {       // receive_verification :373-411
    if (a == s->phase_tag + 1) {
      cd_init_verif(s); s->state = 2;
      for (int i = 0; i < NN; i++) if (i != sl) cd_send(s, i, 1, s->phase_tag, 0);
    }
  }
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 2, &a, &b)) {       // receive_response :414-448
    if (a == s->phase_tag) s->responses[sl] = b;
  }
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 3, &a, &b)) {       // receive_verdict :451-498
    if (b == +1) s->state = 3;
    else { cd_init_state(s); s->phase_tag = a; }
    for (int i = 0; i < NN; i++) if (i != sl) cd_send(s, i, 3, s->phase_tag, b);
  }
  s->state_seen = s->state;
}

// ------------------------------------------------------------------------------------------------
// Legacy counter-based termination (src/utils/conv_detection.c:6-173, driven as in
// src/asynchronous-multisplitting/asynchronous-multisplitting.c.save:283-329).  A block is "pre-converged"
// while its local residual is under the threshold; after MIN_CONVERGENCE_COUNT consecutive pre-converged iterations it is
// "strictly locally converged" (sLocalCV).  It then reports to the one neighbour that has not reported yet (SEND_CV carrying
// the iteration number), cancels that report if it loses the threshold (CANCEL_CV), and declares global convergence when
// every neighbour has reported while it is itself converged.  The host holds the loop for MAX_TRAVERSAL_TIME after
// globalCV and then tells the neighbours (GLOBAL_CV).  Mailbox cells: type 0 = SEND_CV, 1 = CANCEL_CV, 3 = GLOBAL_CV.
// Generalisation of the 2-block code (dest_node = neighbors[0], conv_detection.c:63): the report goes to the first
// neighbour whose newest message is not a report (prevIterNumS <= prevIterNumC).
// ------------------------------------------------------------------------------------------------
__global__ void k_cd_legacy_step(CdState *s, const double *local_norm_sq, double thr, int threshold_slcv, int current_iteration) {
  if (threadIdx.x || blockIdx.x) return;
  const int NN = s->nb_neighbors;
  s->lg_pre = (sqrt(*local_norm_sq) <= thr) ? 1 : 0;
  // comm_async_convDetection conv_detection.c:6-81
  if (!s->lg_slcv) {
    if (s->lg_pre) { s->lg_count += 1; if (s->lg_count == threshold_slcv) s->lg_slcv = 1; }
    else s->lg_count = 0;
  } else if (!s->lg_pre) {
    s->lg_slcv = 0; s->lg_count = 0;
    if (s->lg_dest != -1) cd_send(s, s->lg_dest, 1, current_iteration, 0);
  } else if (s->lg_not_lcv == 0) {
    s->lg_global = 1;
  } else if (s->lg_not_lcv == 1) {
    int dest = 0;
    for (int i = 0; i < NN; i++) if (s->lg_prev_s[i] <= s->lg_prev_c[i]) { dest = i; break; }
    s->lg_dest = dest;
    cd_send(s, dest, 0, current_iteration, 0);
  }
  int a, b;
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 0, &a, &b)) {   // comm_async_recvSPartialCV :83-113
    if (s->lg_prev_s[sl] < s->lg_prev_c[sl] && s->lg_prev_c[sl] < a) { s->lg_not_lcv -= 1; if (s->lg_not_lcv < 0) s->lg_not_lcv = 0; }
    if (s->lg_prev_s[sl] < a) s->lg_prev_s[sl] = a;
  }
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 1, &a, &b)) {   // comm_async_recvCancelSPartialCV :115-146
    if (s->lg_prev_c[sl] < s->lg_prev_s[sl] && s->lg_prev_s[sl] < a) {
      s->lg_not_lcv += 1; if (s->lg_not_lcv > NN) s->lg_not_lcv = NN;
      s->lg_global = 0;
    }
    if (s->lg_prev_c[sl] < a) s->lg_prev_c[sl] = a;
  }
  for (int sl = 0; sl < NN; sl++) if (cd_recv(s, sl, 3, &a, &b)) s->lg_global = a ? 1 : 0;   // comm_async_recvGlobalCV :148-160
  s->state = s->lg_global ? 1 : 0; // what the host reads: 1 = globalCV (it then runs the MAX_TRAVERSAL_TIME hold)
  s->state_seen = 0;
}
// comm_async_sendGlobalCV conv_detection.c:162-172, after the loop
__global__ void k_cd_legacy_send_global(CdState *s) {
  if (threadIdx.x || blockIdx.x) return;
  for (int i = 0; i < s->nb_neighbors; i++) cd_send(s, i, 3, 1, 0);
}

// ------------------------------------------------------------------------------------------------
// asynchronous boundary exchange (comm_async_test_and_send_prime / comm_async_probe_and_receive_prime,
// comm.c:455-554): the payload (one boundary layer) is stored into the neighbour's double-buffered receive
// window, then a header {seq, PhaseTag, iteration} is released.  The receiver takes the newest header it
// sees ("drain the queue, keep the last message"), asks receive_data_dependency whether to accept it, and
// copies the layer into its private halo.  A layer overwritten while it is being copied mixes two iterates
// of the neighbour, which an asynchronous iteration tolerates by construction.
// ------------------------------------------------------------------------------------------------
struct AsyncHdrDev { int seq, tag, iter, pad; };
struct ProbeDecision { int copy, buf, seq, pad; };

__global__ void k_async_release(AsyncHdrDev *peer_hdr, const CdState *cd, int iter, int seq) {
  if (threadIdx.x || blockIdx.x || !peer_hdr) return;
  __threadfence_system();
  *reinterpret_cast<volatile int *>(&peer_hdr->tag) = cd->phase_tag;
  *reinterpret_cast<volatile int *>(&peer_hdr->iter) = iter;
  __threadfence_system();
  *reinterpret_cast<volatile int *>(&peer_hdr->seq) = seq;
  __threadfence_system();
}
__global__ void k_async_probe(const AsyncHdrDev *hdr, CdState *cd, int nslot, int *seen_seq, ProbeDecision *dec) {
  if (threadIdx.x || blockIdx.x) return;
  const volatile AsyncHdrDev *h = hdr;
  const int q = h->seq;
  dec->copy = 0;
  if (q == *seen_seq) return;
  __threadfence_system();
  const int tag = h->tag, iter = h->iter;
  *seen_seq = q;
  dec->buf = q & 1;
  dec->seq = q;
  dec->copy = cd_data_arrival(cd, nslot, tag, iter);
}
__global__ void k_async_copy(const ProbeDecision *dec, int H, const double *win0, const double *win1, double *halo) {
  if (!dec->copy) return;
  const volatile double *src = dec->buf ? win1 : win0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x) halo[i] = src[i];
}
__global__ void k_cd_init(CdState *s, int me, int nblocks, CdMailbox *inbox, CdMailbox *out_lo, CdMailbox *out_hi,
                          const AsyncHdrDev *hdr_lo, const AsyncHdrDev *hdr_hi, int *seen_hdr) {
  if (threadIdx.x || blockIdx.x) return;
  // messages left over from a previous solve are not part of this one
  for (int sl = 0; sl < 2; sl++)
    for (int t = 0; t < 4; t++) s->seen[sl][t] = reinterpret_cast<volatile CdMsg *>(&inbox->m[sl][t])->seq;
  seen_hdr[0] = reinterpret_cast<const volatile AsyncHdrDev *>(hdr_lo)->seq;
  seen_hdr[1] = reinterpret_cast<const volatile AsyncHdrDev *>(hdr_hi)->seq;
  // keep `seen`/`sent` counters across solves: the mailboxes persist
  s->me = me;
  s->nb_neighbors = 0;
  s->outbox[0] = s->outbox[1] = nullptr;
  if (me > 0) { s->neighbors[s->nb_neighbors] = me - 1; s->outbox[s->nb_neighbors] = out_lo; s->my_slot_at[s->nb_neighbors] = (me - 1 > 0) ? 1 : 0; s->nb_neighbors++; }
  if (me < nblocks - 1) { s->neighbors[s->nb_neighbors] = me + 1; s->outbox[s->nb_neighbors] = out_hi; s->my_slot_at[s->nb_neighbors] = 0; s->nb_neighbors++; }
  s->inbox = inbox;
  for (int i = 0; i < 2; i++) { s->last_iter[i] = -1; s->responses[i] = 0; }
  cd_init_state(s);
  s->under = 0; s->phase_tag = 0; s->state_seen = 0; s->response_sent = 0;
  // legacy detector (asynchronous-multisplitting.c.save:136-151): nbNeigNotLCV = nbNeighbors, prevIterNumS = -1, prevIterNumC = 0
  s->lg_not_lcv = s->nb_neighbors; s->lg_count = 0; s->lg_pre = 0; s->lg_slcv = 0; s->lg_global = 0; s->lg_dest = -1;
  for (int i = 0; i < 2; i++) { s->lg_prev_s[i] = -1; s->lg_prev_c[i] = 0; }
}

// cycle_coop.cuh — ONE persistent cooperative kernel per GMRES restart cycle (included by engine.cu after kernels.cuh).
//
// Replaces, for small Jacobi blocks, the 1 + 3 j + 2 launches of a restart cycle of j steps (KSPGMRESCycle: prologue
// r = rhs - A x and its norm, per step SpMV / VecMDot / VecMAXPY + VecNorm + Hessenberg update, KSPGMRESBuildSoln and
// x += V y with the boundary publication of comm_sync_send_and_receive comm.c:126-141) by one launch whose phases are
// separated by grid-wide barriers.  At 131 072 rows per block (BASELINE config 1) a phase moves 1-30 MB out of L2 in 1-3 us
// and the one-kernel-per-phase path is bound by launch and drain latency (8-10 us per launch); here a step costs three
// barriers instead of three launches (five with a second Gram-Schmidt pass: -ksp_gmres_cgs_refinement_type refine_always,
// or refine_ifneeded when the first pass asks for it — decided on the device, identically in every block).
//
// Bit-identical to the one-kernel-per-phase path by construction:
//  * per-row arithmetic (the SpMV fma chain over the sorted row, the MAXPY fma chain in vector order, x += V y) does not
//    depend on which thread block handles the row;
//  * every reduction forms its block partials over the SAME virtual grid the separate kernel would be launched with
//    (msp_grid_for / mdot_geometry, shared with the host): a resident block loops over virtual block indices, each virtual
//    block accumulates its rows in the same order, sums over the block with the same tree and stores its partial in the same
//    workspace row; after the barrier the partials are summed in the same fixed order (mdot_sum_partials, block_sum);
//  * the control block (Hessenberg matrix, rotations, KSPConvergedDefault context) is updated by the same ctl_* functions —
//    by thread 0 of EVERY resident block on a shared-memory copy, so that no second barrier is needed to publish the decision
//    (the copies stay identical: same inputs, same deterministic arithmetic); block 0 writes it back at the end.
// Vectors written earlier in the same launch by other blocks are read through L2 (ld.global.cg; the COH flavour of the
// SpMV trip) — L1 is not coherent between SMs inside one launch.
//
// The barrier is a monotone arrival counter (fence + fire-and-forget relaxed reduction, relaxed polling + fence; the count a
// launch starts from is left behind by the previous one).  All blocks must be co-resident: the kernel is launched with
// cudaLaunchCooperativeKernel.  A block that waits longer than MSPK_COOP_TIMEOUT_NS raises the sticky abort word: every
// block leaves, the solve reports reason -100, later launches return at once and the host turns it into an error (never a
// hang).  No kernel of one engine ever waits for a kernel of another engine.
#pragma once

#include <cstdio>
#define MSPK_COOP_TIMEOUT_NS 8000000000ull
#define MSPK_COOP_REASON_ABORT (-100)

struct CycleCoopArgs {
  SpmvArgs sp;             // matrix view (coded DIA, stencil-shaped); x / y / b are set per phase
  int nb, H, nsteps, num_sms, mdot_gmax;
  int vg_prologue;         // virtual grid of the cycle-prologue SpMV (residual + norm)
  int vg_maxpy;            // virtual grid of VecMAXPY + VecNorm
  long long ld;
  double *V, *x;
  const double *rhs;
  GmresCtl *ctl;
  double *peer_lo, *peer_hi; // neighbours' receive windows for the new iterate's boundary layers (or null)
  ReduceWs ws;
  unsigned int *bar;       // [0] arrival counter, [1] sticky abort flag, [2] arrivals counted when the previous launch ended
};

__device__ __forceinline__ unsigned int coop_ld_relaxed(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long coop_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// fence.acq_rel is all a release / acquire pattern around a relaxed access needs; __threadfence() is the sequentially
// consistent fence (MEMBAR.SC.GPU)
__device__ __forceinline__ void coop_fence() {
#ifdef MSPK_COOP_SC_FENCE
  __threadfence();
#else
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
#ifdef MSPK_COOP_TIMING // debug build only: block 0 accumulates the time it spends per phase and prints it per launch
#define COOP_T(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const unsigned long long t_ = coop_globaltimer(); tacc[k] += t_ - tlast; tlast = t_; } } while (0)
#else
#define COOP_T(k) do { } while (0)
#endif
__device__ __forceinline__ double2 coop_ld2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

// grid-wide barrier; false = aborted (timeout here or in another block).  bar[0] counts arrivals and only ever grows:
// barrier k of a launch is passed once it has reached base + k * gridDim.x (`target`, kept by the caller); the arrival is a
// fire-and-forget reduction (no round trip) behind a fence, the wait polls with relaxed loads and ends with a fence.
__device__ __forceinline__ bool coop_barrier(unsigned int *bar, unsigned int &target, int *s_flag) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    coop_fence(); // release: this block's stores (ordered before by the block barrier) are visible before its arrival
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    int aborted = 0;
    unsigned int polls = 0;
    unsigned long long t0 = 0;
    while (true) {
      const unsigned int cur = coop_ld_relaxed(bar);
      if ((int)(cur - target) >= 0) break;
      if ((++polls & 0xfffu) == 0) {
        if (coop_ld_relaxed(bar + 1)) { aborted = 1; break; }
        const unsigned long long t = coop_globaltimer();
        if (t0 == 0) t0 = t;
        else if (t - t0 > MSPK_COOP_TIMEOUT_NS) { atomicExch(bar + 1, 1u); aborted = 1; break; }
      }
    }
    coop_fence(); // acquire: nothing of this block is read before the other blocks' arrivals were seen
    *s_flag = aborted;
  }
  __syncthreads();
  return *s_flag == 0;
}

// the block partials of up to four vectors summed by one warp, every load issued before the first shuffle: per vector the
// same additions in the same order as mdot_sum_partials
__device__ __forceinline__ void mdot_sum_partials4(const double *row0, long long row_stride, int nrows, int nblocks, int lane, double (&tot)[4]) {
  double s0[4], s1[4], s2[4], s3[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { s0[k] = 0.0; s1[k] = 0.0; s2[k] = 0.0; s3[k] = 0.0; }
  int i = lane;
  for (; i + 96 < nblocks; i += 128) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (k < nrows) {
        const double *row = row0 + k * row_stride;
        s0[k] += __ldcg(row + i); s1[k] += __ldcg(row + i + 32); s2[k] += __ldcg(row + i + 64); s3[k] += __ldcg(row + i + 96);
      }
  }
  for (; i < nblocks; i += 32) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (k < nrows) s0[k] += __ldcg(row0 + k * row_stride + i);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) tot[k] = (k < nrows) ? warp_sum((s0[k] + s1[k]) + (s2[k] + s3[k])) : 0.0;
}

// one pass of the stencil SpMV over the quads of virtual block vb of a grid of vg blocks (the loop of k_spmv_cdia_stencil)
template <int ND, bool RESID, bool SCALE, bool NORM>
__device__ __forceinline__ void coop_spmv_pass(const SpmvArgs &a, double inv, int vb, int vg, double &nrm) {
  const int lane = threadIdx.x & 31;
  const int last4 = (a.nb - 4) & ~3;
  const long long r_lo = -(long long)a.dia.off[0], r_hi = (long long)a.nb - 4 - a.dia.off[ND - 1];
  const long long nquads = ((long long)a.nb + 3) >> 2;
  for (long long q0 = vb * (long long)MSPK_THREADS + (threadIdx.x - lane); q0 < nquads; q0 += (long long)vg * MSPK_THREADS) {
    const long long q = q0 + lane;
    const int nvalid = (q < nquads) ? (int)((a.nb - q * 4 < 4) ? (a.nb - q * 4) : 4) : 0;
    const bool interior = (q0 * 4 >= r_lo) && ((q0 + 31) * 4 <= r_hi);
    if (interior) cdia_stencil_trip<ND, 0, RESID, SCALE, NORM, true, true>(a, q, 4, lane, inv, last4, nrm);
    else cdia_stencil_trip<ND, 0, RESID, SCALE, NORM, false, true>(a, q, nvalid, lane, inv, last4, nrm);
  }
}

#define MSPK_COOP_NV 8  // vectors per MDot work item
#define MSPK_COOP_NX 8  // vectors per MAXPY chunk
// Measured alternatives (profiles/r02_coop_probe.txt): the phases as separate non-inlined functions, 12 or 16 vectors per
// MAXPY chunk (more loads in flight, but spills under the 128-register cap): 3-30 % slower than this form.
#define MSPK_COOP_PHASE __forceinline__

// P0: vtilde_0 = rhs - A_KK x and the block partials of ||vtilde_0||^2 over the virtual grid of the prologue SpMV
template <int ND>
__device__ __forceinline__ void coop_phase_residual(const SpmvArgs &sp, int vg, double *partial, double *red) {
  for (int vb = blockIdx.x; vb < vg; vb += gridDim.x) {
    double nrm = 0.0;
    coop_spmv_pass<ND, true, false, true>(sp, 1.0, vb, vg, nrm);
    const double bs = block_sum(nrm, red);
    if (threadIdx.x == 0) partial[vb] = bs;
  }
}
// P1: w = A v_it with v_it = vtilde_it * inv on the fly
template <int ND>
__device__ __forceinline__ void coop_phase_spmv(const SpmvArgs &sp, double inv) {
  double unused = 0.0;
  coop_spmv_pass<ND, false, true, false>(sp, inv, blockIdx.x, gridDim.x, unused);
}
// sum of the block partials of a norm, thread-strided in index order then the block tree (valid in thread 0)
__device__ MSPK_COOP_PHASE double coop_sum_partials(const double *partial, int n, double *red) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += MSPK_THREADS) v += __ldcg(partial + i);
  return block_sum(v, red);
}
// P2: block partials of <w, vtilde_v>, v < nv.  The rows are split over the virtual x-grid of k_mdot (gx blocks: the partial
// of a vector depends on that split only, not on which vectors share a thread).  Work item = (virtual block, group of vector
// chunks): the chunks of MSPK_COOP_NV vectors are dealt out so that there are about as many items as resident blocks — a
// whole virtual block per item (w loaded once, ONE pair of block barriers) when gx already fills the grid, a few chunks per
// item when the block is so small that gx does not.
__device__ MSPK_COOP_PHASE void coop_phase_mdot(const double *V, long long ld, const double *w, int nb, int nv, int gx, double *partial, double *sm) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long npairs = nb >> 1;
  const long long stride = (long long)gx * MSPK_THREADS;
  const int nchunks = (nv + MSPK_COOP_NV - 1) / MSPK_COOP_NV;
  int ngrp = (int)gridDim.x / gx; // groups of chunks per virtual block
  ngrp = ngrp < 1 ? 1 : (ngrp > nchunks ? nchunks : ngrp);
  const int cpg = (nchunks + ngrp - 1) / ngrp; // chunks per group
  ngrp = (nchunks + cpg - 1) / cpg;
  for (int item = blockIdx.x; item < gx * ngrp; item += gridDim.x) {
    const int grp = item / gx, vbx = item - grp * gx;
    const int vfirst = grp * cpg * MSPK_COOP_NV;
    const int vend = min(nv, vfirst + cpg * MSPK_COOP_NV);
    for (int v0 = vfirst; v0 < vend; v0 += MSPK_COOP_NV) {
      const int nvc = min(MSPK_COOP_NV, vend - v0);
      const double *Vc = V + (long long)v0 * ld;
      double acc[MSPK_COOP_NV];
#pragma unroll
      for (int v = 0; v < MSPK_COOP_NV; v++) acc[v] = 0.0;
      for (long long p = vbx * (long long)MSPK_THREADS + tid; p < npairs; p += stride) {
        const double2 w0 = coop_ld2(w + 2 * p);
        double2 xv[MSPK_COOP_NV];
#pragma unroll
        for (int v = 0; v < MSPK_COOP_NV; v++)
          if (v < nvc) xv[v] = coop_ld2(Vc + v * ld + 2 * p);
#pragma unroll
        for (int v = 0; v < MSPK_COOP_NV; v++)
          if (v < nvc) {
            acc[v] = fma(xv[v].x, w0.x, acc[v]);
            acc[v] = fma(xv[v].y, w0.y, acc[v]);
          }
      }
      if ((nb & 1) && vbx == 0 && tid == 0) {
        const double wl = __ldcg(w + nb - 1);
#pragma unroll
        for (int v = 0; v < MSPK_COOP_NV; v++)
          if (v < nvc) acc[v] = fma(__ldcg(Vc + v * ld + nb - 1), wl, acc[v]);
      }
#pragma unroll
      for (int v = 0; v < MSPK_COOP_NV; v++) {
        if (v < nvc) {
          const double ws_ = warp_sum(acc[v]);
          if (lane == 0) sm[wid * (MSPK_MAXK + 2) + (v0 - vfirst) + v] = ws_;
        }
      }
    }
    __syncthreads();
    if (tid < vend - vfirst) { // thread v adds the eight warp sums of vector vfirst + v in warp order
      double bs = 0.0;
#pragma unroll
      for (int ww = 0; ww < MSPK_THREADS / 32; ww++) bs += sm[ww * (MSPK_MAXK + 2) + tid];
      partial[(64 + vfirst + tid) * (long long)MSPK_MAX_PART + vbx] = bs;
    }
    __syncthreads();
  }
}
// every block finishes every dot product: lhh[v] = -<w, v_v> = -(inv_v <w, vtilde_v>); warp `wid` takes the vectors
// wid, wid + 8, wid + 16, wid + 24 at once
__device__ MSPK_COOP_PHASE void coop_phase_mdot_final(const double *partial, int nv, int gx, const double *inv_arr, double *lhh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int NW = MSPK_THREADS / 32;
  for (int vb0 = wid; vb0 < nv; vb0 += 4 * NW) {
    const int nrows = (nv - vb0 + NW - 1) / NW;
    double tot[4];
    mdot_sum_partials4(partial + (64 + vb0) * (long long)MSPK_MAX_PART, (long long)NW * MSPK_MAX_PART, nrows < 4 ? nrows : 4, gx, lane, tot);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int v = vb0 + k * NW;
        if (v < nv) lhh[v] = -1.0 * (tot[k] * inv_arr[v]);
      }
    }
  }
}
// P3: w += sum_j cf_j vtilde_j and the block partials of ||w||^2 over the virtual grid of k_maxpy_norm
__device__ MSPK_COOP_PHASE void coop_phase_maxpy(const double *V, long long ld, double *w, int nb, int nv, int vg, const double *cf, double *partial,
                                              double *red) {
  const int tid = threadIdx.x;
  const long long npairs = nb >> 1;
  const long long stride = (long long)vg * MSPK_THREADS;
  for (int vb = blockIdx.x; vb < vg; vb += gridDim.x) {
    double nrm = 0.0;
    for (long long p = vb * (long long)MSPK_THREADS + tid; p < npairs; p += stride) {
      double2 t = coop_ld2(w + 2 * p);
      for (int j = 0; j < nv; j += MSPK_COOP_NX) { // every load of a chunk leaves before the fma chain (in vector order) starts
        double2 x0[MSPK_COOP_NX];
#pragma unroll
        for (int u = 0; u < MSPK_COOP_NX; u++)
          if (j + u < nv) x0[u] = coop_ld2(V + (long long)(j + u) * ld + 2 * p);
#pragma unroll
        for (int u = 0; u < MSPK_COOP_NX; u++)
          if (j + u < nv) { t.x = fma(cf[j + u], x0[u].x, t.x); t.y = fma(cf[j + u], x0[u].y, t.y); }
      }
      *reinterpret_cast<double2 *>(w + 2 * p) = t;
      nrm = fma(t.x, t.x, fma(t.y, t.y, nrm));
    }
    if ((nb & 1) && vb == 0 && tid == 0) {
      double t = __ldcg(w + nb - 1);
      for (int j = 0; j < nv; j++) t = fma(cf[j], __ldcg(V + (long long)j * ld + nb - 1), t);
      w[nb - 1] = t;
      nrm = fma(t, t, nrm);
    }
    const double bs = block_sum(nrm, red);
    if (tid == 0) partial[vb] = bs;
  }
}
// x += sum_j cf_j vtilde_j (PETSc: TEMP = 0; TEMP += sum nrs_j v_j; x += TEMP) and the boundary layers of the new iterate
// stored into the neighbours' receive windows
__device__ MSPK_COOP_PHASE void coop_phase_update_x(const double *V, long long ld, double *x, int nb, int H, int nv, const double *cf, double *peer_lo,
                                                 double *peer_hi) {
  for (long long r = blockIdx.x * (long long)MSPK_THREADS + threadIdx.x; r < nb; r += (long long)gridDim.x * MSPK_THREADS) {
    double xv = __ldcg(x + r);
    if (nv > 0) {
      double t = 0.0;
      int j = 0;
      for (; j + 8 <= nv; j += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcg(V + (long long)(j + u) * ld + r);
#pragma unroll
        for (int u = 0; u < 8; u++) t = fma(cf[j + u], v[u], t);
      }
      for (; j < nv; j++) t = fma(cf[j], __ldcg(V + (long long)j * ld + r), t);
      xv = xv + t;
      x[r] = xv;
    }
    if (peer_lo && r < H) peer_lo[r] = xv;
    if (peer_hi && r >= nb - H) peer_hi[r - (nb - H)] = xv;
  }
}

template <int ND>
__global__ void __launch_bounds__(MSPK_THREADS, 2) k_gmres_cycle_coop(CycleCoopArgs a) {
  __shared__ GmresCtl sc;
  __shared__ double sm[(MSPK_THREADS / 32) * (MSPK_MAXK + 2)];
  __shared__ double cf[MSPK_MAXK + 2];
  __shared__ double red[32];
  __shared__ int s_flag;
  static_assert(offsetof(GmresCtl, hh) % 8 == 0 && sizeof(GmresCtl) % 8 == 0, "GmresCtl is copied in 8-byte words");
  const int tid = threadIdx.x;

  if (coop_ld_relaxed(a.bar + 1)) { // an earlier launch timed out in a barrier: the arrival count cannot be trusted any more
    if (blockIdx.x == 0 && tid == 0) { a.ctl->reason = MSPK_COOP_REASON_ABORT; a.ctl->active = 0; a.ctl->it = 0; }
    return;
  }
  unsigned int bar_target = coop_ld_relaxed(a.bar + 2); // arrivals counted by the earlier launches
#ifdef MSPK_COOP_TIMING
  unsigned long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = coop_globaltimer();
#endif
  // ---- private copy of the control block (everything but the Hessenberg columns, which this cycle rewrites before use)
  {
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.ctl);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(&sc);
    for (int i = tid; i < (int)(offsetof(GmresCtl, hh) / 8); i += MSPK_THREADS) dst[i] = __ldcg(src + i);
  }
  __syncthreads();

  // ---- P0: vtilde_0 = rhs - A_KK x, ||vtilde_0||; cycle-begin logic (KSPGMRESCycle prologue)
  {
    SpmvArgs sp = a.sp;
    sp.x = a.x; sp.y = a.V; sp.b = a.rhs;
    coop_phase_residual<ND>(sp, a.vg_prologue, a.ws.partial, red);
  }
  if (!coop_barrier(a.bar, bar_target, &s_flag)) goto aborted;
  {
    const double tot = coop_sum_partials(a.ws.partial, a.vg_prologue, red);
    if (tid == 0) {
      if (blockIdx.x == 0) a.ws.partial[MSPK_MAX_PART - 1] = tot;
      ctl_cycle_begin(&sc, sqrt(tot));
    }
    __syncthreads();
  }

  for (int it = 0; it < a.nsteps; it++) {
    if (!sc.active) break; // identical in every block
    double *w = a.V + (long long)(it + 1) * a.ld;
    const int nv = it + 1;
    // ---- P1: w = A v_it
    {
      SpmvArgs sp = a.sp;
      sp.x = a.V + (long long)it * a.ld; sp.y = w; sp.b = nullptr;
      coop_phase_spmv<ND>(sp, sc.inv_arr[it]);
    }
    COOP_T(1);
    if (!coop_barrier(a.bar, bar_target, &s_flag)) goto aborted;
    COOP_T(2);
    // classical Gram-Schmidt: one pass, or two with -ksp_gmres_cgs_refinement_type refine_always / refine_ifneeded (the
    // decision of the first pass, ctl->refine, is the same in every block)
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1 && !sc.refine) break;
      // ---- P2: lhh = -V^T w
      const MdotGeom gm = mdot_geometry(a.nb, nv, a.mdot_gmax, a.num_sms);
      coop_phase_mdot(a.V, a.ld, w, a.nb, nv, gm.gx, a.ws.partial, sm);
      COOP_T(3);
      if (!coop_barrier(a.bar, bar_target, &s_flag)) goto aborted;
      COOP_T(4);
      coop_phase_mdot_final(a.ws.partial, nv, gm.gx, sc.inv_arr, sc.lhh);
      __syncthreads();
      for (int j = tid; j < nv; j += MSPK_THREADS) cf[j] = sc.lhh[j] * sc.inv_arr[j];
      __syncthreads();
      COOP_T(5);
      // ---- P3: w += V lhh, ||w||; Hessenberg / Givens update, KSPConvergedDefault, next `active`
      coop_phase_maxpy(a.V, a.ld, w, a.nb, nv, a.vg_maxpy, cf, a.ws.partial, red);
      COOP_T(6);
      if (!coop_barrier(a.bar, bar_target, &s_flag)) goto aborted;
      COOP_T(7);
      {
        const double tot = coop_sum_partials(a.ws.partial, a.vg_maxpy, red);
        COOP_T(8);
        if (sc.cgs_refine == 0) {
          // ctl_cgs_pass_end for REFINE_NEVER: hh[j] = 0 - lhh[j] element by element (the sum of squares it also forms
          // only decides about a second pass), then the step is closed by one thread
          if (tid == 0) red[0] = sqrt(tot);
          for (int j = tid; j <= it; j += MSPK_THREADS) sc.hh[(size_t)it * (MSPK_MAXK + 2) + j] = 0.0 - sc.lhh[j];
          __syncthreads();
          if (tid == 0) { sc.refine = 0; ctl_step_end(&sc, red[0]); }
        } else if (tid == 0) {
          ctl_cgs_pass_end(&sc, sqrt(tot), pass); // asks for the second pass (ctl->refine) or closes the step
        }
        __syncthreads();
      }
      COOP_T(9);
    }
  }
  COOP_T(0);
  // ---- KSPGMRESBuildSoln: back substitution (every block, on its own copy), x += sum_j nrs_j v_j, boundary publication
  {
    const int cols = sc.it; // Hessenberg columns this cycle wrote (ctl_build_soln may reset `it` on a zero pivot)
    __syncthreads();
    if (tid == 0) ctl_build_soln(&sc);
    __syncthreads();
    const int nv = sc.it;
    for (int j = tid; j < nv; j += MSPK_THREADS) cf[j] = sc.nrs[j] * sc.inv_arr[j];
    __syncthreads();
    coop_phase_update_x(a.V, a.ld, a.x, a.nb, a.H, nv, cf, a.peer_lo, a.peer_hi);
    // ---- block 0 publishes the control block
    if (blockIdx.x == 0) {
      const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&sc);
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.ctl);
      for (int i = tid; i < (int)(offsetof(GmresCtl, hh) / 8); i += MSPK_THREADS) dst[i] = src[i];
      const int ldh = MSPK_MAXK + 2;
      for (int i = tid; i < cols * ldh; i += MSPK_THREADS) a.ctl->hh[i] = sc.hh[i];
      if (tid == 0) a.bar[2] = bar_target; // where the next launch starts counting
    }
  }
#ifdef MSPK_COOP_TIMING
  COOP_T(10);
  if (blockIdx.x == 0 && tid == 0 && a.nsteps >= 30)
    printf("COOPT steps %d ns: prologue+exit %llu | spmv %llu bar1 %llu | mdot %llu bar2 %llu final %llu | maxpy %llu bar3 %llu sum %llu ctl %llu | update_x %llu\n", sc.it,
           tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[9], tacc[10]);
#endif
  return;

aborted:
  if (blockIdx.x == 0 && tid == 0) { a.ctl->reason = MSPK_COOP_REASON_ABORT; a.ctl->active = 0; a.ctl->it = 0; }
}

/*
 * msplit_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see msplit_oracle.h).
 *
 * Restates, in plain C, the reference's multisplitting solve path.  Every
 * function cites the reference file:line (under /root/reference/src/) or the
 * PETSc 3.22.1 routine (un-vendored; restated from its published algorithm,
 * SURVEY.md Appendix A) it follows.  Nothing here is shipped or timed as the
 * product; bench.py times it only as the CPU baseline.
 */
#include "msplit_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
#define CHUNK 4096 /* fixed reduction chunk: results independent of the OpenMP thread count */

/* ======================================================================= */
/* assembly                                                                */
/* ======================================================================= */

/* utils/utils.c:652-666 computeDimensionRelatedVariables */
API int orc_dimension_related(int nprocs, int npb, int rank, int m, int n, int *njacobi_blocks, int *rank_jacobi_block,
                              int *proc_local_rank, int *n_mesh_points, int *jacobi_block_size) {
  if (npb <= 0 || nprocs <= 0) return 1;
  *njacobi_blocks = nprocs / npb;
  *rank_jacobi_block = rank / npb;
  *proc_local_rank = rank % npb;
  *n_mesh_points = m * n;
  *jacobi_block_size = (*n_mesh_points) / (*njacobi_blocks);
  return 0;
}

/* utils/utils.c:247-293 poisson2DMatrix.  Row Ii of the global 5-point matrix,
 * i = Ii / n_cols, j = Ii % n_cols; PETSc AIJ stores each row sorted by
 * global column: [Ii-n, Ii-1, Ii, Ii+1, Ii+n] where in range. */
static int row2d(int m, int n, int64_t Ii, int32_t *cols, double *vals) {
  int64_t i = Ii / n, j = Ii - i * n;
  int k = 0;
  if (i > 0) { cols[k] = (int32_t)(Ii - n); vals[k++] = -1.0; }
  if (j > 0) { cols[k] = (int32_t)(Ii - 1); vals[k++] = -1.0; }
  cols[k] = (int32_t)Ii; vals[k++] = 4.0;
  if (j < n - 1) { cols[k] = (int32_t)(Ii + 1); vals[k++] = -1.0; }
  if (i < m - 1) { cols[k] = (int32_t)(Ii + n); vals[k++] = -1.0; }
  return k;
}

API int64_t orc_poisson2d_nnz(int m, int n, int block, int nblocks) {
  int64_t nb = ((int64_t)m * n) / nblocks, r0 = nb * block, nnz = 0;
  int32_t c[5]; double v[5];
  for (int64_t Ii = r0; Ii < r0 + nb; Ii++) nnz += row2d(m, n, Ii, c, v);
  return nnz;
}

API int orc_poisson2d(int m, int n, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val) {
  if (nblocks <= 0 || block < 0 || block >= nblocks) return 1;
  int64_t nb = ((int64_t)m * n) / nblocks, r0 = nb * block, nnz = 0;
  rowptr[0] = 0;
  for (int64_t r = 0; r < nb; r++) {
    nnz += row2d(m, n, r0 + r, colidx + nnz, val + nnz);
    rowptr[r + 1] = (int32_t)nnz;
  }
  return 0;
}

/* utils/utils.c:383-445 poisson2DMatrix_complete: row stride N = n_mesh_lines (:390,:397),
 * i.e. the reference assumes a square mesh; we refuse anything else. */
API int orc_poisson2d_complete(int m, int n, int32_t *rowptr, int32_t *colidx, double *val) {
  if (m != n) return 2;
  return orc_poisson2d(m, n, 0, 1, rowptr, colidx, val);
}

/* utils/utils.c:30-121 poisson3DMatrix: row = i + j*nx + k*nx*ny (:63), diag 6, -1 at +-1, +-nx, +-nx*ny.
 * Block split: reference hard-wires 2 blocks and splits with n_grid_columns/2 (:45,:51: cubes only);
 * generalised here to z-slabs of nz/nblocks planes (identical for cubes at 2 blocks). */
static int row3d(int nx, int ny, int nz, int64_t row, int32_t *cols, double *vals) {
  int64_t pl = (int64_t)nx * ny;
  int64_t k = row / pl, rem = row - k * pl, j = rem / nx, i = rem - j * nx;
  int c = 0;
  if (k > 0) { cols[c] = (int32_t)(row - pl); vals[c++] = -1.0; }
  if (j > 0) { cols[c] = (int32_t)(row - nx); vals[c++] = -1.0; }
  if (i > 0) { cols[c] = (int32_t)(row - 1); vals[c++] = -1.0; }
  cols[c] = (int32_t)row; vals[c++] = 6.0;
  if (i < nx - 1) { cols[c] = (int32_t)(row + 1); vals[c++] = -1.0; }
  if (j < ny - 1) { cols[c] = (int32_t)(row + nx); vals[c++] = -1.0; }
  if (k < nz - 1) { cols[c] = (int32_t)(row + pl); vals[c++] = -1.0; }
  return c;
}

API int64_t orc_poisson3d_nnz(int nx, int ny, int nz, int block, int nblocks) {
  int64_t nb = ((int64_t)nx * ny * nz) / nblocks, r0 = nb * block, nnz = 0;
  int32_t c[7]; double v[7];
  for (int64_t r = r0; r < r0 + nb; r++) nnz += row3d(nx, ny, nz, r, c, v);
  return nnz;
}

API int orc_poisson3d(int nx, int ny, int nz, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val) {
  if (nblocks <= 0 || block < 0 || block >= nblocks) return 1;
  int64_t nb = ((int64_t)nx * ny * nz) / nblocks, r0 = nb * block, nnz = 0;
  rowptr[0] = 0;
  for (int64_t r = 0; r < nb; r++) {
    nnz += row3d(nx, ny, nz, r0 + r, colidx + nnz, val + nnz);
    rowptr[r + 1] = (int32_t)nnz;
  }
  return 0;
}

/* utils/utils.c:450-478 divideSubDomainIntoBlockMatrices (MatCreateSubMatrix with a stride
 * column IS): keep columns in [col_lo, col_hi), renumber to col - col_lo. */
API int64_t orc_submatrix_nnz(int nrows, const int32_t *rowptr, const int32_t *colidx, int col_lo, int col_hi) {
  int64_t nnz = 0;
  for (int64_t k = 0; k < rowptr[nrows]; k++) nnz += (colidx[k] >= col_lo && colidx[k] < col_hi);
  return nnz;
}

API int orc_submatrix(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val, int col_lo, int col_hi,
                      int32_t *out_rowptr, int32_t *out_colidx, double *out_val) {
  int64_t nnz = 0;
  out_rowptr[0] = 0;
  for (int r = 0; r < nrows; r++) {
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++)
      if (colidx[k] >= col_lo && colidx[k] < col_hi) {
        out_colidx[nnz] = colidx[k] - col_lo;
        out_val[nnz++] = val[k];
      }
    out_rowptr[r + 1] = (int32_t)nnz;
  }
  return 0;
}

/* ======================================================================= */
/* vector / matrix kernels                                                 */
/* ======================================================================= */

/* PETSc MatMult_SeqAIJ: per row, sequential sum over the sorted non-zeros.  FMA is used
 * explicitly so that the CUDA kernels (which use __fma_rn in the same order) are bit-identical. */
API void orc_spmv(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nrows; r++) {
    double sum = 0.0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) sum = fma(val[k], x[colidx[k]], sum);
    y[r] = sum;
  }
}

/* PETSc MatResidual default: r = b - (A x)  (MatMult then VecAYPX(r,-1,b)) */
API void orc_residual(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val, const double *b,
                      const double *x, double *r) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nrows; i++) {
    double sum = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; k++) sum = fma(val[k], x[colidx[k]], sum);
    r[i] = b[i] - sum;
  }
}

API double orc_dot(int64_t n, const double *a, const double *b) {
  int64_t nch = (n + CHUNK - 1) / CHUNK;
  double total = 0.0;
  if (nch <= 1) {
    for (int64_t i = 0; i < n; i++) total = fma(a[i], b[i], total);
    return total;
  }
  double *part = (double *)malloc(sizeof(double) * nch);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nch; c++) {
    int64_t lo = c * CHUNK, hi = lo + CHUNK < n ? lo + CHUNK : n;
    double s = 0.0;
    for (int64_t i = lo; i < hi; i++) s = fma(a[i], b[i], s);
    part[c] = s;
  }
  for (int64_t c = 0; c < nch; c++) total += part[c];
  free(part);
  return total;
}

API double orc_norm2(int64_t n, const double *a) { return sqrt(orc_dot(n, a, a)); }

/* the per-block term of utils/utils.c:575-595 computeFinalResidualNorm: ||b_K - A_K,: x||_2 */
API double orc_block_residual_norm(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val,
                                   const double *b, const double *x) {
  double *r = (double *)malloc(sizeof(double) * (size_t)nrows);
  orc_residual(nrows, rowptr, colidx, val, b, x, r);
  double nrm = orc_norm2(nrows, r);
  free(r);
  return nrm;
}

static void axpy(int64_t n, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) y[i] = fma(a, x[i], y[i]);
}
static void scale(int64_t n, double a, double *x) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) x[i] *= a;
}
/* y += sum_j a[j] * X[j]   (VecMAXPY) */
static void maxpy(int64_t n, int nv, const double *a, double *const *X, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double t = y[i];
    for (int j = 0; j < nv; j++) t = fma(a[j], X[j][i], t);
    y[i] = t;
  }
}

/* ======================================================================= */
/* KSPConvergedDefault (PETSc src/ksp/ksp/interface/iterativ.c) */
/* ======================================================================= */
typedef struct {
  const orc_ksp_opts *o;
  double rnorm0, ttol;
  double bnorm; /* ||b|| for the nonzero-guess, non-UIR case */
  int guess_zero;
} cvg_ctx;

static int converged_default(cvg_ctx *c, int n, double rnorm) {
  const orc_ksp_opts *o = c->o;
  if (n == 0) {
    if (!c->guess_zero && !o->initial_rtol) {
      double snorm = c->bnorm;
      if (snorm == 0.0) snorm = rnorm;
      c->rnorm0 = snorm;
    } else {
      c->rnorm0 = rnorm;
    }
    c->ttol = fmax(o->rtol * c->rnorm0, o->abstol);
  }
  if (n <= o->min_it) return ORC_CONVERGED_ITERATING;
  if (isnan(rnorm) || isinf(rnorm)) return ORC_DIVERGED_NANORINF;
  if (rnorm <= c->ttol) return (rnorm < o->abstol) ? ORC_CONVERGED_ATOL : ORC_CONVERGED_RTOL;
  if (rnorm >= o->divtol * c->rnorm0) return ORC_DIVERGED_DTOL;
  return ORC_CONVERGED_ITERATING;
}

API void orc_ksp_defaults(orc_ksp_opts *o) {
  /* tmp/petscmpiexec_help:336-342,602-607,615 */
  o->restart = 30; o->max_it = 10000; o->rtol = 1e-5; o->abstol = 1e-50; o->divtol = 1e4;
  o->initial_rtol = 0; o->guess_nonzero = 0; o->cgs_refine = 0; o->mgs = 0; o->min_it = 0;
}

/* ======================================================================= */
/* restarted GMRES  (PETSc gmres.c KSPSolve_GMRES / KSPGMRESCycle /         */
/* KSPGMRESUpdateHessenberg / KSPGMRESBuildSoln, borthog2.c) */
/* ======================================================================= */
static struct { int n, m; double **V; } g_gmres_ws = {0, 0, NULL};
API void orc_gmres_release_workspace(void) {
  if (g_gmres_ws.V) {
    for (int i = 0; i < g_gmres_ws.m + 2; i++) free(g_gmres_ws.V[i]);
    free(g_gmres_ws.V);
  }
  g_gmres_ws.V = NULL; g_gmres_ws.n = g_gmres_ws.m = 0;
}
static double **gmres_workspace(int n, int m) {
  if (g_gmres_ws.V && g_gmres_ws.n == n && g_gmres_ws.m == m) return g_gmres_ws.V;
  orc_gmres_release_workspace();
  g_gmres_ws.V = (double **)malloc(sizeof(double *) * (m + 2));
  for (int i = 0; i < m + 2; i++) g_gmres_ws.V[i] = (double *)malloc(sizeof(double) * (size_t)n);
  g_gmres_ws.n = n; g_gmres_ws.m = m;
  return g_gmres_ws.V;
}
API int orc_gmres(int n, const int32_t *rowptr, const int32_t *colidx, const double *val, const double *b, double *x,
                  const orc_ksp_opts *o, int *its_out, int *reason_out, double *rnorm_out, double *hist, int hist_cap) {
  const int m = o->restart;
  const double haptol = 1e-30, breakdowntol = 0.1;
  /* Krylov workspace kept between calls (PETSc allocates VEC_VV once per KSP, not once per KSPSolve): at 67 M rows a
   * fresh malloc per inner solve would page-fault ~11 GB every time and inflate the CPU baseline. */
  double **V = gmres_workspace(n, m);
  double *temp = V[m + 1];
  /* HH is (m+1) x m, column-major with leading dimension m+1 */
  double *HH = (double *)calloc((size_t)(m + 2) * (m + 1), sizeof(double));
  double *cc = (double *)calloc(m + 2, sizeof(double)), *ss = (double *)calloc(m + 2, sizeof(double));
  double *grs = (double *)calloc(m + 2, sizeof(double)), *lhh = (double *)calloc(m + 2, sizeof(double));
#define H(i, j) HH[(size_t)(j) * (m + 2) + (i)]
  cvg_ctx cv = {o, 0.0, 0.0, 0.0, !o->guess_nonzero};
  if (!cv.guess_zero && !o->initial_rtol) cv.bnorm = orc_norm2(n, b);

  int its = 0, itcount = 0, reason = 0, nhist = 0;
  double ksp_rnorm = -1.0, gm_rnorm0 = 0.0;
  int guess_zero = cv.guess_zero;

  while (!reason) {
    /* KSPInitialResidual: r = b - A x (nonzero guess) or r = b */
    if (guess_zero) memcpy(V[0], b, sizeof(double) * (size_t)n);
    else orc_residual(n, rowptr, colidx, val, b, x, V[0]);

    /* ---- KSPGMRESCycle ---- */
    int it = 0, hapend = 0;
    double res = orc_norm2(n, V[0]);
    if (res > 0.0) scale(n, 1.0 / res, V[0]); /* VecNormalize */
    if (isnan(res) || isinf(res)) { reason = ORC_DIVERGED_NANORINF; break; }
    if (ksp_rnorm > 0.0 && fabs(res - ksp_rnorm) > breakdowntol * gm_rnorm0) { reason = ORC_DIVERGED_BREAKDOWN; break; }
    grs[0] = gm_rnorm0 = res;
    ksp_rnorm = res;
    if (hist && nhist < hist_cap) hist[nhist++] = res;
    if (res == 0.0) { reason = ORC_CONVERGED_ATOL; break; }
    reason = converged_default(&cv, its, res);
    while (!reason && it < m && its < o->max_it) {
      if (it && hist && nhist < hist_cap) hist[nhist++] = res;
      orc_spmv(n, rowptr, colidx, val, V[it], V[it + 1]); /* PC none */
      /* orthogonalisation */
      for (int j = 0; j <= it; j++) H(j, it) = 0.0;
      if (o->mgs) {
        for (int j = 0; j <= it; j++) {
          double d = orc_dot(n, V[it + 1], V[j]);
          H(j, it) = d;
          axpy(n, -d, V[j], V[it + 1]);
        }
      } else {
        int passes = (o->cgs_refine == 2) ? 2 : 1;
        for (int pass = 0; pass < passes; pass++) {
          for (int j = 0; j <= it; j++) lhh[j] = -orc_dot(n, V[it + 1], V[j]); /* VecMDot, negated */
          maxpy(n, it + 1, lhh, V, V[it + 1]);
          for (int j = 0; j <= it; j++) H(j, it) -= lhh[j];
          if (pass == 0 && o->cgs_refine == 1) {
            double hnrm = 0.0;
            for (int j = 0; j <= it; j++) hnrm += lhh[j] * lhh[j];
            hnrm = sqrt(hnrm);
            double wnrm = orc_norm2(n, V[it + 1]);
            if (wnrm < hnrm) passes = 2;
          }
        }
      }
      double tt = orc_norm2(n, V[it + 1]);
      if (tt > 0.0) scale(n, 1.0 / tt, V[it + 1]);
      if (isnan(tt) || isinf(tt)) { reason = ORC_DIVERGED_NANORINF; break; }
      H(it + 1, it) = tt;
      double hapbnd = fabs(tt / grs[it]);
      if (hapbnd > haptol) hapbnd = haptol;
      if (tt < hapbnd) hapend = 1;
      /* KSPGMRESUpdateHessenberg */
      {
        double *hh = &H(0, it);
        for (int j = 1; j <= it; j++) {
          double t = hh[j - 1];
          hh[j - 1] = cc[j - 1] * t + ss[j - 1] * hh[j];
          hh[j] = cc[j - 1] * hh[j] - ss[j - 1] * t;
        }
        if (!hapend) {
          double t2 = sqrt(hh[it] * hh[it] + hh[it + 1] * hh[it + 1]);
          if (t2 == 0.0) { reason = ORC_DIVERGED_NULL; }
          else {
            cc[it] = hh[it] / t2;
            ss[it] = hh[it + 1] / t2;
            grs[it + 1] = -(ss[it] * grs[it]);
            grs[it] = cc[it] * grs[it];
            hh[it] = cc[it] * hh[it] + ss[it] * hh[it + 1];
            res = fabs(grs[it + 1]);
          }
        } else {
          res = 0.0;
        }
      }
      it++;
      its++;
      ksp_rnorm = res;
      if (reason) break;
      reason = converged_default(&cv, its, res);
      if (hapend && !reason) { reason = ORC_DIVERGED_BREAKDOWN; break; }
    }
    /* KSPGMRESBuildSoln(it - 1) — always, even when cut short by max_it */
    if (it > 0) {
      int k1 = it - 1, bad = 0;
      double *nrs = lhh;
      if (H(k1, k1) != 0.0) nrs[k1] = grs[k1] / H(k1, k1);
      else { bad = 1; }
      for (int ii = 1; ii <= k1 && !bad; ii++) {
        int k = k1 - ii;
        double t = grs[k];
        for (int j = k + 1; j <= k1; j++) t = t - H(k, j) * nrs[j];
        if (H(k, k) == 0.0) { bad = 1; break; }
        nrs[k] = t / H(k, k);
      }
      if (bad) { reason = ORC_DIVERGED_BREAKDOWN; }
      else {
        memset(temp, 0, sizeof(double) * (size_t)n);
        maxpy(n, it, nrs, V, temp);
        axpy(n, 1.0, temp, x);
      }
    }
    if (hist && reason && nhist < hist_cap) hist[nhist++] = res;
    itcount += it;
    if (itcount >= o->max_it) {
      if (!reason) reason = ORC_DIVERGED_ITS;
      break;
    }
    guess_zero = 0;
  }
#undef H
  if (its_out) *its_out = its;
  if (reason_out) *reason_out = reason;
  if (rnorm_out) *rnorm_out = ksp_rnorm;
  free(HH); free(cc); free(ss); free(grs); free(lhh);
  return 0;
}

/* ======================================================================= */
/* LSQR on a dense column-major R  (PETSc lsqr.c KSPSolve_LSQR)  */
/* called as utils/utils.c:1061-1078 outer_solver_norm_equation does:       */
/* UIR norm, zero initial guess, PC none.                                   */
/* ======================================================================= */
static void dense_mv(int64_t nrows, int s, const double *R, int64_t ld, const double *v, double *u) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; i++) {
    double t = 0.0;
    for (int j = 0; j < s; j++) t = fma(R[(size_t)j * ld + i], v[j], t);
    u[i] = t;
  }
}
static void dense_mtv(int64_t nrows, int s, const double *R, int64_t ld, const double *u, double *v) {
  for (int j = 0; j < s; j++) v[j] = orc_dot(nrows, R + (size_t)j * ld, u);
}

API int orc_lsqr(int64_t nrows, int s, const double *R, int64_t ldr, const double *b, double *x, const orc_outer_opts *o,
                 int *its_out, int *reason_out, double *rnorm_out) {
  double *U = (double *)malloc(sizeof(double) * (size_t)nrows), *U1 = (double *)malloc(sizeof(double) * (size_t)nrows);
  double *V = (double *)calloc(s, sizeof(double)), *V1 = (double *)calloc(s, sizeof(double)), *W = (double *)calloc(s, sizeof(double));
  orc_ksp_opts ko;
  orc_ksp_defaults(&ko);
  ko.rtol = o->rtol; ko.abstol = o->abstol; ko.initial_rtol = 1; ko.max_it = o->max_it;
  cvg_ctx cv = {&ko, 0.0, 0.0, 0.0, 1};
  int reason = 0, its = 0;
  for (int j = 0; j < s; j++) x[j] = 0.0;
  memcpy(U, b, sizeof(double) * (size_t)nrows);
  double rnorm = orc_norm2(nrows, U), anorm = 0.0, arnorm;
  reason = converged_default(&cv, 0, rnorm);
  if (!reason && rnorm > 0.0) {
    double beta = rnorm, alpha;
    scale(nrows, 1.0 / beta, U);
    dense_mtv(nrows, s, R, ldr, U, V);
    alpha = orc_norm2(s, V);
    if (alpha > 0.0) for (int j = 0; j < s; j++) V[j] /= alpha;
    memcpy(W, V, sizeof(double) * s);
    arnorm = alpha * beta;
    double phibar = beta, rhobar = alpha;
    int i = 0;
    do {
      dense_mv(nrows, s, R, ldr, V, U1);
      axpy(nrows, -alpha, U, U1);
      beta = orc_norm2(nrows, U1);
      if (beta > 0.0) {
        scale(nrows, 1.0 / beta, U1);
        anorm = sqrt(anorm * anorm + alpha * alpha + beta * beta);
      }
      dense_mtv(nrows, s, R, ldr, U1, V1);
      for (int j = 0; j < s; j++) V1[j] = fma(-beta, V[j], V1[j]);
      alpha = orc_norm2(s, V1);
      if (alpha > 0.0) for (int j = 0; j < s; j++) V1[j] /= alpha;
      double rho = sqrt(rhobar * rhobar + beta * beta);
      double c = rhobar / rho, sn = beta / rho;
      double theta = sn * alpha;
      rhobar = -c * alpha;
      double phi = c * phibar;
      phibar = sn * phibar;
      double tau = sn * phi;
      for (int j = 0; j < s; j++) x[j] = fma(phi / rho, W[j], x[j]);
      for (int j = 0; j < s; j++) W[j] = V1[j] + (-theta / rho) * W[j]; /* VecAYPX */
      arnorm = alpha * fabs(tau);
      rnorm = phibar;
      its++;
      reason = converged_default(&cv, i + 1, rnorm);
      if (!reason && !o->lsqr_default_test) { /* KSPLSQRConvergedDefault */
        if (arnorm < ko.abstol) reason = ORC_CONVERGED_ATOL_NORMAL;
        else if (arnorm < ko.rtol * anorm * rnorm) reason = ORC_CONVERGED_RTOL_NORMAL;
      }
      if (reason) break;
      double *t = U1; U1 = U; U = t;
      t = V1; V1 = V; V = t;
      i++;
    } while (i < o->max_it);
    if (i >= o->max_it && !reason) reason = ORC_DIVERGED_ITS;
  }
  (void)arnorm;
  if (its_out) *its_out = its;
  if (reason_out) *reason_out = reason;
  if (rnorm_out) *rnorm_out = rnorm;
  free(U); free(U1); free(V); free(V1); free(W);
  return 0;
}

/* exact least squares by Householder QR (the "exact LS" minimiser of BASELINE config 3);
 * rnorm = ||b - R alpha||_2 from the trailing part of Q^T b. */
API int orc_lstsq_qr(int64_t nrows, int s, const double *R, int64_t ldr, const double *b, double *alpha, double *rnorm) {
  if (nrows < s) return 1;
  double *A = (double *)malloc(sizeof(double) * (size_t)nrows * s);
  double *c = (double *)malloc(sizeof(double) * (size_t)nrows);
  for (int j = 0; j < s; j++) memcpy(A + (size_t)j * nrows, R + (size_t)j * ldr, sizeof(double) * (size_t)nrows);
  memcpy(c, b, sizeof(double) * (size_t)nrows);
  for (int k = 0; k < s; k++) {
    double *a = A + (size_t)k * nrows;
    double nrm = orc_norm2(nrows - k, a + k);
    if (nrm == 0.0) { free(A); free(c); return 2; }
    double akk = a[k], beta = (akk >= 0.0) ? -nrm : nrm;
    /* v = a[k:] - beta e1, stored in place; H = I - 2 v v^T / (v^T v) */
    a[k] = akk - beta;
    double vtv = orc_dot(nrows - k, a + k, a + k);
    for (int j = k + 1; j < s; j++) {
      double *aj = A + (size_t)j * nrows;
      double f = 2.0 * orc_dot(nrows - k, a + k, aj + k) / vtv;
      axpy(nrows - k, -f, a + k, aj + k);
    }
    double f = 2.0 * orc_dot(nrows - k, a + k, c + k) / vtv;
    axpy(nrows - k, -f, a + k, c + k);
    /* store U(k,k) out of band: keep v in place, remember beta in alpha[] temporarily */
    alpha[k] = beta;
  }
  /* back substitution: U(k,k) = alpha[k] (beta), U(k,j) = A[j*nrows + k] for j > k */
  double diag[64];
  if (s > 64) { free(A); free(c); return 3; }
  for (int k = 0; k < s; k++) diag[k] = alpha[k];
  for (int k = s - 1; k >= 0; k--) {
    double t = c[k];
    for (int j = k + 1; j < s; j++) t -= A[(size_t)j * nrows + k] * alpha[j];
    alpha[k] = t / diag[k];
  }
  if (rnorm) *rnorm = orc_norm2(nrows - s, c + s);
  free(A); free(c);
  return 0;
}

/* ======================================================================= */
/* asynchronous convergence detection  (utils/conv_detection_prime.c,       */
/* Algorithm 5.15 of Bahi/Contassot-Vivier/Couturier).    */
/* Spanning tree over block roots = chain K-1, K, K+1 (reference: 2 blocks,  */
/* utils/conv_detection.c:180-196).  "Messages" are last-value mailboxes:    */
/* every reference handler drains its queue and keeps the last message       */
/* (:328-332, :383-388, :429-434, :466-470).                                 */
/* ======================================================================= */
enum { ST_NORMAL = 0, ST_WAIT4VERIFICATION = 1, ST_VERIFICATION = 2, ST_FINISHED = 3 };
enum { MSG_PARTIAL_CV = 0, MSG_VERIFICATION = 1, MSG_RESPONSE = 2, MSG_VERDICT = 3 };
typedef struct { int valid, a, b; } cd_msg;
typedef struct {
  int state, phase_tag, under, pp_begin, pp_end, local_cv, elected, partial_cv_sent, response_sent;
  int nb_not_recvd, nb_neighbors, neighbors[2], recvd_pcv[2], responses[2];
  int nb_deps, newer_dep[2], last_iter[2];
  cd_msg inbox[2][4]; /* [neighbour slot][message type] */
} cd_node;
struct orc_cd { int nblocks; cd_node *nd; };

static int slot_of(const cd_node *n, int rank) {
  for (int i = 0; i < n->nb_neighbors; i++) if (n->neighbors[i] == rank) return i;
  return -1;
}
static void cd_send(orc_cd *cd, int from, int to, int type, int a, int b) {
  cd_node *dst = &cd->nd[to];
  int sl = slot_of(dst, from);
  if (sl < 0) return;
  dst->inbox[sl][type].valid = 1; dst->inbox[sl][type].a = a; dst->inbox[sl][type].b = b;
}
/* :275-287 */
static void reinit_pseudo_period(cd_node *n) {
  n->pp_begin = 0; n->pp_end = 0;
  for (int i = 0; i < n->nb_deps; i++) n->newer_dep[i] = 0;
}
/* :251-266 */
static void initialize_state(cd_node *n) {
  n->nb_not_recvd = n->nb_neighbors;
  for (int i = 0; i < n->nb_neighbors; i++) n->recvd_pcv[i] = 0;
  n->elected = 0; n->local_cv = 0; n->partial_cv_sent = 0;
  reinit_pseudo_period(n);
  n->state = ST_NORMAL;
}
/* :300-312 */
static void initialize_verification(cd_node *n) {
  reinit_pseudo_period(n);
  n->phase_tag += 1;
  for (int i = 0; i < n->nb_neighbors; i++) n->responses[i] = 0;
  n->response_sent = 0;
}
static int all_newer(const cd_node *n) {
  for (int i = 0; i < n->nb_deps; i++) if (!n->newer_dep[i]) return 0;
  return 1;
}
static int count_resp(const cd_node *n, int v) {
  int c = 0;
  for (int i = 0; i < n->nb_neighbors; i++) c += (n->responses[i] == v);
  return c;
}

API orc_cd *orc_cd_create(int nblocks) {
  orc_cd *cd = (orc_cd *)calloc(1, sizeof(orc_cd));
  cd->nblocks = nblocks;
  cd->nd = (cd_node *)calloc(nblocks, sizeof(cd_node));
  for (int k = 0; k < nblocks; k++) {
    cd_node *n = &cd->nd[k];
    n->nb_neighbors = 0;
    if (k > 0) n->neighbors[n->nb_neighbors++] = k - 1;
    if (k < nblocks - 1) n->neighbors[n->nb_neighbors++] = k + 1;
    n->nb_deps = n->nb_neighbors;
    for (int i = 0; i < 2; i++) { n->last_iter[i] = -1; n->responses[i] = 0; }
    initialize_state(n); /* …multisplitting_prime.c:196-198 */
    n->under = 0; n->phase_tag = 0;
  }
  return cd;
}
API void orc_cd_destroy(orc_cd *cd) { if (cd) { free(cd->nd); free(cd); } }
API int orc_cd_state(const orc_cd *cd, int k) { return cd->nd[k].state; }
API int orc_cd_phase_tag(const orc_cd *cd, int k) { return cd->nd[k].phase_tag; }

/* :603-633 receive_data_dependency */
API int orc_cd_data_arrival(orc_cd *cd, int k, int src, int src_tag, int src_iter) {
  cd_node *n = &cd->nd[k];
  int sl = slot_of(n, src);
  if (sl < 0) return 0;
  if (n->last_iter[sl] < src_iter && (n->state != ST_VERIFICATION || src_tag == n->phase_tag)) {
    n->last_iter[sl] = src_iter;
    n->newer_dep[sl] = 1;
    return 1;
  }
  return 0;
}

API void orc_cd_step(orc_cd *cd, int k, int under) {
  cd_node *n = &cd->nd[k];
  n->under = under;
  /* ---- comm_async_convDetection_prime :11-249 ---- */
  if (n->state == ST_NORMAL) {
    if (!n->under) reinit_pseudo_period(n);
    else if (!n->pp_begin) n->pp_begin = 1;
    else if (n->pp_end) {
      n->local_cv = 1;
      if (n->nb_not_recvd == 0) {
        n->elected = 1;
        initialize_verification(n);
        for (int i = 0; i < n->nb_neighbors; i++) cd_send(cd, k, n->neighbors[i], MSG_VERIFICATION, n->phase_tag, 0);
        n->state = ST_VERIFICATION;
      } else if (n->nb_not_recvd == 1) {
        for (int i = 0; i < n->nb_neighbors; i++)
          if (!n->recvd_pcv[i]) { cd_send(cd, k, n->neighbors[i], MSG_PARTIAL_CV, n->phase_tag, 0); break; }
        n->partial_cv_sent = 1;
        n->state = ST_WAIT4VERIFICATION;
      }
    } else if (all_newer(n)) n->pp_end = 1;
  } else if (n->state == ST_WAIT4VERIFICATION) {
    /* :84 compares the POINTER UnderThreashold with PETSC_FALSE: never true; branch is dead. Replicated. */
  } else if (n->state == ST_VERIFICATION) {
    if (n->elected) {
      int neg = count_resp(n, -1) > 0;
      /* :97 pointer compare: the "!under" term never fires */
      if (!n->local_cv || neg) {
        n->phase_tag += 1;
        for (int i = 0; i < n->nb_neighbors; i++) cd_send(cd, k, n->neighbors[i], MSG_VERDICT, n->phase_tag, -1);
        initialize_state(n);
      } else if (n->pp_end) {
        if (count_resp(n, 0) == 0) {
          if (count_resp(n, -1) == 0) {
            for (int i = 0; i < n->nb_neighbors; i++) cd_send(cd, k, n->neighbors[i], MSG_VERDICT, n->phase_tag, +1);
            n->state = ST_FINISHED;
          } else {
            n->phase_tag += 1;
            for (int i = 0; i < n->nb_neighbors; i++) cd_send(cd, k, n->neighbors[i], MSG_VERDICT, n->phase_tag, -1);
            initialize_state(n);
          }
        }
      } else if (all_newer(n)) n->pp_end = 1;
    } else if (!n->response_sent) {
      int neg = count_resp(n, -1) > 0;
      if (!n->local_cv || neg) { /* :173 pointer compare again */
        for (int i = 0; i < n->nb_neighbors; i++)
          if (!n->recvd_pcv[i]) { cd_send(cd, k, n->neighbors[i], MSG_RESPONSE, n->phase_tag, -1); break; }
        n->response_sent = 1;
      } else if (n->pp_end) {
        if (count_resp(n, 0) == 1) {
          int asking = -1;
          for (int i = 0; i < n->nb_neighbors; i++) if (n->responses[i] == 0) { asking = n->neighbors[i]; break; }
          int v = (count_resp(n, +1) == n->nb_neighbors - 1) ? +1 : -1;
          cd_send(cd, k, asking, MSG_RESPONSE, n->phase_tag, v);
          n->response_sent = 1;
        }
      } else if (all_newer(n)) n->pp_end = 1;
    }
  }
  /* ---- receive_partial_CV :314-370 ---- */
  for (int sl = 0; sl < n->nb_neighbors; sl++) {
    cd_msg *mm = &n->inbox[sl][MSG_PARTIAL_CV];
    if (!mm->valid) continue;
    mm->valid = 0;
    if (mm->a == n->phase_tag) {
      n->recvd_pcv[sl] = 1;
      n->nb_not_recvd -= 1;
      int src = n->neighbors[sl];
      int leader = src > k ? src : k; /* choose_leader :500-508 */
      if (n->nb_not_recvd == 0 && n->partial_cv_sent && leader == k) {
        n->elected = 1;
        initialize_verification(n);
        for (int i = 0; i < n->nb_neighbors; i++) cd_send(cd, k, n->neighbors[i], MSG_VERIFICATION, n->phase_tag, 0);
        n->state = ST_VERIFICATION;
      }
    }
  }
  /* ---- receive_verification :373-411 ---- */
  for (int sl = 0; sl < n->nb_neighbors; sl++) {
    cd_msg *mm = &n->inbox[sl][MSG_VERIFICATION];
    if (!mm->valid) continue;
    mm->valid = 0;
    if (mm->a == n->phase_tag + 1) {
      initialize_verification(n);
      n->state = ST_VERIFICATION;
      for (int i = 0; i < n->nb_neighbors; i++)
        if (i != sl) cd_send(cd, k, n->neighbors[i], MSG_VERIFICATION, n->phase_tag, 0);
    }
  }
  /* ---- receive_response :414-448 ---- */
  for (int sl = 0; sl < n->nb_neighbors; sl++) {
    cd_msg *mm = &n->inbox[sl][MSG_RESPONSE];
    if (!mm->valid) continue;
    mm->valid = 0;
    if (mm->a == n->phase_tag) n->responses[sl] = mm->b;
  }
  /* ---- receive_verdict :451-498 ---- */
  for (int sl = 0; sl < n->nb_neighbors; sl++) {
    cd_msg *mm = &n->inbox[sl][MSG_VERDICT];
    if (!mm->valid) continue;
    mm->valid = 0;
    if (mm->b == +1) n->state = ST_FINISHED;
    else { initialize_state(n); n->phase_tag = mm->a; }
    for (int i = 0; i < n->nb_neighbors; i++)
      if (i != sl) cd_send(cd, k, n->neighbors[i], MSG_VERDICT, n->phase_tag, mm->b);
  }
}

/* ======================================================================= */
/* outer loops                                                             */
/* ======================================================================= */
typedef struct {
  int nb, off;                   /* rows owned, first global row */
  int32_t *rp, *ci; double *va;  /* strip A_K,: (global columns)      utils.c:247 */
  int32_t *drp, *dci; double *dva; /* A_KK (local columns)            utils.c:450 */
  int32_t *orp, *oci; double *ova; /* A_K,: minus A_KK (global cols): sum_J A_KJ */
  double *b, *rhs, *view;        /* b_K; local_right_side_vector; block's copy of the global iterate */
  double *r;                     /* scratch nb */
} blk;

static void blk_free(blk *B) {
  free(B->rp); free(B->ci); free(B->va); free(B->drp); free(B->dci); free(B->dva);
  free(B->orp); free(B->oci); free(B->ova); free(B->b); free(B->rhs); free(B->view); free(B->r);
}

static int blk_setup(const orc_config *c, int K, blk *B, int64_t ntot) {
  int G = c->nblocks;
  B->nb = (int)(ntot / G);
  B->off = K * B->nb;
  int64_t nnz = (c->dim == 2) ? orc_poisson2d_nnz(c->m, c->n, K, G) : orc_poisson3d_nnz(c->m, c->n, c->p, K, G);
  B->rp = (int32_t *)malloc(sizeof(int32_t) * (B->nb + 1));
  B->ci = (int32_t *)malloc(sizeof(int32_t) * nnz);
  B->va = (double *)malloc(sizeof(double) * nnz);
  if (c->dim == 2) orc_poisson2d(c->m, c->n, K, G, B->rp, B->ci, B->va);
  else orc_poisson3d(c->m, c->n, c->p, K, G, B->rp, B->ci, B->va);
  int64_t dn = orc_submatrix_nnz(B->nb, B->rp, B->ci, B->off, B->off + B->nb);
  B->drp = (int32_t *)malloc(sizeof(int32_t) * (B->nb + 1));
  B->dci = (int32_t *)malloc(sizeof(int32_t) * (dn + 1));
  B->dva = (double *)malloc(sizeof(double) * (dn + 1));
  orc_submatrix(B->nb, B->rp, B->ci, B->va, B->off, B->off + B->nb, B->drp, B->dci, B->dva);
  int64_t on = nnz - dn, q = 0;
  B->orp = (int32_t *)malloc(sizeof(int32_t) * (B->nb + 1));
  B->oci = (int32_t *)malloc(sizeof(int32_t) * (on + 1));
  B->ova = (double *)malloc(sizeof(double) * (on + 1));
  B->orp[0] = 0;
  for (int r = 0; r < B->nb; r++) {
    for (int k = B->rp[r]; k < B->rp[r + 1]; k++)
      if (B->ci[k] < B->off || B->ci[k] >= B->off + B->nb) { B->oci[q] = B->ci[k]; B->ova[q++] = B->va[k]; }
    B->orp[r + 1] = (int32_t)q;
  }
  B->b = (double *)malloc(sizeof(double) * B->nb);
  B->rhs = (double *)malloc(sizeof(double) * B->nb);
  B->r = (double *)malloc(sizeof(double) * B->nb);
  B->view = (double *)calloc(ntot, sizeof(double)); /* PETSc Vecs are zero-initialised */
  /* utils.c:623-626: b_K = A_K,: * u, u = 1 */
  double *ones = (double *)malloc(sizeof(double) * ntot);
  for (int64_t i = 0; i < ntot; i++) ones[i] = 1.0;
  orc_spmv(B->nb, B->rp, B->ci, B->va, ones, B->b);
  free(ones);
  /* sensitivity experiments only: ORC_PERTURB_B=eps multiplies b_i by (1 + eps sin(i)) */
  {
    const char *pe = getenv("ORC_PERTURB_B");
    if (pe && *pe) {
      const double eps = atof(pe);
      for (int i = 0; i < B->nb; i++) B->b[i] *= (1.0 + eps * sin((double)(B->off + i)));
    }
  }
  return 0;
}

/* utils.c:943-948 updateLocalRHS: rhs_K = b_K - A_KJ x_J */
static void update_rhs(blk *B) { orc_residual(B->nb, B->orp, B->oci, B->ova, B->b, B->view, B->rhs); }
/* utils.c:950-970 inner_solver: UIR norm, nonzero guess */
static int inner_solve(const orc_config *c, blk *B) {
  orc_ksp_opts o = c->inner;
  o.initial_rtol = 1; o.guess_nonzero = 1;
  int its = 0, reason = 0; double rn = 0;
  orc_gmres(B->nb, B->drp, B->dci, B->dva, B->rhs, B->view + B->off, &o, &its, &reason, &rn, NULL, 0);
  return its;
}
/* ||rhs_K - A_KK x_K|| */
static double local_resid(blk *B) {
  orc_residual(B->nb, B->drp, B->dci, B->dva, B->rhs, B->view + B->off, B->r);
  return orc_norm2(B->nb, B->r);
}
/* comm.c:126-141 comm_sync_send_and_receive, generalised: every block learns every other block's slice */
static void exchange_all(blk *B, int G) {
  for (int K = 0; K < G; K++)
    for (int J = 0; J < G; J++)
      if (J != K) memcpy(B[K].view + B[J].off, B[J].view + B[J].off, sizeof(double) * B[J].nb);
}
/* utils.c:575-595 computeFinalResidualNorm over all blocks */
static double global_resid(blk *B, int G) {
  double acc = 0.0;
  for (int K = 0; K < G; K++) {
    double ln = orc_block_residual_norm(B[K].nb, B[K].rp, B[K].ci, B[K].va, B[K].b, B[K].view);
    acc += ln * ln;
  }
  return sqrt(acc);
}
static int ls_solve(const orc_config *c, int64_t nrows, int s, const double *R, int64_t ld, const double *b, double *alpha,
                    double *rnorm) {
  if (c->outer.type == ORC_OUTER_LSQR) {
    int its, reason;
    return orc_lsqr(nrows, s, R, ld, b, alpha, &c->outer, &its, &reason, rnorm);
  }
  return orc_lstsq_qr(nrows, s, R, ld, b, alpha, rnorm);
}

/* Exact-LS path only (not the reference's LSQR path): replace the basis [x^1 .. x^s] by [x^1, x^2-x^1, .., x^s-x^(s-1)].
 * Same span, hence the same minimiser in exact arithmetic, but the successive corrections are far less collinear than
 * the iterates themselves, which removes most of the cancellation from alpha and from x = S alpha. */
static void diff_basis(double *S, int64_t rows, int s, int64_t ld) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < rows; i++) {
    double prev = S[i];
    for (int t = 1; t < s; t++) {
      double cur = S[(size_t)t * ld + i];
      S[(size_t)t * ld + i] = cur - prev;
      prev = cur;
    }
  }
}

static void push_hist(orc_result *res, double v) {
  if (res->hist_len < 4096) res->hist[res->hist_len++] = v;
}

static int solve_gmres_standalone(const orc_config *c, orc_result *res, double *x_out) {
  /* gmres_solution/gmres_solution.c:50-85 */
  if (c->dim != 2 || c->m != c->n) return 2;
  int64_t n = (int64_t)c->m * c->n;
  int64_t nnz = orc_poisson2d_nnz(c->m, c->n, 0, 1);
  int32_t *rp = (int32_t *)malloc(sizeof(int32_t) * (n + 1)), *ci = (int32_t *)malloc(sizeof(int32_t) * nnz);
  double *va = (double *)malloc(sizeof(double) * nnz);
  orc_poisson2d_complete(c->m, c->n, rp, ci, va);
  double *u = (double *)malloc(sizeof(double) * n), *b = (double *)malloc(sizeof(double) * n), *x = (double *)calloc(n, sizeof(double));
  for (int64_t i = 0; i < n; i++) u[i] = 1.0;
  orc_spmv((int)n, rp, ci, va, u, b);
  orc_ksp_opts o = c->inner;
  o.guess_nonzero = 0;
  res->norm0 = orc_norm2(n, b);
  double t0 = now_s();
  orc_gmres((int)n, rp, ci, va, b, x, &o, &res->gmres_its, &res->gmres_reason, &res->gmres_rnorm, res->hist, 4096);
  res->elapsed_s = now_s() - t0;
  res->last_norm = res->gmres_rnorm;
  res->outer_its = res->gmres_its;
  res->final_residual = orc_block_residual_norm((int)n, rp, ci, va, b, x);
  double e = 0.0;
  for (int64_t i = 0; i < n; i++) e += (x[i] - 1.0) * (x[i] - 1.0);
  res->error = sqrt(e);
  if (x_out) memcpy(x_out, x, sizeof(double) * n);
  free(rp); free(ci); free(va); free(u); free(b); free(x);
  return 0;
}

API int orc_solve(const orc_config *c, orc_result *res, double *x_out) {
  memset(res, 0, sizeof(*res));
#ifdef _OPENMP
  if (c->nthreads > 0) omp_set_num_threads(c->nthreads);
#endif
  if (c->alg == ORC_ALG_GMRES) return solve_gmres_standalone(c, res, x_out);
  const int G = c->nblocks, s = c->s;
  if (G < 1 || G > 16) return 1;
  int64_t ntot = (c->dim == 2) ? (int64_t)c->m * c->n : (int64_t)c->m * c->n * c->p;
  if (ntot % G) return 1;
  const double atol = 1e-100; /* hard-coded absolute_tolerance, e.g. …-global.c:34 */
  blk *B = (blk *)calloc(G, sizeof(blk));
  for (int K = 0; K < G; K++) blk_setup(c, K, &B[K], ntot);
  const int nb = B[0].nb;
  res->norm0 = global_resid(B, G); /* x = 0 => ||b|| */
  const double thr_global = fmax(atol, c->rtol * res->norm0);
  const double thr_local = fmax(atol, (c->rtol / sqrt((double)G)) * 1.0 * res->norm0); /* …-semi-local.c:330 (sqrt(2) at G = 2) */
  const int max_outer = c->max_outer > 0 ? c->max_outer : 100000;
  const int is_async = (c->alg >= ORC_ALG_AM);
  const int minim = (c->alg == ORC_ALG_SMSM_GLOBAL || c->alg == ORC_ALG_AMAM_GLOBAL) ? 1
                  : (c->alg == ORC_ALG_SMSM_SEMI_LOCAL || c->alg == ORC_ALG_AMAM_SEMI_LOCAL) ? 2
                  : (c->alg == ORC_ALG_SMSM_LOCAL || c->alg == ORC_ALG_AMAM_LOCAL) ? 3 : 0;
  double *S = NULL, *R = NULL, *alpha = (double *)calloc(s > 0 ? s : 1, sizeof(double));
  double *bglob = NULL, *xmin = NULL;
  int sig[16] = {0};
  int rc = 0;
  const double t_start = now_s();

  if (!is_async) {
    if (minim == 1 || minim == 2) {
      S = (double *)calloc((size_t)ntot * s, sizeof(double));
      R = (double *)calloc((size_t)ntot * s, sizeof(double));
      xmin = (double *)calloc(ntot, sizeof(double));
      bglob = (double *)malloc(sizeof(double) * ntot);
      for (int K = 0; K < G; K++) memcpy(bglob + B[K].off, B[K].b, sizeof(double) * nb);
    } else if (minim == 3) {
      S = (double *)calloc((size_t)nb * s * G, sizeof(double)); /* per block nb x s */
      R = (double *)calloc((size_t)nb * s * G, sizeof(double));
    }
    if (minim == 0) for (int K = 0; K < G; K++) update_rhs(&B[K]); /* …multisplitting.c:164 */
    int done = 0;
    while (!done && res->outer_its < max_outer && !(c->max_seconds > 0.0 && now_s() - t_start >= c->max_seconds)) {
      if (minim == 0) {
        /* …multisplitting.c:170-206 */
        for (int K = 0; K < G; K++) { int it = inner_solve(c, &B[K]); if (K == 0) res->inner_its_total += it; }
        exchange_all(B, G);
        double acc = 0.0;
        for (int K = 0; K < G; K++) { update_rhs(&B[K]); double ln = local_resid(&B[K]); acc += ln * ln; }
        double norm = sqrt(acc);
        res->last_norm = norm; push_hist(res, norm);
        if (norm <= thr_global) done = 1;
        if (res->outer_its < 256) res->t_outer[res->outer_its] = now_s() - t_start;
        res->outer_its++;
        continue;
      }
      /* build S: …-global.c:297-315, …-semi-local.c:287-309, …-local.c:232-246 */
      for (int t = 0; t < s; t++) {
        for (int K = 0; K < G; K++) update_rhs(&B[K]);
        for (int K = 0; K < G; K++) { int it = inner_solve(c, &B[K]); if (K == 0) res->inner_its_total += it; }
        exchange_all(B, G);
        if (minim == 3) for (int K = 0; K < G; K++) memcpy(S + ((size_t)K * s + t) * nb, B[K].view + B[K].off, sizeof(double) * nb);
        else memcpy(S + (size_t)t * ntot, B[0].view, sizeof(double) * ntot); /* all views agree after a synchronous exchange */
      }
      if (c->outer.type == ORC_OUTER_QR) {
        if (minim == 3) for (int K = 0; K < G; K++) diff_basis(S + (size_t)K * s * nb, nb, s, nb);
        else diff_basis(S, ntot, s, ntot);
      }
      if (minim == 1) {
        /* …-global.c:325-354: R = A S (block rows), complete R everywhere, LS on (R, b), x = S alpha */
        for (int t = 0; t < s; t++)
          for (int K = 0; K < G; K++) orc_spmv(nb, B[K].rp, B[K].ci, B[K].va, S + (size_t)t * ntot, R + (size_t)t * ntot + B[K].off);
        double norm = 0.0;
        rc = ls_solve(c, ntot, s, R, ntot, bglob, alpha, &norm);
        if (rc) break;
        dense_mv(ntot, s, S, ntot, alpha, xmin);
        for (int K = 0; K < G; K++) memcpy(B[K].view, xmin, sizeof(double) * ntot);
        res->last_norm = norm; push_hist(res, norm);
        if (norm <= thr_global) done = 1;
      } else if (minim == 2) {
        /* …-semi-local.c:319-341 */
        int all = 1;
        double worst = 0.0;
        for (int K = 0; K < G; K++) {
          for (int t = 0; t < s; t++) orc_spmv(nb, B[K].rp, B[K].ci, B[K].va, S + (size_t)t * ntot, R + (size_t)t * nb);
          double dummy;
          rc = ls_solve(c, nb, s, R, nb, B[K].b, alpha, &dummy);
          if (rc) break;
          double ln = local_resid(&B[K]); /* pre-minimisation x_K against the stale rhs_K (:326) */
          if (ln > worst) worst = ln;
          if (ln <= thr_local) sig[K] = 1; /* sticky send_signal (:330-333) */
          dense_mv(ntot, s, S, ntot, alpha, xmin);
          memcpy(B[K].view, xmin, sizeof(double) * ntot); /* :335-338 */
          all &= sig[K];
        }
        if (rc) break;
        res->last_norm = worst; push_hist(res, worst);
        if (all) done = 1;
      } else {
        /* …-local.c:256-274 */
        int all = 1;
        double worst = 0.0;
        for (int K = 0; K < G; K++) {
          double *SK = S + (size_t)K * s * nb, *RK = R + (size_t)K * s * nb;
          for (int t = 0; t < s; t++) orc_spmv(nb, B[K].drp, B[K].dci, B[K].dva, SK + (size_t)t * nb, RK + (size_t)t * nb);
          update_rhs(&B[K]);
          double dummy;
          rc = ls_solve(c, nb, s, RK, nb, B[K].rhs, alpha, &dummy);
          if (rc) break;
          dense_mv(nb, s, SK, nb, alpha, B[K].view + B[K].off);
          double ln = local_resid(&B[K]);
          if (ln > worst) worst = ln;
          if (ln <= thr_local) sig[K] = 1;
          all &= sig[K];
        }
        if (rc) break;
        res->last_norm = worst; push_hist(res, worst);
        if (all) done = 1;
      }
      if (res->outer_its < 256) res->t_outer[res->outer_its] = now_s() - t_start;
      res->outer_its++;
    }
  } else {
    /* ---------------- asynchronous variants, simulated with a deterministic schedule ---------------- */
    /* …multisplitting_prime.c:321-393 and …-{global,semi-local,local}_prime.c loops.  A tick-based
     * simulation: block K executes one full outer step at tick t iff t % period[K] == 0; a sent message
     * is visible to the destination's next probe (comm.c:455-554, newest message wins). */
    orc_cd *cd = orc_cd_create(G);
    typedef struct { int valid, tag, iter; double *x; } dmsg;
    dmsg *mail = (dmsg *)calloc((size_t)G * G, sizeof(dmsg)); /* mail[dst*G + src] */
    for (int i = 0; i < G * G; i++) mail[i].x = (double *)malloc(sizeof(double) * nb);
    int iters[16] = {0}, inner_outer[16] = {0}, state_seen[16] = {0};
    double **SS = NULL, **RR = NULL, **Rpub = NULL;
    if (minim) {
      SS = (double **)calloc(G, sizeof(double *)); RR = (double **)calloc(G, sizeof(double *)); Rpub = (double **)calloc(G, sizeof(double *));
      for (int K = 0; K < G; K++) {
        size_t rows = (minim == 3) ? (size_t)nb : (size_t)ntot;
        SS[K] = (double *)calloc(rows * s, sizeof(double));
        RR[K] = (double *)calloc(((minim == 1) ? (size_t)ntot : (size_t)nb) * s, sizeof(double));
        Rpub[K] = (double *)calloc((size_t)nb * s, sizeof(double)); /* newest published slab of block K */
      }
      xmin = (double *)calloc(ntot, sizeof(double));
      bglob = (double *)malloc(sizeof(double) * ntot);
      for (int K = 0; K < G; K++) memcpy(bglob + B[K].off, B[K].b, sizeof(double) * nb);
    }
    int rpub_valid[16] = {0};
    for (int K = 0; K < G; K++) update_rhs(&B[K]);
    int nfinished = 0;
    for (int64_t tick = 0; nfinished < G && tick < (int64_t)max_outer * 64; tick++) {
      for (int K = 0; K < G; K++) {
        int per = c->period[K] > 0 ? c->period[K] : 1;
        if (tick % per) continue;
        if (orc_cd_state(cd, K) == ST_FINISHED) continue;
        blk *Bk = &B[K];
        int nsteps = minim ? s : 1;
        for (int t = 0; t < nsteps; t++) {
          for (int pass = 0; pass < (minim ? 2 : 1); pass++) {
            if (pass == 1) {
              /* comm_async_test_and_send_prime: publish (PhaseTag, iteration, x_K) to the neighbours */
              for (int J = 0; J < G; J++) if (J == K - 1 || J == K + 1) {
                dmsg *mm = &mail[J * G + K];
                mm->valid = 1; mm->tag = orc_cd_phase_tag(cd, K); mm->iter = minim ? inner_outer[K] : iters[K];
                memcpy(mm->x, Bk->view + Bk->off, sizeof(double) * nb);
              }
            }
            /* comm_async_probe_and_receive_prime */
            for (int J = 0; J < G; J++) if (J == K - 1 || J == K + 1) {
              dmsg *mm = &mail[K * G + J];
              if (!mm->valid) continue;
              mm->valid = 0;
              /* the handler sees the state broadcast at the end of the previous outer iteration */
              int saved = cd->nd[K].state; cd->nd[K].state = state_seen[K];
              int cp = orc_cd_data_arrival(cd, K, J, mm->tag, mm->iter);
              cd->nd[K].state = saved;
              if (cp) memcpy(Bk->view + B[J].off, mm->x, sizeof(double) * nb);
            }
            if (pass == 0) {
              update_rhs(Bk);
              int it = inner_solve(c, Bk);
              if (K == 0) res->inner_its_total += it;
              if (!minim) {
                for (int J = 0; J < G; J++) if (J == K - 1 || J == K + 1) {
                  dmsg *mm = &mail[J * G + K];
                  mm->valid = 1; mm->tag = orc_cd_phase_tag(cd, K); mm->iter = iters[K];
                  memcpy(mm->x, Bk->view + Bk->off, sizeof(double) * nb);
                }
              }
            }
          }
          if (minim == 3) memcpy(SS[K] + (size_t)t * nb, Bk->view + Bk->off, sizeof(double) * nb);
          else if (minim) memcpy(SS[K] + (size_t)t * ntot, Bk->view, sizeof(double) * ntot);
          inner_outer[K]++;
        }
        if (minim && c->outer.type == ORC_OUTER_QR) diff_basis(SS[K], (minim == 3) ? nb : ntot, s, (minim == 3) ? nb : ntot);
        double ln;
        if (minim == 0) {
          ln = local_resid(Bk);
        } else if (minim == 1) {
          /* …-global_prime.c:421-444: own slab fresh, other slabs = newest published by their owners */
          for (int t = 0; t < s; t++) orc_spmv(nb, Bk->rp, Bk->ci, Bk->va, SS[K] + (size_t)t * ntot, RR[K] + (size_t)t * ntot + Bk->off);
          for (int t = 0; t < s; t++) memcpy(Rpub[K] + (size_t)t * nb, RR[K] + (size_t)t * ntot + Bk->off, sizeof(double) * nb);
          rpub_valid[K] = 1;
          for (int J = 0; J < G; J++) if (J != K && rpub_valid[J])
            for (int t = 0; t < s; t++) memcpy(RR[K] + (size_t)t * ntot + B[J].off, Rpub[J] + (size_t)t * nb, sizeof(double) * nb);
          double dummy;
          rc = ls_solve(c, ntot, s, RR[K], ntot, bglob, alpha, &dummy);
          if (rc) break;
          dense_mv(ntot, s, SS[K], ntot, alpha, xmin);
          ln = orc_block_residual_norm(nb, Bk->rp, Bk->ci, Bk->va, Bk->b, xmin);
          memcpy(Bk->view, xmin, sizeof(double) * ntot);
        } else if (minim == 2) {
          /* …-semi-local_prime.c:391-393 */
          for (int t = 0; t < s; t++) orc_spmv(nb, Bk->rp, Bk->ci, Bk->va, SS[K] + (size_t)t * ntot, RR[K] + (size_t)t * nb);
          double dummy;
          rc = ls_solve(c, nb, s, RR[K], nb, Bk->b, alpha, &dummy);
          if (rc) break;
          dense_mv(ntot, s, SS[K], ntot, alpha, xmin);
          ln = orc_block_residual_norm(nb, Bk->rp, Bk->ci, Bk->va, Bk->b, xmin);
          memcpy(Bk->view, xmin, sizeof(double) * ntot);
        } else {
          /* …-local_prime.c:400-404 */
          for (int t = 0; t < s; t++) orc_spmv(nb, Bk->drp, Bk->dci, Bk->dva, SS[K] + (size_t)t * nb, RR[K] + (size_t)t * nb);
          update_rhs(Bk);
          double dummy;
          rc = ls_solve(c, nb, s, RR[K], nb, Bk->rhs, alpha, &dummy);
          if (rc) break;
          dense_mv(nb, s, SS[K], nb, alpha, Bk->view + Bk->off);
          ln = local_resid(Bk);
        }
        orc_cd_step(cd, K, ln <= thr_local);
        state_seen[K] = orc_cd_state(cd, K);
        iters[K]++;
        if (K == 0) push_hist(res, ln);
        res->last_norm = ln;
        if (orc_cd_state(cd, K) == ST_FINISHED) nfinished++;
      }
      if (rc) break;
    }
    for (int K = 0; K < G; K++) { res->outer_its_block[K] = iters[K]; if (iters[K] > res->outer_its) res->outer_its = iters[K]; }
    if (nfinished < G && !rc) rc = 4; /* schedule cap reached */
    for (int i = 0; i < G * G; i++) free(mail[i].x);
    free(mail);
    if (minim) { for (int K = 0; K < G; K++) { free(SS[K]); free(RR[K]); free(Rpub[K]); } free(SS); free(RR); free(Rpub); }
    orc_cd_destroy(cd);
  }

  res->elapsed_s = now_s() - t_start;
  /* closing synchronous exchange + true residual + error (e.g. …-global.c:376-388) */
  exchange_all(B, G);
  res->final_residual = global_resid(B, G);
  {
    double e = 0.0;
    for (int K = 0; K < G; K++)
      for (int i = 0; i < nb; i++) { double d = B[K].view[B[K].off + i] - 1.0; e += d * d; }
    res->error = sqrt(e);
  }
  if (x_out) for (int K = 0; K < G; K++) memcpy(x_out + B[K].off, B[K].view + B[K].off, sizeof(double) * nb);
  for (int K = 0; K < G; K++) blk_free(&B[K]);
  orc_gmres_release_workspace();
  free(B); free(S); free(R); free(alpha); free(bglob); free(xmin);
  return rc;
}

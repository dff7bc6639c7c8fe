/*
 * msplit_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the reference's multisplitting solve path
 * (craftman22/medane_tchakorom_ufc_thesis_repository, C on PETSc 3.22.1 + MPICH).
 * Only tests/ (and the checker scripts the tests run or that produce their
 * evidence: tools/mgpu_check.py, tools/parity_margins.py, tools/sensitivity_*.py,
 * tools/history_1024_vs_oracle.py),
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library — always as the checker or the CPU baseline, never as the
 * thing measured or shipped.  The product path (libmsplit.so and the package
 * medane_tchakorom_ufc_thesis_repository_b200/) never links, imports or calls it
 * (tests/test_cabi_and_host.py::test_product_never_imports_oracle).
 *
 * PARITY STATUS
 *   pinned   : CSR assembly (poisson2D/3D), block partition arithmetic and the
 *              residual-norm reduction — against the reference's four Unity
 *              known-answer tests (src/tests/utils_test.c:38-64, :66-170,
 *              :172-221, :225-228 with inputs :285-316).
 *   unpinned by the reference: GMRES / LSQR / minimisation / outer loops / convergence
 *              detection.  The arithmetic lives in PETSc 3.22.1 (un-vendored dependency,
 *              not on disk, reference unbuildable here: every source includes
 *              <petscts.h>); the reference's own tests hold no vector for it.  These
 *              parts restate PETSc's published algorithm (gmres.c, borthog2.c,
 *              iterativ.c, lsqr.c) as summarised in SURVEY.md Appendix A: "parity
 *              unpinned" against the reference.
 *   cross-checked (round 2): the capped restarted GMRES, the MSM and SMSM-global outer
 *              loops and LSQR agree with an independent numpy/scipy implementation that
 *              shares no code and no algorithmic shortcut with this file
 *              (tests/independent_reference.py, tests/test_independent_pin.py: iteration
 *              counts equal or +-1, iterates to 1e-8..1e-10, scipy.sparse.linalg.lsqr to
 *              1e-7 / phibar to 1e-10, the three points VERDICT r01 measured).  The
 *              asynchronous detection state machine has no second implementation.
 */
#ifndef MSPLIT_ORACLE_H
#define MSPLIT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* KSPConvergedReason values (PETSc 3.22 petscksp.h) */
enum {
  ORC_CONVERGED_ITERATING = 0,
  ORC_CONVERGED_RTOL_NORMAL = 1,
  ORC_CONVERGED_RTOL = 2,
  ORC_CONVERGED_ATOL = 3,
  ORC_CONVERGED_ITS = 4,
  ORC_CONVERGED_HAPPY_BREAKDOWN = 7,
  ORC_CONVERGED_ATOL_NORMAL = 9,
  ORC_DIVERGED_NULL = -2,
  ORC_DIVERGED_ITS = -3,
  ORC_DIVERGED_DTOL = -4,
  ORC_DIVERGED_BREAKDOWN = -5,
  ORC_DIVERGED_NANORINF = -9
};

/* algorithms (iSolve --alg names, iSolve:30-57) */
enum {
  ORC_ALG_SM = 0,              /* synchronous multisplitting (MSM) */
  ORC_ALG_SMSM_GLOBAL = 1,
  ORC_ALG_SMSM_SEMI_LOCAL = 2,
  ORC_ALG_SMSM_LOCAL = 3,
  ORC_ALG_GMRES = 4,           /* src/gmres_solution */
  ORC_ALG_AM = 5,              /* asynchronous multisplitting */
  ORC_ALG_AMAM_GLOBAL = 6,
  ORC_ALG_AMAM_SEMI_LOCAL = 7,
  ORC_ALG_AMAM_LOCAL = 8
};

enum { ORC_OUTER_LSQR = 0, ORC_OUTER_QR = 1 };

typedef struct {
  int restart;       /* -ksp_gmres_restart (30) */
  int max_it;        /* -ksp_max_it (10000) */
  double rtol;       /* -ksp_rtol (1e-5) */
  double abstol;     /* -ksp_atol (1e-50) */
  double divtol;     /* -ksp_divtol (1e4) */
  int initial_rtol;  /* KSPConvergedDefaultSetUIRNorm / -ksp_converged_use_initial_residual_norm */
  int guess_nonzero; /* KSPSetInitialGuessNonzero */
  int cgs_refine;    /* 0 REFINE_NEVER (default), 1 IFNEEDED, 2 ALWAYS */
  int mgs;           /* -ksp_gmres_modifiedgramschmidt */
  int min_it;        /* 0 */
} orc_ksp_opts;

typedef struct {
  int type;          /* ORC_OUTER_LSQR | ORC_OUTER_QR */
  int max_it;        /* LSQR iterations (scripts: 40..200) */
  double rtol;       /* LSQR rtol (scripts: 1e-14..1e-50) */
  double abstol;
  int lsqr_default_test; /* -ksp_convergence_test default (else KSPLSQRConvergedDefault) */
} orc_outer_opts;

typedef struct {
  int alg;
  int dim;           /* 2 or 3 */
  int m, n, p;       /* 2-D: m grid lines x n grid columns; 3-D: m lines (fastest index), n columns, p depth */
  int nblocks;
  int s;
  double rtol;
  orc_ksp_opts inner;
  orc_outer_opts outer;
  int max_outer;     /* safety cap on outer iterations (reference: none) */
  /* async schedule (oracle simulation only): block K runs one step at tick t iff t % period[K] == 0 */
  int period[16];
  int nthreads;      /* OpenMP threads for elementwise loops (0 = leave) */
  double max_seconds; /* synchronous drivers: leave the outer loop once it has run this long (0 = no cap; bench.py's CPU arm) */
} orc_config;

typedef struct {
  int outer_its;            /* number_of_iterations of the reference drivers */
  int outer_its_block[16];  /* async: per block */
  int64_t inner_its_total;  /* sum of KSPGetIterationNumber over all inner solves of block 0 */
  double norm0;             /* global_norm_0 = ||b|| */
  double last_norm;         /* the stopping quantity at exit */
  double final_residual;    /* computeFinalResidualNorm after the closing exchange */
  double error;             /* ||x - 1||_2 (computeError) */
  int hist_len;
  double hist[4096];        /* stopping quantity per outer iteration */
  int gmres_its;            /* alg GMRES: KSP its */
  int gmres_reason;
  double gmres_rnorm;
  double elapsed_s;         /* wall-clock of the outer loop only (the reference's MPI_Wtime region) */
  double t_outer[256];      /* synchronous drivers: seconds since the start of the outer loop at the end of outer iteration i (i < 256) */
} orc_result;

/* ---- assembly (bit-exact gate) ---- */
int64_t orc_poisson2d_nnz(int m, int n, int block, int nblocks);
int orc_poisson2d(int m, int n, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val);
int orc_poisson2d_complete(int m, int n, int32_t *rowptr, int32_t *colidx, double *val);
int64_t orc_poisson3d_nnz(int nx, int ny, int nz, int block, int nblocks);
int orc_poisson3d(int nx, int ny, int nz, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val);
int64_t orc_submatrix_nnz(int nrows, const int32_t *rowptr, const int32_t *colidx, int col_lo, int col_hi);
int orc_submatrix(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val, int col_lo, int col_hi,
                  int32_t *out_rowptr, int32_t *out_colidx, double *out_val);
int orc_dimension_related(int nprocs, int npb, int rank, int m, int n, int *njacobi_blocks, int *rank_jacobi_block,
                          int *proc_local_rank, int *n_mesh_points, int *jacobi_block_size);

/* ---- vector / matrix kernels ---- */
void orc_spmv(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val, const double *x, double *y);
void orc_residual(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val, const double *b,
                  const double *x, double *r);
double orc_dot(int64_t n, const double *a, const double *b);
double orc_norm2(int64_t n, const double *a);
double orc_block_residual_norm(int nrows, const int32_t *rowptr, const int32_t *colidx, const double *val,
                               const double *b, const double *x);

/* ---- Krylov solvers ---- */
void orc_ksp_defaults(orc_ksp_opts *o);
void orc_gmres_release_workspace(void); /* the Krylov basis is kept between orc_gmres calls of the same size */
int orc_gmres(int n, const int32_t *rowptr, const int32_t *colidx, const double *val, const double *b, double *x,
              const orc_ksp_opts *o, int *its, int *reason, double *rnorm, double *hist, int hist_cap);
int orc_lsqr(int64_t nrows, int s, const double *R, int64_t ldr, const double *b, double *alpha, const orc_outer_opts *o,
             int *its, int *reason, double *rnorm);
int orc_lstsq_qr(int64_t nrows, int s, const double *R, int64_t ldr, const double *b, double *alpha, double *rnorm);

/* ---- whole outer loops ---- */
int orc_solve(const orc_config *cfg, orc_result *res, double *x_out /* n_tot or NULL */);

/* ---- asynchronous convergence detection (conv_detection_prime.c) ---- */
typedef struct orc_cd orc_cd;
orc_cd *orc_cd_create(int nblocks);
void orc_cd_destroy(orc_cd *);
/* one root-only detection step of block k: conv-detection + 4 receive handlers, in the reference's order */
void orc_cd_step(orc_cd *, int k, int under_threshold);
/* data message arrival: returns 1 if the payload must be copied (receive_data_dependency) */
int orc_cd_data_arrival(orc_cd *, int k, int src, int src_tag, int src_iter);
int orc_cd_state(const orc_cd *, int k);
int orc_cd_phase_tag(const orc_cd *, int k);

#ifdef __cplusplus
}
#endif
#endif

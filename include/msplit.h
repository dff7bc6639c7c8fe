/*
 * msplit.h — C-ABI of libmsplit.so, the B200-native multisplitting solve path.
 *
 * Drop-in boundary for the hot path of craftman22/medane_tchakorom_ufc_thesis_repository
 * (reference paths below are relative to /root/reference/).  The reference has no FFI of
 * its own: its boundary is the C operator surface of include/utils.h + include/comm.h on
 * PETSc handles.  Each entry point here names the reference function it replaces.  PETSc
 * Mat/Vec/KSP handles become one opaque per-block engine (one block = one GPU); every
 * function returns int, 0 = success (PETSC_SUCCESS), non-zero = error with a message in
 * msp_last_error().  Plain pointers and sizes only; all pointers are HOST pointers unless
 * a name ends in _dev.  One host thread per engine; the engine owns its CUDA stream.
 */
#ifndef MSPLIT_H
#define MSPLIT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSP_VERSION 200 /* round 2: msp_problem.npb, msp_solve_opts.{max_seconds,detector,...}, msp_result.{hist_dropped,stop_reason} */
#define MSP_MAX_RESTART 64
#define MSP_MAX_S 32
#define MSP_MAX_BLOCKS 64

typedef struct msp_engine msp_engine;
typedef struct msp_group msp_group;

/* iSolve --alg names (iSolve:30-57; binaries makefile:120-148) */
enum {
  MSP_ALG_SM = 0,              /* src/synchronous-multisplitting/synchronous-multisplitting.c */
  MSP_ALG_SMSM_GLOBAL = 1,     /* …-synchronous-minimization-global.c */
  MSP_ALG_SMSM_SEMI_LOCAL = 2, /* …-synchronous-minimization-semi-local.c */
  MSP_ALG_SMSM_LOCAL = 3,      /* …-synchronous-minimization-local.c */
  MSP_ALG_GMRES = 4,           /* src/gmres_solution/gmres_solution.c */
  MSP_ALG_AM = 5,              /* src/asynchronous-multisplitting/asynchronous-multisplitting_prime.c */
  MSP_ALG_AMAM_GLOBAL = 6,     /* …-asynchronous-minimization-global_prime.c */
  MSP_ALG_AMAM_SEMI_LOCAL = 7, /* …-semi-local_prime.c */
  MSP_ALG_AMAM_LOCAL = 8       /* …-local_prime.c */
};

/* KSPConvergedReason (PETSc 3.22 petscksp.h) */
enum {
  MSP_CONVERGED_ITERATING = 0, MSP_CONVERGED_RTOL = 2, MSP_CONVERGED_ATOL = 3, MSP_CONVERGED_ITS = 4,
  MSP_CONVERGED_HAPPY_BREAKDOWN = 7, MSP_DIVERGED_NULL = -2, MSP_DIVERGED_ITS = -3, MSP_DIVERGED_DTOL = -4,
  MSP_DIVERGED_BREAKDOWN = -5, MSP_DIVERGED_NANORINF = -9
};

/* which matrix: the strip A_K,: (poisson2DMatrix), the diagonal block A_KK or the coupling part
 * sum_J A_KJ (divideSubDomainIntoBlockMatrices, utils.c:450-478) */
enum { MSP_MAT_STRIP = 0, MSP_MAT_DIAG = 1, MSP_MAT_OFFDIAG = 2 };

/* KSP options actually used by the reference command lines (running_bulk_test_local:72-310,
 * running_bulk_test_g5k:230-320); defaults = PETSc 3.22.1 (tmp/petscmpiexec_help:336-342,602-615) */
typedef struct {
  int restart;       /* -ksp_gmres_restart            30 */
  int max_it;        /* -ksp_max_it                   10000 */
  double rtol;       /* -ksp_rtol                     1e-5 */
  double abstol;     /* -ksp_atol                     1e-50 */
  double divtol;     /* -ksp_divtol                   1e4 */
  int initial_rtol;  /* -ksp_converged_use_initial_residual_norm / KSPConvergedDefaultSetUIRNorm */
  int guess_nonzero; /* KSPSetInitialGuessNonzero */
  int cgs_refine;    /* -ksp_gmres_cgs_refinement_type: 0 never, 1 ifneeded, 2 always */
  int mgs;           /* -ksp_gmres_modifiedgramschmidt */
  int min_it;        /* 0 */
} msp_ksp_opts;

typedef struct {
  int dim;     /* 2 or 3 */
  int m, n, p; /* 2-D: m grid lines x n grid columns (-m -n); 3-D: m lines (fastest), n columns, p depth */
  int block;   /* rank_jacobi_block */
  int nblocks; /* njacobi_blocks (reference: 2; generalised to a 1-D strip partition) */
  int s;       /* -s : minimisation basis size (0 if unused) */
  int max_restart; /* storage for the Krylov basis; >= every restart used later */
  int keep_csr;    /* keep the strip CSR on the device after setup (needed by msp_get_csr) */
  /* -npb: GPUs per Jacobi block (the reference's ranks per block, computeDimensionRelatedVariables utils.c:652-666).
   * 0 or 1: one GPU per block.  P > 1: `nblocks` counts the GPUs (strips), Jacobi blocks = nblocks / P, GPU r belongs to block
   * r / P; the block's inner GMRES is then distributed over its P GPUs (boundary layers of every Krylov vector between the
   * GPUs of a block, MDot and norm results summed over them).  Supported by SM, SMSM_GLOBAL and the stand-alone GMRES. */
  int npb;
} msp_problem;

typedef struct {
  int alg;
  int s;
  double rtol;          /* -rtol */
  msp_ksp_opts inner;   /* -inner{K}_ksp_* */
  int max_outer;        /* safety cap (reference: none); 0 = 1000000 */
  int record_history;
  int profile;          /* bracket every hot-kernel launch with CUDA events and fill the per-class t_, b_ and n_ fields */
  /* minimiser (outer_solver_norm_equation utils.c:1061-1103): 0 = exact least squares by TSQR (default),
   * 1 = PETSc-faithful LSQR on R with zero initial guess and the initial-residual-norm test, as every shipped
   * command line selects (-outer{K}_ksp_type lsqr -outer{K}_ksp_max_it 40..200 -outer{K}_ksp_rtol 1e-14..1e-50),
   * 2 = normal equations on the Gram matrix R'R solved by Cholesky (the reference's `outer_solver`, utils.c:972-996), s <= 8,
   * 3 = the same Gram system solved by PETSc's CG (`outer_solver` with -outer_ksp_type cg, config/default_run_variables), s <= 8,
   * 4 = CGNE on R without forming R'R (`outer_solver_cgne`, utils.c:1020-1043) */
  int outer_type;
  int outer_max_it;
  double outer_rtol, outer_abstol;
  /* async emulation in one process: block K runs a step at tick t iff t % period[K] == 0 */
  int period[MSP_MAX_BLOCKS];
  /* wall-clock cap of the outer loop in seconds (reference: the `timeout -k 5s 3600` wrapper of
   * running_bulk_test_local:3-7); 0 = none.  Synchronous variants agree on it collectively (one extra 1-double
   * allreduce per outer iteration, only when the cap is set), asynchronous blocks stop on their own clock. */
  double max_seconds;
  /* asynchronous termination: 0 = conv_detection_prime.c (verification phases; what the compiled *_prime drivers use),
   * 1 = the legacy counter-based detector of conv_detection.c (-min_convergence_count consecutive iterations under the
   * threshold, SEND_CV / CANCEL_CV / GLOBAL_CV messages, exit after globalCV has held for max_traversal_ms;
   * asynchronous-multisplitting.c.save:280-329) */
  int detector;
  int min_convergence_count;   /* MIN_CONVERGENCE_COUNT, config/default_run_variables: 4 (0 = 4) */
  double max_traversal_ms;     /* MAX_TRAVERSAL_TIME; the reference measures a ping between the two block roots (13.21 ms on its
                                  cluster); 0 = 0.5 ms, generous for NVLink peers */
} msp_solve_opts;

/* why the outer loop ended (msp_result.stop_reason) */
enum { MSP_STOP_CONVERGED = 0, MSP_STOP_MAX_OUTER = 1, MSP_STOP_MAX_SECONDS = 2 };

typedef struct {
  int outer_its;              /* number_of_iterations printed by utils.c:703-729 */
  int64_t inner_its_total;    /* sum of KSPGetIterationNumber of this block's inner solves */
  double norm0;               /* global_norm_0 */
  double last_norm;           /* stopping quantity at exit */
  double final_residual;      /* computeFinalResidualNorm after the closing exchange */
  double error;               /* computeError: ||x - 1||_2 */
  double elapsed_s;           /* device time of the outer loop (CUDA events), the reference's MPI_Wtime region */
  int gmres_its, gmres_reason;
  double gmres_rnorm;
  int hist_len;
  double hist[4096];
  int64_t kernel_launches;    /* kernels launched by this engine inside the timed region */
  /* per-kernel-class device time inside the timed region (ms) and launches, when profiling is on */
  double t_spmv_ms, t_mdot_ms, t_maxpy_ms, t_other_ms;
  double b_spmv, b_mdot, b_maxpy, b_other;   /* algorithmic bytes moved by each class (SURVEY.md §8d formulas) */
  int64_t n_spmv, n_mdot, n_maxpy, n_other;  /* launches per class */
  /* the reference's PetscLogStage split (…-global.c:81-89): host time spent in the inner solves ("I_Solver stage")
   * and in the exchange + minimisation + convergence test ("O_Solver stage") */
  double stage_inner_s, stage_outer_s;
  int64_t outer_solver_its;   /* LSQR iterations summed over the outer iterations (0 with TSQR) */
  int hist_dropped;           /* history entries that did not fit hist[4096] (the first 4096 are kept) */
  int stop_reason;            /* MSP_STOP_* */
} msp_result;

int msp_version(void);
const char *msp_last_error(void);
int msp_device_count(void);

/* ---- CSR assembly on the device, returned to host arrays (bit-exact gate) ----
 * replaces poisson2DMatrix utils.c:247-293, poisson2DMatrix_complete utils.c:383-445,
 * poisson3DMatrix utils.c:30-121 (+ MatAssemblyBegin/End).  rowptr[nb+1], colidx/val[nnz]. */
int64_t msp_poisson2d_nnz(int m, int n, int block, int nblocks);
int64_t msp_poisson3d_nnz(int nx, int ny, int nz, int block, int nblocks);
int msp_assemble_poisson2d(int device, int m, int n, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val);
int msp_assemble_poisson2d_complete(int device, int m, int n, int32_t *rowptr, int32_t *colidx, double *val);
int msp_assemble_poisson3d(int device, int nx, int ny, int nz, int block, int nblocks, int32_t *rowptr, int32_t *colidx, double *val);
/* computeDimensionRelatedVariables utils.c:652-666 */
int msp_dimension_related(int nprocs, int npb, int rank, int m, int n, int *njacobi_blocks, int *rank_jacobi_block,
                          int *proc_local_rank, int *n_mesh_points, int *jacobi_block_size);

/* ---- engine: one Jacobi block resident on one GPU ---- */
/* create = create_matrix_sparse + poissonXDMatrix + divideSubDomainIntoBlockMatrices + create_vector(s)
 * (…multisplitting.c:101-153); b_K = A_K,: * 1 (computeTheRightHandSideWithInitialGuess utils.c:623-626) */
int msp_create(const msp_problem *prob, int device, msp_engine **out);
int msp_destroy(msp_engine *e);
int msp_rows(const msp_engine *e);      /* jacobi_block_size */
int msp_halo_size(const msp_engine *e); /* one grid line (2-D) / plane (3-D) */
/* storage the hot SpMV reads: returns 2 = coded DIA (<= 8 diagonals, each one constant wherever present: one presence
 * byte per row, 1 + 16 bytes per row), 1 = DIA (<= 8 diagonals: values only, 8*width + 16 bytes per row),
 * 0 = slot-major ELL (values + indices, 12*width + 16 bytes per row); *width = diagonals / slots.
 * Environment: MSPLIT_NO_CDIA=1 keeps the plain DIA view, MSPLIT_NO_DIA=1 the ELL view. */
int msp_spmv_format(const msp_engine *e, int *width);
/* 1 when this engine runs each GMRES restart cycle (KSPGMRESCycle: the loop inside inner_solver utils.c:950-970) as ONE
 * persistent cooperative kernel with grid barriers instead of one kernel per phase — chosen for small blocks, where the
 * phases are launch-bound; the iterates are bit-identical either way.  Needs the coded-DIA stencil view, one GPU per
 * Jacobi block, classical Gram-Schmidt (any refinement type; modified Gram-Schmidt keeps one kernel per phase).  Environment: MSPLIT_COOP=0 never, MSPLIT_COOP=1 for every
 * eligible block, default: blocks of at most MSPLIT_COOP_MAX_ROWS rows. */
int msp_persistent_cycles(const msp_engine *e);
int64_t msp_mat_nnz(msp_engine *e, int which);
int msp_get_csr(msp_engine *e, int which, int32_t *rowptr, int32_t *colidx, double *val);
int msp_set_b(msp_engine *e, const double *b);
int msp_get_b(msp_engine *e, double *b);
int msp_set_x(msp_engine *e, const double *x);
int msp_get_x(msp_engine *e, double *x);
/* pipelined variants for callers that step the solver from PAGE-LOCKED host buffers: the uploads are only enqueued on the
 * engine's stream, the download leaves on a second stream from a snapshot of x, so the result of step k travels while step
 * k + 1 uploads and computes.  The host buffers must stay valid (and the output unread) until msp_copies_wait returns. */
int msp_set_b_async(msp_engine *e, const double *b_pinned);
int msp_set_x_async(msp_engine *e, const double *x_pinned);
int msp_get_x_async(msp_engine *e, double *x_pinned);
int msp_copies_wait(msp_engine *e);
/* neighbour boundary values this block currently holds (side 0 = block K-1, 1 = block K+1) */
int msp_set_halo(msp_engine *e, int side, const double *h);
int msp_get_halo(msp_engine *e, int side, double *h);
int msp_get_rhs(msp_engine *e, double *rhs);

/* ---- operator surface (include/utils.h) ---- */
int msp_update_local_rhs(msp_engine *e);                                   /* updateLocalRHS utils.c:943-948 */
int msp_inner_solve(msp_engine *e, const msp_ksp_opts *o, int *its, int *reason, double *rnorm); /* inner_solver utils.c:950-970; with npb > 1 collective over the GPUs of the block, like KSPSolve on comm_jacobi_block */
int msp_local_residual_norm(msp_engine *e, double *nrm);  /* ||rhs_K - A_KK x_K||  (…multisplitting.c:187-188) */
int msp_block_residual_norm(msp_engine *e, double *nrm);  /* ||b_K - A_K,: x||     (computeFinalResidualNorm utils.c:579-580) */
int msp_error_norm_sq(msp_engine *e, double *sq);         /* computeError utils.c:1045-1059, block part squared */
int msp_push_iterate(msp_engine *e, int t);               /* S[:,t] = x  (MatSetValuesLocal …-global.c:310-312) */
int msp_spmm_AS(msp_engine *e, int kind);                 /* MatMatMult(A,S,&R) …-global.c:326 / …-semi-local.c:319 / …-local.c:256 */
/* block-local part of the least-squares solve (TSQR leaf): QR of [R_K | rhs] -> (s+1)x(s+1) upper factor,
 * column-major, written to u_aug.  kind selects the right-hand side like the reference drivers do. */
int msp_minimize_local_qr(msp_engine *e, int kind, double *u_aug);
/* x = S alpha on own rows and on the stored neighbour boundaries (MatMult(S,alpha,x) utils.c:1075,1100) */
int msp_apply_alpha(msp_engine *e, int kind, const double *alpha);
/* small stacked least squares: nfac factors u_aug[(s+1)^2] -> alpha[s], ||b - R alpha|| (outer_solver_norm_equation utils.c:1061-1078) */
int msp_tsqr_combine(int s, int nfac, const double *u_aug_all, double *alpha, double *resnorm);

/* ---- the fine-grained exchange and minimiser a reference main() drives itself (include/comm.h, include/utils.h) ---- */
int msp_get_solution(msp_engine *e, double *x);            /* = msp_get_x: the block's slice of the iterate */
int msp_split_blocks(msp_engine *e, int which, int32_t *rowptr, int32_t *colidx, double *val); /* divideSubDomainIntoBlockMatrices utils.c:450-478 (= msp_get_csr) */
int msp_compute_rhs_ones(msp_engine *e);                   /* computeTheRightHandSideWithInitialGuess utils.c:623-650: b_K = A_K,: 1, rhs_K = b_K, halos 0 */
int msp_residual_norm(msp_engine *e, double *nrm);         /* computeFinalResidualNorm utils.c:575-595: sqrt(sum_K ||b_K - A_K,: x||^2), collective */
/* wire two engines of ONE process as neighbours (side 0: `neighbour` is block K-1, side 1: block K+1); msp_group_create and
 * msp_comm_connect do the same for a whole group / across processes */
int msp_connect_local(msp_engine *e, int side, msp_engine *neighbour);
/* comm_sync_send_and_receive comm.c:126-141: store my boundary layers into the neighbours' receive windows (P2P), wait until
 * every block has done so, copy the received layers into the private halos.  Collective: one caller per block (threads of a
 * group, or one process per GPU). */
int msp_exchange_sync(msp_engine *e);
/* comm_async_test_and_send_prime comm.c:531-554: store the layers, then release the header {seq, PhaseTag, iteration}.
 * comm_async_probe_and_receive_prime comm.c:455-529: take the newest header of each neighbour, let receive_data_dependency
 * (conv_detection_prime.c:603-633) decide, copy the layer; accepted[side] = 1 if a layer was taken.  Never block. */
int msp_async_reset(msp_engine *e);                        /* state of a new asynchronous run (…multisplitting_prime.c:280-315) */
int msp_exchange_async_publish(msp_engine *e, int iteration);
int msp_exchange_async_poll(msp_engine *e, int *accepted /* [2], may be null */);
/* outer_solver_norm_equation[_modify] utils.c:1061-1103 and the rest of the outer-solver menu (outer_solver utils.c:972-996,
 * outer_solver_cgne utils.c:1020-1043): least squares on R = A S (msp_spmm_AS) against b_K (rhs_K for LOCAL), then x = S alpha.
 * kind: MSP_ALG_SMSM_GLOBAL (collective: one problem over all blocks) | _SEMI_LOCAL | _LOCAL.
 * outer_type: 0 exact LS (TSQR), 1 LSQR, 2 normal equations + Cholesky, 3 CG on the normal equations, 4 CGNE. */
int msp_minimize(msp_engine *e, int kind, int outer_type, int outer_max_it, double outer_rtol, double *alpha, double *resnorm);

/* raw kernels on host data, for parity tests and micro-benchmarks */
int msp_op_spmv(msp_engine *e, int which, const double *x, const double *halo_lo, const double *halo_hi, double *y);
int msp_op_mdot(msp_engine *e, int nv, const double *V /* nv x nb */, const double *w, double *h);
int msp_op_maxpy(msp_engine *e, int nv, const double *V, const double *coef, double *w /* in/out */, double *norm);
/* device micro-benchmark of the hot kernels on this engine's resident data: returns average ms per launch.
 * op: 0 spmv(ELL), 1 mdot(nv), 2 maxpy+norm(nv), 3 spmm(s), 4 copy (STREAM), 5 spmv with the input scaled on the fly,
 *     6 Gram contraction of nv columns, 7 the same in panels of 8 columns (wide bases), 8 C := C T in panels (T = identity) */
int msp_bench_kernel(msp_engine *e, int op, int nv, int iters, int flush_l2, double *ms_avg);

/* standalone GMRES (gmres_solution.c:50-85): b = A 1, x0 = 0, one KSPSolve */
int msp_gmres_solve(msp_engine *e, const msp_ksp_opts *o, msp_result *res);

/* ---- multi-block ---- */
/* (a) all blocks in one process (one host thread per block; any mix of GPUs, peer access enabled when
 *     two blocks sit on different GPUs).  replaces comm_sync_send_and_receive comm.c:126-141,
 *     comm_sync_convergence_detection comm.c:235-250, MPI_Allreduce …multisplitting.c:192. */
int msp_group_create(const msp_problem *prob /* block ignored */, int nblocks, const int *devices, msp_group **out);
int msp_group_destroy(msp_group *g);
msp_engine *msp_group_engine(msp_group *g, int k);
int msp_group_solve(msp_group *g, const msp_solve_opts *o, msp_result *res /* [nblocks] */);

/* (b) one process per GPU (torchrun / any launcher): NCCL for the scalar, Gram/TSQR and flag reductions,
 *     CUDA-IPC mapped peer buffers for the P2P halo stores and the async mailboxes. */
int msp_comm_unique_id(void *id128);                                  /* ncclGetUniqueId */
int msp_comm_init(msp_engine *e, const void *id128, int rank, int nranks); /* ncclCommInitRank */
int msp_comm_export(msp_engine *e, void *handle64);                   /* cudaIpcGetMemHandle of the receive window */
int msp_comm_connect(msp_engine *e, int side, const void *handle64);  /* cudaIpcOpenMemHandle of neighbour side */
int msp_comm_connect_block(msp_engine *e, int block, const void *handle64); /* any block: the async global minimisation
                                                                               publishes its TSQR factor to every block */
int msp_solve(msp_engine *e, const msp_solve_opts *o, msp_result *res); /* the drivers' do { } while loops */

/* async convergence detection, one step of this block's state machine on the device
 * (comm_async_convDetection_prime + 4 receive handlers, conv_detection_prime.c:11-498) */
int msp_conv_detect_step(msp_engine *e, int under_threshold, int *state, int *phase_tag);

#ifdef __cplusplus
}
#endif
#endif

/*
 * msolve.c — C host driver of the multisplitting solve path (one binary for every --alg).
 *
 * Replaces the reference's nine main() programs (src/<alg>/<alg>.c, makefile:120-148) and the command line
 * iSolve builds (iSolve:347-401).  It accepts the reference's PETSc-style options — application options
 * -m -n -s -npb -rtol (…-global.c:74-78), -atol (gmres_solution.c:45), -min_convergence_count (parsed, unused by
 * the _prime drivers), per-block KSP prefixes inner{K}_ / outer{K}_ (…multisplitting.c:129-143) and the iSolve
 * spelling inner_ / outer_ for "all blocks", un-prefixed -ksp_* for the stand-alone GMRES — ignores unknown keys
 * like the PETSc options database does, and prints the log lines the reference's log scrapers rely on
 * (utils.c:668-729).  All numerics happen behind the C-ABI of libmsplit.so (include/msplit.h).
 *
 * Extensions: -minimizer tsqr|lsqr|gram|cg|cgne (default tsqr = exact least squares; gram = normal equations on R'R by
 * Cholesky; cg = the same system by PETSc's CG, the reference's `outer_solver` with its default -outer_ksp_type cg;
 * cgne = `outer_solver_cgne`; lsqr = the reference's PETSc LSQR; the iterative ones are driven by
 * -outer{K}_ksp_max_it / -outer{K}_ksp_rtol / -outer{K}_ksp_atol), -detector prime|legacy (asynchronous termination:
 * conv_detection_prime.c or the counter-based conv_detection.c with -min_convergence_count and -max_traversal_ms),
 * -max_seconds T (wall-clock cap of the outer loop, the scripts' `timeout` wrapper), -alg <iSolve name | reference binary name>, -p <depth> (3-D, poisson3DMatrix), -nblocks G
 * (default: the reference's np/npb = 2; 1 for GMRES), -devices 0,1,... (one entry per block, default: block K on
 * GPU K mod #GPUs), -max_outer N, -period a,b,... (deterministic asynchronous schedule, tests only).
 */
#include "../../include/msplit.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { int argc; char **argv; } optdb;

static const char *opt_find(const optdb *db, const char *key, int *present) {
  /* last occurrence wins, like the PETSc options database */
  const char *val = NULL;
  if (present) *present = 0;
  for (int i = 1; i < db->argc; i++) {
    if (strcmp(db->argv[i], key) == 0) {
      if (present) *present = 1;
      val = (i + 1 < db->argc && !(db->argv[i + 1][0] == '-' && isalpha((unsigned char)db->argv[i + 1][1]))) ? db->argv[i + 1] : "";
    }
  }
  return val;
}
static int opt_int(const optdb *db, const char *key, int *out) {
  const char *v = opt_find(db, key, NULL);
  if (!v || !*v) return 0;
  *out = atoi(v);
  return 1;
}
static int opt_real(const optdb *db, const char *key, double *out) {
  const char *v = opt_find(db, key, NULL);
  if (!v || !*v) return 0;
  *out = atof(v);
  return 1;
}
static int opt_flag(const optdb *db, const char *key) {
  int present = 0;
  const char *v = opt_find(db, key, &present);
  if (!present) return 0;
  if (v && *v && (strcmp(v, "0") == 0 || strcasecmp(v, "false") == 0 || strcasecmp(v, "no") == 0)) return 0;
  return 1;
}

static void ksp_defaults(msp_ksp_opts *o) {
  /* PETSc 3.22.1 defaults, tmp/petscmpiexec_help:336-342,602-615 */
  o->restart = 30; o->max_it = 10000; o->rtol = 1e-5; o->abstol = 1e-50; o->divtol = 1e4;
  o->initial_rtol = 0; o->guess_nonzero = 0; o->cgs_refine = 0; o->mgs = 0; o->min_it = 0;
}

/* read -<prefix>ksp_* into o; returns -1 on an unsupported value */
static int ksp_from_options(const optdb *db, const char *prefix, msp_ksp_opts *o) {
  char key[128];
  const char *v;
#define KEY(name) (snprintf(key, sizeof key, "-%s%s", prefix, name), key)
  opt_int(db, KEY("ksp_gmres_restart"), &o->restart);
  opt_int(db, KEY("ksp_max_it"), &o->max_it);
  opt_real(db, KEY("ksp_rtol"), &o->rtol);
  opt_real(db, KEY("ksp_atol"), &o->abstol);
  opt_real(db, KEY("ksp_divtol"), &o->divtol);
  opt_int(db, KEY("ksp_min_it"), &o->min_it);
  if (opt_flag(db, KEY("ksp_converged_use_initial_residual_norm"))) o->initial_rtol = 1;
  if (opt_flag(db, KEY("ksp_gmres_modifiedgramschmidt"))) o->mgs = 1;
  v = opt_find(db, KEY("ksp_gmres_cgs_refinement_type"), NULL);
  if (v && *v) {
    if (!strcasecmp(v, "refine_never")) o->cgs_refine = 0;
    else if (!strcasecmp(v, "refine_ifneeded")) o->cgs_refine = 1;
    else if (!strcasecmp(v, "refine_always")) o->cgs_refine = 2;
    else { fprintf(stderr, "msolve: unknown %s %s\n", key, v); return -1; }
  }
  v = opt_find(db, KEY("ksp_type"), NULL);
  if (v && *v && !strcasecmp(v, "preonly")) {
    /* -innerK_ksp_type preonly -innerK_pc_type lu (running_bulk_test_local:241-242): an EXACT inner solve.  There is no sparse
     * LU on the device path; the same iterate (to rounding) comes from GMRES(64) run to rtol 1e-12 of the entry residual,
     * capped at 20 cycles.  Anything but lu / cholesky behind preonly is refused. */
    const char *pc = opt_find(db, KEY("pc_type"), NULL);
    if (!pc || (strcasecmp(pc, "lu") && strcasecmp(pc, "cholesky"))) { fprintf(stderr, "msolve: -%sksp_type preonly needs -%spc_type lu (exact inner solve); other preconditioners are not on the device path\n", prefix, prefix); return -1; }
    o->restart = MSP_MAX_RESTART; o->max_it = 20 * MSP_MAX_RESTART; o->rtol = 1e-12; o->abstol = 1e-300; o->initial_rtol = 1;
    return 0;
  }
  if (v && *v && strcasecmp(v, "gmres")) { fprintf(stderr, "msolve: -%sksp_type %s is not on the device path (gmres | preonly + lu)\n", prefix, v); return -1; }
  v = opt_find(db, KEY("pc_type"), NULL);
  if (v && *v && strcasecmp(v, "none")) { fprintf(stderr, "msolve: -%spc_type %s is not supported (none only)\n", prefix, v); return -1; }
  v = opt_find(db, KEY("ksp_norm_type"), NULL);
  if (v && *v && strcasecmp(v, "unpreconditioned") && strcasecmp(v, "preconditioned") && strcasecmp(v, "default")) {
    fprintf(stderr, "msolve: -%sksp_norm_type %s is not supported\n", prefix, v); return -1; /* pc none: both norms coincide */
  }
#undef KEY
  return 0;
}

static int alg_from_name(const char *s) {
  static const struct { const char *name; int alg; } tab[] = {
    {"SM", MSP_ALG_SM}, {"MSM", MSP_ALG_SM}, {"synchronous-multisplitting", MSP_ALG_SM},
    {"SMSM_GLOBAL", MSP_ALG_SMSM_GLOBAL}, {"synchronous-multisplitting-synchronous-minimization-global", MSP_ALG_SMSM_GLOBAL},
    {"SMSM_SEMI_LOCAL", MSP_ALG_SMSM_SEMI_LOCAL}, {"synchronous-multisplitting-synchronous-minimization-semi-local", MSP_ALG_SMSM_SEMI_LOCAL},
    {"SMSM_LOCAL", MSP_ALG_SMSM_LOCAL}, {"synchronous-multisplitting-synchronous-minimization-local", MSP_ALG_SMSM_LOCAL},
    {"GMRES", MSP_ALG_GMRES}, {"gmres_solution", MSP_ALG_GMRES},
    {"AM", MSP_ALG_AM}, {"asynchronous-multisplitting_prime", MSP_ALG_AM}, {"asynchronous-multisplitting", MSP_ALG_AM},
    {"AMAM_GLOBAL", MSP_ALG_AMAM_GLOBAL}, {"asynchronous-multisplitting-asynchronous-minimization-global_prime", MSP_ALG_AMAM_GLOBAL},
    /* iSolve:50-56 names the asynchronous binaries without the _prime suffix the makefile gives them (makefile:135,148) */
    {"asynchronous-multisplitting-asynchronous-minimization-global", MSP_ALG_AMAM_GLOBAL},
    {"asynchronous-multisplitting-asynchronous-minimization-semi-local", MSP_ALG_AMAM_SEMI_LOCAL},
    {"asynchronous-multisplitting-asynchronous-minimization-local", MSP_ALG_AMAM_LOCAL},
    /* iSolve:56 maps AMAM_SEMI_LOCAL to the *local* binary; dispatched correctly here */
    {"AMAM_SEMI_LOCAL", MSP_ALG_AMAM_SEMI_LOCAL}, {"asynchronous-multisplitting-asynchronous-minimization-semi-local_prime", MSP_ALG_AMAM_SEMI_LOCAL},
    {"AMAM_LOCAL", MSP_ALG_AMAM_LOCAL}, {"asynchronous-multisplitting-asynchronous-minimization-local_prime", MSP_ALG_AMAM_LOCAL},
  };
  for (size_t i = 0; i < sizeof tab / sizeof tab[0]; i++) if (!strcasecmp(s, tab[i].name)) return tab[i].alg;
  return -1;
}

static int parse_int_list(const char *s, int *out, int cap) {
  int n = 0;
  while (s && *s && n < cap) {
    out[n++] = atoi(s);
    s = strchr(s, ',');
    if (s) s++;
  }
  return n;
}

#define CHECK(call)                                                      \
  do {                                                                   \
    if ((call) != 0) {                                                   \
      fprintf(stderr, "msolve: %s\n  at %s\n", msp_last_error(), #call); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(int argc, char **argv) {
  optdb db = {argc, argv};
  /* defaults of config/default_run_variables */
  int m = 1024, n = 1024, p = 1, s = 4, npb = 1, nblocks = -1, max_outer = 0, min_cc = 4;
  double rtol = 1e-3, atol = 1e-100;
  const char *algname = opt_find(&db, "-alg", NULL);
  if (!algname || !*algname) algname = opt_find(&db, "--alg", NULL);
  if (!algname || !*algname) {
    const char *base = strrchr(argv[0], '/');
    algname = base ? base + 1 : argv[0]; /* reference binaries are one per algorithm */
    if (alg_from_name(algname) < 0) algname = "AM"; /* DEFAULT_ALGORITHM */
  }
  const int alg = alg_from_name(algname);
  if (alg < 0) { fprintf(stderr, "msolve: unknown algorithm %s\n", algname); return 2; }
  opt_int(&db, "-m", &m); opt_int(&db, "-n", &n); opt_int(&db, "-p", &p); opt_int(&db, "-s", &s);
  opt_int(&db, "-npb", &npb); opt_real(&db, "-rtol", &rtol); opt_real(&db, "-atol", &atol);
  opt_int(&db, "-min_convergence_count", &min_cc); opt_int(&db, "-max_outer", &max_outer);
  opt_int(&db, "-nblocks", &nblocks);
  if (npb < 1) npb = 1;
  if (nblocks < 0) nblocks = (alg == MSP_ALG_GMRES) ? 1 : 2; /* iSolve:332-338: np/npb == 2 */
  /* -npb P: P GPUs (strips) per Jacobi block, i.e. nblocks * P engines; the inner GMRES of a block is then distributed over them
   * (the synchronous drivers and GMRES).  The asynchronous drivers keep one strip per block and say so. */
  if (npb > 1 && alg >= MSP_ALG_AM) { /* the asynchronous drivers */
    fprintf(stderr, "msolve: -npb %d ignored for this algorithm: a Jacobi block is one GPU (strip) here\n", npb);
    npb = 1;
  }
  const int njacobi = nblocks;
  nblocks *= npb; /* engines */
  if (nblocks > MSP_MAX_BLOCKS) { fprintf(stderr, "msolve: too many blocks\n"); return 2; }
  const int uses_s = !(alg == MSP_ALG_SM || alg == MSP_ALG_AM || alg == MSP_ALG_GMRES);

  msp_ksp_opts inner;
  ksp_defaults(&inner);
  if (alg == MSP_ALG_GMRES) {
    if (ksp_from_options(&db, "", &inner)) return 2;
  } else {
    /* "inner_" (iSolve:349) = all blocks; inner{K}_ per block.  One option set drives every block: the reference's
     * command lines always give identical values; differing sets are rejected instead of silently merged. */
    if (ksp_from_options(&db, "inner_", &inner)) return 2;
    msp_ksp_opts first = inner;
    for (int k = 1; k <= njacobi; k++) {
      char pre[32];
      snprintf(pre, sizeof pre, "inner%d_", k);
      msp_ksp_opts o = inner;
      if (ksp_from_options(&db, pre, &o)) return 2;
      if (k == 1) first = o;
      else if (memcmp(&o, &first, sizeof o)) { fprintf(stderr, "msolve: -inner%d_* differs from -inner1_*: per-block inner options must agree\n", k); return 2; }
    }
    inner = first;
    /* outer{K}_ksp_* select the reference's LSQR settings; they drive the device LSQR when -minimizer lsqr is given and
     * are otherwise unused (the default minimiser is an exact least-squares solve, TSQR) */
  }
  int outer_type = 0, outer_max_it = 100;
  double outer_rtol = 1e-15, outer_atol = 1e-100;
  {
    const char *mz = opt_find(&db, "-minimizer", NULL);
    if (mz && *mz) {
      if (!strcasecmp(mz, "lsqr")) outer_type = 1;
      else if (!strcasecmp(mz, "gram") || !strcasecmp(mz, "normal")) outer_type = 2;
      else if (!strcasecmp(mz, "cg")) outer_type = 3;
      else if (!strcasecmp(mz, "cgne")) outer_type = 4;
      else if (strcasecmp(mz, "tsqr") && strcasecmp(mz, "qr")) { fprintf(stderr, "msolve: -minimizer %s unknown (tsqr | lsqr | gram | cg | cgne)\n", mz); return 2; }
    }
    const char *pre[] = {"outer_", "outer1_"};
    for (int i = 0; i < 2; i++) {
      char key[64];
      snprintf(key, sizeof key, "-%sksp_max_it", pre[i]); opt_int(&db, key, &outer_max_it);
      snprintf(key, sizeof key, "-%sksp_rtol", pre[i]); opt_real(&db, key, &outer_rtol);
      snprintf(key, sizeof key, "-%sksp_atol", pre[i]); opt_real(&db, key, &outer_atol);
    }
  }

  int devices[MSP_MAX_BLOCKS], periods[MSP_MAX_BLOCKS];
  const int ngpu = msp_device_count();
  if (ngpu < 1) { fprintf(stderr, "msolve: no CUDA device (there is no CPU fallback)\n"); return 3; }
  for (int k = 0; k < nblocks; k++) devices[k] = k % ngpu;
  const char *dl = opt_find(&db, "-devices", NULL);
  if (dl && *dl) { int got = parse_int_list(dl, devices, nblocks); for (int k = got; k < nblocks; k++) devices[k] = devices[k % (got ? got : 1)]; }
  memset(periods, 0, sizeof periods);
  const char *pl = opt_find(&db, "-period", NULL);
  if (pl && *pl) parse_int_list(pl, periods, nblocks);

  msp_problem prob;
  memset(&prob, 0, sizeof prob);
  prob.dim = p > 1 ? 3 : 2; prob.m = m; prob.n = n; prob.p = p; prob.nblocks = nblocks;
  prob.s = uses_s ? s : 0; prob.max_restart = inner.restart; prob.keep_csr = 0; prob.npb = npb;

  if (alg == MSP_ALG_GMRES && nblocks == 1) {
    /* gmres_solution.c:50-85 */
    msp_engine *e = NULL;
    msp_result res;
    if (m != n) { fprintf(stderr, "msolve: poisson2DMatrix_complete assumes a square mesh (utils.c:390)\n"); return 2; }
    CHECK(msp_create(&prob, devices[0], &e));
    msp_ksp_opts o = inner;
    if (opt_find(&db, "-ksp_rtol", NULL) == NULL) o.rtol = 1e-5;
    printf("Start solving...\n");
    CHECK(msp_gmres_solve(e, &o, &res));
    printf("End solving...\n\n\n");
    printf("Elapsed time (iterations):   %f  seconds \n", res.elapsed_s);
    printf("======================== \n");
    printf("Number of iterations of GMRES : %d \n", res.gmres_its);
    printf("Right hand side norm : %e \n", res.norm0);
    printf("GMRES residual norm : %e \n", res.gmres_rnorm);
    printf("||r(i)||/||b|| : %e \n", res.gmres_rnorm / res.norm0);
    printf("======================== \n");
    printf("Erreur : %e \n", res.error);
    printf("\n\n");
    msp_destroy(e);
    return 0;
  }

  msp_group *g = NULL;
  struct timespec ts0, ts1;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  CHECK(msp_group_create(&prob, nblocks, devices, &g));
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  const double stage_loading_s = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec); /* "Loading" stage: assembly, split, vectors, b = A 1 */
  msp_solve_opts so;
  memset(&so, 0, sizeof so);
  {
    double max_seconds = 0.0, trav = 0.0;
    const char *det = opt_find(&db, "-detector", NULL);
    opt_real(&db, "-max_seconds", &max_seconds); opt_real(&db, "-max_traversal_ms", &trav);
    so.max_seconds = max_seconds; so.max_traversal_ms = trav; so.min_convergence_count = min_cc;
    if (det && *det) {
      if (!strcasecmp(det, "legacy")) so.detector = 1;
      else if (strcasecmp(det, "prime")) { fprintf(stderr, "msolve: -detector %s unknown (prime | legacy)\n", det); return 2; }
    }
  }
  so.alg = alg; so.s = uses_s ? s : 0; so.rtol = rtol; so.inner = inner; so.max_outer = max_outer; so.record_history = 1;
  if (alg == MSP_ALG_GMRES && opt_find(&db, "-ksp_rtol", NULL) == NULL) so.inner.rtol = 1e-5;
  so.outer_type = outer_type; so.outer_max_it = outer_max_it; so.outer_rtol = outer_rtol; so.outer_abstol = outer_atol;
  so.profile = opt_flag(&db, "-log_view"); /* per-event device times, like PETSc's -log_view (events bracket every launch) */
  for (int k = 0; k < nblocks; k++) so.period[k] = periods[k];
  msp_result *res = (msp_result *)calloc((size_t)nblocks, sizeof(msp_result));
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  CHECK(msp_group_solve(g, &so, res));
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  const double solve_wall_s = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec);

  if (alg == MSP_ALG_GMRES) {
    /* gmres_solution.c:78-85 with the matrix spread over -npb GPUs */
    printf("Elapsed time (iterations):   %f  seconds \n", res[0].elapsed_s);
    printf("======================== \n");
    printf("Number of iterations of GMRES : %d \n", res[0].gmres_its);
    printf("Right hand side norm : %e \n", res[0].norm0);
    printf("GMRES residual norm : %e \n", res[0].gmres_rnorm);
    printf("||r(i)||/||b|| : %e \n", res[0].gmres_rnorm / res[0].norm0);
    printf("======================== \n");
    printf("Erreur : %e \n", res[0].error);
    printf("[msolve] alg=%s gpus_per_block=%d rel_residual=%e\n", algname, npb, res[0].final_residual / res[0].norm0);
    free(res);
    msp_group_destroy(g);
    return 0;
  }
  printf("Global norm of b %e \n", res[0].norm0); /* …multisplitting.c:157 */
  if (opt_flag(&db, "-print_history"))
    for (int i = 0; i < res[0].hist_len; i++) printf("Final residual norm 2 = %e \n", res[0].hist[i]); /* …multisplitting.c:195 */
  double elapsed = 0.0;
  for (int k = 0; k < nblocks; k++) if (res[k].elapsed_s > elapsed) elapsed = res[k].elapsed_s;
  printf("Elapsed time (iterations):   %f  seconds \n", elapsed); /* utils.c:671 */
  for (int k = 0; k < njacobi; k++) { /* one line per Jacobi block (its first strip when a block spans -npb GPUs) */
    const msp_result *rk = &res[k * npb];
    if (uses_s) printf("[ Block rank %d ] Total number of iterations (outer_iterations * s) = %d * %d = %d \n", k, rk->outer_its, s, s * rk->outer_its); /* utils.c:727 */
    else printf("[ Block rank %d ] Total number of iterations (outer_iterations) = %d \n", k, rk->outer_its); /* utils.c:706 */
  }
  printf("Final residual norm 2 = %e \n", res[0].final_residual); /* utils.c:699 */
  printf("Erreur : %e \n", res[0].error);                        /* …multisplitting.c:229 */
  if (opt_flag(&db, "-log_view")) {
    /* the reference's four PetscLogStages (…multisplitting.c:52-62, …-global.c:81-89), host seconds of block 0 */
    printf("Stage Loading (assembly, splitting, vectors, right-hand side): %f s\n", stage_loading_s);
    printf("Stage I_Solver (inner GMRES solves): %f s\n", res[0].stage_inner_s);
    printf("Stage O_Solver (exchange, A*S, minimisation, convergence test): %f s\n", res[0].stage_outer_s);
    {
      double last = solve_wall_s - res[0].stage_inner_s - res[0].stage_outer_s; /* closing exchange, true residual, error */
      printf("Stage Last (closing exchange, final residual, error): %f s\n", last > 0.0 ? last : 0.0);
    }
    if (res[0].outer_solver_its) printf("Outer solver (LSQR / CG / CGNE) iterations: %lld\n", (long long)res[0].outer_solver_its);
    /* self time per event in microseconds, in the layout of `-log_view ::ascii_flamegraph` (tmp/function-calling-stack:6-13) */
    printf("total solving;I_Solver stage;KSPSolve;KSPGMRESOrthog;VecMDot %.0f\n", res[0].t_mdot_ms * 1e3);
    printf("total solving;I_Solver stage;KSPSolve;KSPGMRESOrthog;VecMAXPY %.0f\n", res[0].t_maxpy_ms * 1e3);
    printf("total solving;I_Solver stage;KSPSolve;MatMult %.0f\n", res[0].t_spmv_ms * 1e3);
    printf("total solving;O_Solver stage;exchange %.0f\n", res[0].t_other_ms * 1e3);
    if (res[0].t_mdot_ms > 0)
      printf("[msolve] algorithmic GB/s: VecMDot %.0f  VecMAXPY(+VecNorm) %.0f  MatMult %.0f\n", res[0].b_mdot / res[0].t_mdot_ms * 1e-6,
             res[0].b_maxpy / res[0].t_maxpy_ms * 1e-6, res[0].b_spmv / res[0].t_spmv_ms * 1e-6);
  }
  long long launches = 0;
  for (int k = 0; k < nblocks; k++) launches += (long long)res[k].kernel_launches;
  printf("[msolve] alg=%s blocks=%d gpus=%d rel_residual=%e kernel_launches=%lld\n", algname, njacobi, ngpu,
         res[0].final_residual / res[0].norm0, launches);
  free(res);
  msp_group_destroy(g);
  (void)atol;
  return 0;
}

// abi.cuh — extern "C" entry points declared in include/msplit.h (included by engine.cu).
#pragma once
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
int msp_spmv_format(const msp_engine *e, int *width) {
  if (!e) return -1;
  if (width) *width = (e->dval || e->dmask) ? e->dia.nd : e->W;
  return e->dmask ? 2 : e->dval ? 1 : 0;
}
int msp_persistent_cycles(const msp_engine *e) {
  if (!e) return -1;
  return (e->coop_mode == 1 || (e->coop_mode == 2 && e->nb <= e->coop_max_rows)) ? 1 : 0;
}
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
// ---- pipelined host <-> device transfers (page-locked host buffers; valid until msp_copies_wait) ----
int msp_set_b_async(msp_engine *e, const double *b) {
  if (!e || !b) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  CK(cudaMemcpyAsync(e->b, b, sizeof(double) * (size_t)e->nb, cudaMemcpyHostToDevice, e->st));
  return 0;
}
int msp_set_x_async(msp_engine *e, const double *x) {
  if (!e || !x) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  CK(cudaMemcpyAsync(e->x, x, sizeof(double) * (size_t)e->nb, cudaMemcpyHostToDevice, e->st));
  return 0;
}
int msp_get_x_async(msp_engine *e, double *x) {
  if (!e || !x) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  if (!e->st_copy) {
    CK(cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_copy, cudaEventDisableTiming));
    CK(cudaMalloc(&e->xstage, sizeof(double) * (size_t)e->ld));
  }
  // the previous download must have left the staging buffer before it is overwritten
  CK(cudaEventRecord(e->ev_copy, e->st_copy));
  CK(cudaStreamWaitEvent(e->st, e->ev_copy, 0));
  CK(cudaMemcpyAsync(e->xstage, e->x, sizeof(double) * (size_t)e->nb, cudaMemcpyDeviceToDevice, e->st)); // snapshot: x may be overwritten next
  CK(cudaEventRecord(e->ev_copy, e->st));
  CK(cudaStreamWaitEvent(e->st_copy, e->ev_copy, 0));
  CK(cudaMemcpyAsync(x, e->xstage, sizeof(double) * (size_t)e->nb, cudaMemcpyDeviceToHost, e->st_copy));
  return 0;
}
int msp_copies_wait(msp_engine *e) {
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  CK(cudaStreamSynchronize(e->st));
  if (e->st_copy) CK(cudaStreamSynchronize(e->st_copy));
  return 0;
}

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

// ---- the §8(b) export list under its own names ----
int msp_get_solution(msp_engine *e, double *x) { return msp_get_x(e, x); }
int msp_split_blocks(msp_engine *e, int which, int32_t *rowptr, int32_t *colidx, double *val) { return msp_get_csr(e, which, rowptr, colidx, val); }
int msp_compute_rhs_ones(msp_engine *e) {
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  return op_compute_rhs_ones(e);
}
int msp_residual_norm(msp_engine *e, double *nrm) {
  // computeFinalResidualNorm utils.c:575-595: sqrt(sum over blocks of ||b_K - A_K,: x||^2), collective
  if (!e || !nrm) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  RC(op_resid_sumsq(e, true, 0));
  RC(allreduce_host(e, 0, 1));
  *nrm = std::sqrt(e->hsc[0]);
  return 0;
}
int msp_exchange_sync(msp_engine *e) {
  // comm_sync_send_and_receive comm.c:126-141: my boundary layers into the neighbours' windows, wait for theirs, collect
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  for (int side = 0; side < 2; side++)
    if (e->has_nb[side] && !e->peer[side].base) MSP_FAIL("neighbour window not connected");
  RC(op_publish_boundary(e));
  RC(exchange_sync(e));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_exchange_async_publish(msp_engine *e, int iteration) {
  // comm_async_test_and_send_prime comm.c:531-554
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  RC(async_publish(e, iteration));
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_exchange_async_poll(msp_engine *e, int *accepted /* [2] or null */) {
  // comm_async_probe_and_receive_prime comm.c:455-529 + receive_data_dependency conv_detection_prime.c:603-633
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  RC(async_probe(e));
  ProbeDecision hd[2];
  CK(cudaMemcpyAsync(hd, e->dec, sizeof(hd), cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  if (accepted) for (int side = 0; side < 2; side++) accepted[side] = e->has_nb[side] ? hd[side].copy : 0;
  return 0;
}
int msp_async_reset(msp_engine *e) {
  // start of an asynchronous run: convergence-detection state, mailboxes and header counters (…multisplitting_prime.c:280-315)
  if (!e) MSP_FAIL("null engine");
  cudaSetDevice(e->device);
  k_cd_init<<<1, 32, 0, e->st>>>(e->cd, e->prob.block, e->prob.nblocks, e->win.mailbox(), e->peer[0].base ? e->peer[0].mailbox() : nullptr,
                                 e->peer[1].base ? e->peer[1].mailbox() : nullptr, e->win.hdr(0), e->win.hdr(1), e->aint);
  e->launches++;
  CK(cudaStreamSynchronize(e->st));
  return 0;
}
int msp_connect_local(msp_engine *e, int side, msp_engine *neighbour) {
  // both engines live in this process: the neighbour's receive window is addressed directly (peer access when on another GPU)
  if (!e || !neighbour || side < 0 || side > 1) MSP_FAIL("bad argument");
  if (!e->has_nb[side]) MSP_FAIL("no neighbour on that side");
  const int want = side == 0 ? e->prob.block - 1 : e->prob.block + 1;
  if (neighbour->prob.block != want || neighbour->prob.nblocks != e->prob.nblocks || neighbour->H != e->H) MSP_FAIL("that engine is not the neighbour block of this side");
  cudaSetDevice(e->device);
  if (neighbour->device != e->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, e->device, neighbour->device);
    if (!can) MSP_FAIL("peer access between the two GPUs is not available");
    cudaError_t er = cudaDeviceEnablePeerAccess(neighbour->device, 0);
    if (er != cudaSuccess && er != cudaErrorPeerAccessAlreadyEnabled) MSP_FAIL("cudaDeviceEnablePeerAccess failed");
    cudaGetLastError();
  }
  e->peer[side] = neighbour->win; e->peer_ipc[side] = false;
  e->peer_any[want] = neighbour->win; e->peer_any_ipc[want] = false;
  return 0;
}
int msp_minimize(msp_engine *e, int kind, int outer_type, int outer_max_it, double outer_rtol, double *alpha, double *resnorm) {
  // outer_solver_norm_equation[_modify] utils.c:1061-1103 (+ the LSQR / CG / CGNE menu, utils.c:972-1043) on R = A S
  // computed by msp_spmm_AS: least squares, then x = S alpha.  kind GLOBAL is collective over the blocks.
  if (!e || !alpha) MSP_FAIL("null argument");
  cudaSetDevice(e->device);
  int alg = kind;
  if (kind == MSP_ALG_AMAM_GLOBAL) alg = MSP_ALG_SMSM_GLOBAL;
  else if (kind == MSP_ALG_AMAM_SEMI_LOCAL) alg = MSP_ALG_SMSM_SEMI_LOCAL;
  else if (kind == MSP_ALG_AMAM_LOCAL) alg = MSP_ALG_SMSM_LOCAL;
  if (alg != MSP_ALG_SMSM_GLOBAL && alg != MSP_ALG_SMSM_SEMI_LOCAL && alg != MSP_ALG_SMSM_LOCAL) MSP_FAIL("kind must be a GLOBAL, SEMI_LOCAL or LOCAL minimisation");
  if (e->smax < 1) MSP_FAIL("engine was created without a minimisation basis (s = 0)");
  const MinimizeOpts mo{outer_type, outer_max_it > 0 ? outer_max_it : 100, outer_rtol > 0 ? outer_rtol : 1e-15, 1e-100};
  double norm = 0.0;
  RC(op_minimize(e, alg, e->smax, mo, alpha, &norm, nullptr));
  CK(cudaStreamSynchronize(e->st));
  if (resnorm) *resnorm = norm;
  return 0;
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
  if ((op == 7 || op == 8) && (nv < 2 || nv > e->smax + 1)) MSP_FAIL("wide gram / apply: 2 <= columns <= s+1");
  if (op == 8) { // identity factor: the product leaves the columns unchanged
    std::vector<double> T((size_t)nv * nv, 0.0);
    for (int i = 0; i < nv; i++) T[(size_t)i * nv + i] = 1.0;
    CK(cudaMemcpyAsync(e->dfac, T.data(), sizeof(double) * nv * nv, cudaMemcpyHostToDevice, e->st));
    CK(cudaStreamSynchronize(e->st));
  }
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
      case 7: launch_gram_wide(e, nv, e->R, e->dfac + (size_t)nv * nv); break;
      case 8: launch_apply_upper(e, nv, e->R, e->dfac); break;
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
  const int npb = prob->npb > 1 ? prob->npb : 1;
  if (nblocks % npb) MSP_FAIL("the number of engines must be a multiple of npb");
  msp_group *g = new msp_group();
  g->G = nblocks;
  g->sh = new LocalShared(nblocks);
  if (npb > 1) for (int K = 0; K < nblocks / npb; K++) g->bsh.push_back(new LocalShared(npb));
  for (int k = 0; k < nblocks; k++) {
    msp_problem p = *prob;
    p.block = k; p.nblocks = nblocks;
    msp_engine *e = nullptr;
    int rc = engine_create(&p, devices ? devices[k] : 0, &e);
    if (rc) { for (auto *x : g->eng) engine_free(x); for (auto *b : g->bsh) delete b; delete g->sh; delete g; return rc; }
    g->eng.push_back(e);
  }
  // the block threads run their inner solves at the same time: engines on one GPU share it (persistent cycle kernel: cycle_coop.cuh)
  for (auto *e : g->eng) {
    e->coop_share = 0;
    for (auto *o : g->eng) e->coop_share += (o->device == e->device) ? 1 : 0;
    coop_configure(e);
  }
  int rc = group_wire(g);
  if (rc) { for (auto *x : g->eng) engine_free(x); for (auto *b : g->bsh) delete b; delete g->sh; delete g; return rc; }
  *out = g;
  return 0;
}
int msp_group_destroy(msp_group *g) {
  if (!g) return 0;
  for (auto *e : g->eng) engine_free(e);
  for (auto *b : g->bsh) delete b;
  delete g->sh;
  delete g;
  return 0;
}
msp_engine *msp_group_engine(msp_group *g, int k) { return (g && k >= 0 && k < g->G) ? g->eng[k] : nullptr; }


int msp_group_solve(msp_group *g, const msp_solve_opts *o, msp_result *res) {
  if (!g || !o || !res) MSP_FAIL("null argument");
  if (o->alg >= MSP_ALG_AM) return engine_solve_async_group(g, o, res);
  std::vector<int> rcs(g->G, 0);
  std::vector<std::string> errs(g->G);
  std::vector<std::thread> th;
  g->sh->reset();
  for (auto *b : g->bsh) b->reset();
  for (int k = 0; k < g->G; k++)
    th.emplace_back([&, k] {
      cudaSetDevice(g->eng[k]->device);
      rcs[k] = (o->alg == MSP_ALG_GMRES) ? engine_gmres(g->eng[k], &o->inner, &res[k]) : engine_solve_sync(g->eng[k], o, &res[k]);
      if (rcs[k]) { errs[k] = g_err; g->sh->abort(); for (auto *b : g->bsh) b->abort(); } // wake the blocks waiting for this one in a collective
    });
  for (auto &t : th) t.join();
  return group_first_error(rcs, errs);
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
  if (e->own_bcomm && e->bcomm) delete e->bcomm;
  if (e->own_comm && e->comm) delete e->comm;
  e->comm = c; e->own_comm = true;
  e->bcomm = c; e->own_bcomm = false;
  if (e->npb > 1) {
    // the communicator of my Jacobi block (the reference's comm_jacobi_block, PetscSubcomm CONTIGUOUS …multisplitting.c:66-73)
    if (!g_nccl.CommSplit) MSP_FAIL("ncclCommSplit not available in this NCCL (needed for npb > 1)");
    NcclComm *b = new NcclComm();
    b->rank = rank % e->npb; b->nranks = e->npb;
    rc = g_nccl.CommSplit(c->comm, rank / e->npb, rank % e->npb, &b->comm, nullptr);
    if (rc) { delete b; MSP_FAIL(std::string("ncclCommSplit: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?")); }
    e->bcomm = b; e->own_bcomm = true;
  }
  // first collective of a communicator sets up its peer connections (tens of milliseconds): here, not inside a timed solve
  RC(e->comm->barrier(e->st));
  if (e->bcomm != e->comm) RC(e->bcomm->barrier(e->st));
  CK(cudaStreamSynchronize(e->st));
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


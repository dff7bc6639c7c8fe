// drivers.cuh — the reference drivers' outer loops: synchronous, stand-alone GMRES, block group, asynchronous (included by engine.cu).
#pragma once
// ------------------------------------------------------------------------------------------------
// the drivers' outer loops (one engine = one block; comm provides barrier / allreduce)
// ------------------------------------------------------------------------------------------------
static int exchange_sync(msp_engine *e) {
  // the boundary layers were already stored into the neighbours' windows by k_update_x / k_publish_boundary
  // (class 3 of the profile = synchronisation + collection of the received layers; bytes = what crossed NVLink into this block)
  e->prof_begin(3, 8.0 * e->H * ((e->has_nb[0] ? 1 : 0) + (e->has_nb[1] ? 1 : 0)));
  int rc = 0;
  StreamWaitValue64Fn wait = e->comm->neighbour_flags() ? stream_wait_value64() : nullptr;
  if (wait) {
    // neighbour-only synchronisation: publish my sequence number to K-1 / K+1, wait for theirs (no collective)
    const unsigned long long seq = ++e->ex_seq;
    unsigned long long *f_lo = e->peer[0].base ? e->peer[0].flags() + 1 : nullptr; // I am the lower block's upper neighbour
    unsigned long long *f_hi = e->peer[1].base ? e->peer[1].flags() + 0 : nullptr;
    if (f_lo || f_hi) { k_signal_neighbours<<<1, 32, 0, e->st>>>(f_lo, f_hi, seq); e->launches++; }
    for (int side = 0; side < 2 && !rc; side++)
      if (e->has_nb[side] && wait(e->st, (unsigned long long)(uintptr_t)(e->win.flags() + side), seq, 0 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) {
        g_err = "cuStreamWaitValue64 failed"; rc = 1;
      }
  } else {
    rc = e->comm->barrier(e->st);
  }
  if (!rc) rc = op_collect_halos(e);
  e->prof_end();
  return rc;
}

static int allreduce_host(msp_engine *e, int first, int n) {
  // sum dsc[first..first+n) over blocks and bring it to hsc
  RC(e->comm->allreduce_sum(e->dsc + first, n, e->st));
  return read_scalars(e, first, n);
}

// outer_solver_norm_equation[_modify] utils.c:1061-1103 on one block: least squares on (R, rhs) with the selected
// minimiser, then x = S alpha on the own rows and the stored neighbour layers.  `global`: one least-squares problem over
// all blocks (collective: every block must call), otherwise block-local.  Returns alpha (host), the residual norm
// ||rhs - R alpha|| where the minimiser provides one (TSQR global / LSQR / normal equations; 0 otherwise).
struct MinimizeOpts { int outer_type, max_it; double rtol, abstol; };
static int op_minimize(msp_engine *e, int alg, int s, const MinimizeOpts &mo, double *alpha, double *norm_out, int *lits_out) {
  const int G = e->prob.nblocks;
  const bool global = (alg == MSP_ALG_SMSM_GLOBAL);
  if (e->npb > 1 && !global && mo.outer_type != 0)
    MSP_FAIL("with npb > 1 the (semi-)local minimisation is solved exactly (TSQR over the block's GPUs); LSQR / CG / normal equations run block-local with npb = 1 only");
  double norm = 0.0;
  int lits = 0;
  if (mo.outer_type == 1) {
    RC(op_lsqr(e, alg, s, global, mo.max_it, mo.rtol, mo.abstol, alpha, &norm, &lits));
  } else if (mo.outer_type == 2) {
    RC(op_normal_equations(e, alg, s, global, alpha, &norm));
  } else if (mo.outer_type == 3 || mo.outer_type == 4) {
    RC(op_cg_normal(e, alg, s, global, mo.outer_type == 4, mo.max_it, mo.rtol, mo.abstol, alpha, &norm, &lits));
  } else {
    std::vector<double> uaug((size_t)(s + 1) * (s + 1));
    RC(op_local_qr(e, alg, s, uaug.data()));
    if (global && G > 1) {
      // TSQR: gather every block's factor (zero-padded allreduce = allgather), identical small solve everywhere
      const int nn = (s + 1) * (s + 1);
      std::vector<double> all((size_t)G * nn, 0.0);
      memcpy(all.data() + (size_t)e->prob.block * nn, uaug.data(), sizeof(double) * nn);
      CK(cudaMemcpyAsync(e->dfac, all.data(), sizeof(double) * (size_t)G * nn, cudaMemcpyHostToDevice, e->st));
      RC(e->comm->allreduce_sum(e->dfac, G * nn, e->st));
      CK(cudaMemcpyAsync(all.data(), e->dfac, sizeof(double) * (size_t)G * nn, cudaMemcpyDeviceToHost, e->st));
      CK(cudaStreamSynchronize(e->st));
      RC(tsqr_combine(s, G, all.data(), alpha, &norm));
    } else if (!global && e->npb > 1) {
      // (semi-)local least squares of a Jacobi block spread over npb GPUs: the TSQR tree of the block's strips only
      const int P = e->npb, nn = (s + 1) * (s + 1);
      std::vector<double> all((size_t)P * nn, 0.0);
      memcpy(all.data() + (size_t)(e->prob.block % P) * nn, uaug.data(), sizeof(double) * nn);
      CK(cudaMemcpyAsync(e->dfac, all.data(), sizeof(double) * (size_t)P * nn, cudaMemcpyHostToDevice, e->st));
      RC(e->bcomm->allreduce_sum(e->dfac, P * nn, e->st));
      CK(cudaMemcpyAsync(all.data(), e->dfac, sizeof(double) * (size_t)P * nn, cudaMemcpyDeviceToHost, e->st));
      CK(cudaStreamSynchronize(e->st));
      RC(tsqr_combine(s, P, all.data(), alpha, &norm));
    } else {
      RC(tsqr_combine(s, 1, uaug.data(), alpha, &norm));
    }
  }
  RC(op_apply_alpha(e, alg, s, alpha));
  if (norm_out) *norm_out = norm;
  if (lits_out) *lits_out = lits;
  return 0;
}

static inline void record_hist(const msp_solve_opts *o, msp_result *res, double v) {
  if (!o->record_history) return;
  if (res->hist_len < 4096) res->hist[res->hist_len++] = v;
  else res->hist_dropped++;
}
struct EventPair { // destroyed on every return path
  cudaEvent_t a = nullptr, b = nullptr;
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};
// wall-clock cap: every block must leave the loop at the same outer iteration, so the local verdicts are summed
static int agree_time_up(msp_engine *e, const msp_solve_opts *o, std::chrono::steady_clock::time_point t_start, bool *stop) {
  *stop = false;
  if (!(o->max_seconds > 0.0)) return 0;
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  e->hsc[7] = (el >= o->max_seconds) ? 1.0 : 0.0;
  if (e->comm->nranks > 1) {
    CK(cudaMemcpyAsync(e->dsc + 7, e->hsc + 7, sizeof(double), cudaMemcpyHostToDevice, e->st));
    RC(allreduce_host(e, 7, 1));
  }
  *stop = e->hsc[7] > 0.0;
  return 0;
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
  // local thresholds rtol / sqrt(number of JACOBI blocks) (…-semi-local.c:330: sqrt(2)); G counts the strips (GPUs)
  const double thr_local = std::max(atol, (o->rtol / std::sqrt((double)(G / e->npb))) * 1.0 * res->norm0);
  // ||rhs_K - A_KK x_K|| of my Jacobi block: the strips' sums of squares added over the block's GPUs
  auto block_local_norm = [&](double *ln) -> int {
    RC(op_resid_sumsq(e, false, 1));
    if (e->npb > 1) RC(e->bcomm->allreduce_sum(e->dsc + 1, 1, e->st));
    RC(read_scalars(e, 1, 1));
    *ln = std::sqrt(e->hsc[1]);
    return 0;
  };
  RC(e->comm->barrier(e->st)); // PetscBarrier before MPI_Wtime
  EventPair ev;
  CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
  CK(cudaEventRecord(ev.a, e->st));
  const int64_t launches0 = e->launches;
  CK(cudaMemsetAsync(reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, its_total), 0, sizeof(long long), e->st));
  e->prof = o->profile != 0;
  struct ProfOff { msp_engine *e; ~ProfOff() { e->prof = false; } } prof_off{e};
  bool done = false, time_up = false;
  const auto t_start = std::chrono::steady_clock::now();
  int sticky = 0;
  const bool lsqr = o->outer_type == 1;
  const int lsqr_max_it = o->outer_max_it > 0 ? o->outer_max_it : 100;
  const double lsqr_rtol = o->outer_rtol > 0 ? o->outer_rtol : 1e-15, lsqr_abstol = o->outer_abstol > 0 ? o->outer_abstol : 1e-100;
  typedef std::chrono::steady_clock clk;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  std::vector<double> alpha(std::max(s, 1));
  const MinimizeOpts mo{o->outer_type, lsqr_max_it, lsqr_rtol, lsqr_abstol};
  if (alg == MSP_ALG_SM) RC(op_update_rhs(e)); // …multisplitting.c:164
  while (!done && !time_up && res->outer_its < max_outer) {
    if (alg == MSP_ALG_SM) {
      auto t0 = clk::now();
      { StageRange stage("I_Solver"); RC(op_inner_solve(e, &in, true, nullptr, nullptr, nullptr, !e->prof)); }
      auto t1 = clk::now();
      res->stage_inner_s += secs(t0, t1);
      RC(exchange_sync(e));
      RC(op_update_rhs(e));
      RC(op_resid_sumsq(e, false, 0));
      RC(allreduce_host(e, 0, 1));
      const double norm = std::sqrt(e->hsc[0]);
      res->last_norm = norm;
      record_hist(o, res, norm);
      if (norm <= thr_global) done = true;
      res->outer_its++;
      res->stage_outer_s += secs(t1, clk::now());
      if (!done) RC(agree_time_up(e, o, t_start, &time_up));
      continue;
    }
    auto t_outer0 = clk::now();
    double inner_this = 0.0;
    for (int t = 0; t < s; t++) {
      RC(op_update_rhs(e));
      auto t0 = clk::now();
      { StageRange stage("I_Solver"); RC(op_inner_solve(e, &in, true, nullptr, nullptr, nullptr, !e->prof)); }
      inner_this += secs(t0, clk::now());
      RC(exchange_sync(e));
      RC(op_push_iterate(e, t));
    }
    // R = A S (MatMatMult …-global.c:326); the LSQR path keeps the reference's raw basis [x^1 .. x^s], the exact solvers
    // use the basis of successive corrections
    StageRange stage_outer("O_Solver");
    RC(op_spmm(e, alg, s, !lsqr));
    int lits = 0;
    double norm = 0.0, ln = 0.0;
    if (alg == MSP_ALG_SMSM_GLOBAL) {
      RC(op_minimize(e, alg, s, mo, alpha.data(), &norm, &lits));
      res->last_norm = norm; // KSPGetResidualNorm(outer_ksp) = phibar (…-global.c:343) = ||b - R alpha||
      record_hist(o, res, norm);
      if (norm <= thr_global) done = true;
    } else {
      if (alg == MSP_ALG_SMSM_SEMI_LOCAL) {
        // pre-minimisation x_K against the stale rhs_K (…-semi-local.c:326)
        RC(block_local_norm(&ln));
      } else if (alg == MSP_ALG_SMSM_LOCAL) {
        RC(op_update_rhs(e)); // …-local.c:258
      } else {
        MSP_FAIL("algorithm not handled by the synchronous driver");
      }
      RC(op_minimize(e, alg, s, mo, alpha.data(), nullptr, &lits));
      if (alg == MSP_ALG_SMSM_LOCAL) RC(block_local_norm(&ln));
      if (ln <= thr_local) sticky = 1;
      e->hsc[2] = sticky; e->hsc[3] = ln * ln;
      CK(cudaMemcpyAsync(e->dsc + 2, e->hsc + 2, 16, cudaMemcpyHostToDevice, e->st));
      RC(allreduce_host(e, 2, 2));
      res->last_norm = ln;
      record_hist(o, res, ln);
      if ((int)std::lround(e->hsc[2]) == G) done = true; // comm_sync_convergence_detection comm.c:235-250
    }
    res->outer_solver_its += lits;
    res->outer_its++;
    res->stage_inner_s += inner_this;
    res->stage_outer_s += secs(t_outer0, clk::now()) - inner_this;
    if (!done) RC(agree_time_up(e, o, t_start, &time_up));
  }
  res->stop_reason = done ? MSP_STOP_CONVERGED : time_up ? MSP_STOP_MAX_SECONDS : MSP_STOP_MAX_OUTER;
  RC(e->comm->barrier(e->st));
  CK(cudaEventRecord(ev.b, e->st));
  CK(cudaEventSynchronize(ev.b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
  res->elapsed_s = ms * 1e-3;
  res->kernel_launches = e->launches - launches0;
  {
    long long tot = 0;
    CK(cudaMemcpy(&tot, reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, its_total), sizeof(long long), cudaMemcpyDeviceToHost));
    res->inner_its_total = tot;
  }
  RC(coop_check(e)); // the deferred inner solves read nothing back: a barrier timeout of the persistent kernel surfaces here
  e->prof_collect(res);
  e->prof = false;
  // closing exchange + true residual + error (comm_sync_send_and_receive_final comm.c:199, utils.c:575, :1045)
  StageRange stage_last("Last");
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
  if (e->prob.nblocks != e->npb) MSP_FAIL("stand-alone GMRES runs on a single block (one GPU, or npb = the number of GPUs)");
  CK(cudaMemsetAsync(e->x, 0, sizeof(double) * e->ld, e->st));
  CK(cudaMemcpyAsync(e->rhs, e->b, sizeof(double) * e->ld, cudaMemcpyDeviceToDevice, e->st));
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->b, 0.0, e->ws, 2, e->dsc + 0);
  RC(allreduce_host(e, 0, 1));
  res->norm0 = std::sqrt(e->hsc[0]);
  msp_ksp_opts in = *o;
  in.guess_nonzero = 0;
  RC(e->comm->barrier(e->st));
  EventPair ev;
  CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
  CK(cudaEventRecord(ev.a, e->st));
  const int64_t l0 = e->launches;
  RC(op_inner_solve(e, &in, false, &res->gmres_its, &res->gmres_reason, &res->gmres_rnorm));
  RC(e->comm->barrier(e->st));
  CK(cudaEventRecord(ev.b, e->st));
  CK(cudaEventSynchronize(ev.b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
  res->elapsed_s = ms * 1e-3;
  res->kernel_launches = e->launches - l0;
  res->outer_its = res->gmres_its;
  res->last_norm = res->gmres_rnorm;
  // true residual and error over all strips of the block (closing exchange of the boundary layers when there are several)
  RC(op_publish_boundary(e));
  RC(exchange_sync(e));
  RC(op_resid_sumsq(e, true, 0));
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->x, 1.0, e->ws, 2, e->dsc + 1);
  RC(allreduce_host(e, 0, 2));
  res->final_residual = std::sqrt(e->hsc[0]);
  res->error = std::sqrt(e->hsc[1]);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// group: all blocks in one process
// ------------------------------------------------------------------------------------------------
struct msp_group {
  int G = 0; // engines (= strips = GPUs); Jacobi blocks = G / npb
  std::vector<msp_engine *> eng;
  LocalShared *sh = nullptr;
  std::vector<LocalShared *> bsh; // one per Jacobi block when a block spans several engines (npb > 1)
};

// the error a group call reports: the first block failure that is not just the echo of another block's abort
static int group_first_error(const std::vector<int> &rcs, const std::vector<std::string> &errs) {
  int first = -1;
  for (size_t k = 0; k < rcs.size(); k++)
    if (rcs[k]) {
      if (first < 0) first = (int)k;
      if (errs[k].find("group aborted") == std::string::npos) { first = (int)k; break; }
    }
  if (first < 0) return 0;
  g_err = errs[first];
  return rcs[first];
}

static int group_wire(msp_group *g) {
  for (int k = 0; k < g->G; k++) {
    msp_engine *e = g->eng[k];
    if (e->own_bcomm && e->bcomm) delete e->bcomm;
    if (e->own_comm && e->comm) delete e->comm;
    e->comm = new LocalComm(g->sh, k); e->own_comm = true;
    if (e->npb > 1) { e->bcomm = new LocalComm(g->bsh[k / e->npb], k % e->npb); e->own_bcomm = true; }
    else { e->bcomm = e->comm; e->own_bcomm = false; }
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
// asynchronous variants (…multisplitting_prime.c:321-393, …-minimization-{global,semi-local,local}_prime.c)
// ------------------------------------------------------------------------------------------------
struct AsyncRun {
  int iters = 0;         // number_of_iterations
  int inner_outer = 0;   // number_of_inner_times_outer_iterations
  int state = 0;
  double thr_local = 0;
  double last_norm = 0;
  std::vector<double> uaug, alpha, all;
  // legacy detector: time since globalCV became (and stayed) true (asynchronous-multisplitting.c.save:307-329)
  bool lg_timer_on = false;
  std::chrono::steady_clock::time_point lg_t0;
};
// after a step: has the run ended?  prime detector: state FINISHED; legacy: globalCV held for MAX_TRAVERSAL_TIME
static bool async_finished(const msp_solve_opts *o, AsyncRun *run) {
  if (o->detector != 1) return run->state == 3;
  if (run->state != 1) { run->lg_timer_on = false; return false; }
  if (!run->lg_timer_on) { run->lg_timer_on = true; run->lg_t0 = std::chrono::steady_clock::now(); return false; }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - run->lg_t0).count();
  return ms > (o->max_traversal_ms > 0.0 ? o->max_traversal_ms : 0.5);
}

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
  if (e->npb > 1) MSP_FAIL("the asynchronous variants run one GPU per Jacobi block (npb = 1)");
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
  if (o->detector == 1)
    k_cd_legacy_step<<<1, 32, 0, e->st>>>(e->cd, e->dsc + 1, run->thr_local, o->min_convergence_count > 0 ? o->min_convergence_count : 4, run->iters);
  else
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
  record_hist(o, res, run->last_norm);
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
  e->prof = o->profile != 0; // per-class CUDA events around every hot launch, like the synchronous driver
  struct ProfOff { msp_engine *e; ~ProfOff() { e->prof = false; } } prof_off{e};
  auto t0 = std::chrono::steady_clock::now();
  bool time_up = false, fin = false;
  while (!fin && run.iters < max_outer && !time_up) {
    RC(async_step(e, o, res, &run));
    fin = async_finished(o, &run);
    if (o->max_seconds > 0.0) time_up = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= o->max_seconds;
  }
  if (o->detector == 1 && fin) { k_cd_legacy_send_global<<<1, 32, 0, e->st>>>(e->cd); e->launches++; } // comm_async_sendGlobalCV
  res->stop_reason = fin ? MSP_STOP_CONVERGED : time_up ? MSP_STOP_MAX_SECONDS : MSP_STOP_MAX_OUTER;
  CK(cudaStreamSynchronize(e->st));
  res->elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  res->kernel_launches = e->launches - l0;
  e->prof_collect(res);
  e->prof = false;
  RC(e->comm->barrier(e->st));
  return async_finish(e, res, &run);
}

static int engine_solve_async_group(msp_group *g, const msp_solve_opts *o, msp_result *res) {
  const int G = g->G;
  bool scheduled = false;
  for (int k = 0; k < G; k++) scheduled |= (o->period[k] > 0);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  g->sh->reset();
  if (!scheduled) {
    std::vector<std::thread> th;
    for (int k = 0; k < G; k++)
      th.emplace_back([&, k] {
        cudaSetDevice(g->eng[k]->device);
        rcs[k] = engine_solve_async(g->eng[k], o, &res[k]);
        if (rcs[k]) { errs[k] = g_err; g->sh->abort(); }
      });
    for (auto &t : th) t.join();
    return group_first_error(rcs, errs);
  }
  // deterministic schedule (tests): block K runs one step at tick t iff t % period[K] == 0, blocks in index order;
  // collective phases (norm0, closing exchange) still need one thread per block
  std::vector<AsyncRun> runs(G);
  auto par_all = [&](auto fn) {
    std::vector<std::thread> th;
    for (int k = 0; k < G; k++)
      th.emplace_back([&, k] { cudaSetDevice(g->eng[k]->device); rcs[k] = fn(k); if (rcs[k]) { errs[k] = g_err; g->sh->abort(); } });
    for (auto &t : th) t.join();
    return group_first_error(rcs, errs);
  };
  RC(par_all([&](int k) { int rc = async_begin(g->eng[k], o, &res[k], &runs[k]); if (!rc) rc = g->eng[k]->comm->barrier(g->eng[k]->st); return rc; }));
  const long long max_ticks = (long long)(o->max_outer > 0 ? o->max_outer : 1000000) * 64;
  int nfin = 0;
  for (long long tick = 0; nfin < G && tick < max_ticks; tick++) {
    for (int k = 0; k < G; k++) {
      const int per = o->period[k] > 0 ? o->period[k] : 1;
      if (tick % per || runs[k].state == 3) continue; // (legacy detector: state 3 is set by the host once the hold has passed)
      cudaSetDevice(g->eng[k]->device);
      RC(async_step(g->eng[k], o, &res[k], &runs[k]));
      CK(cudaStreamSynchronize(g->eng[k]->st));
      if (o->detector == 1) {
        // deterministic schedule: the wall-clock hold becomes "globalCV seen at two consecutive steps of this block"
        if (runs[k].state == 1 && runs[k].lg_timer_on) { runs[k].state = 3; k_cd_legacy_send_global<<<1, 32, 0, g->eng[k]->st>>>(g->eng[k]->cd); }
        runs[k].lg_timer_on = (runs[k].state == 1);
      }
      if (runs[k].state == 3) nfin++;
    }
  }
  if (nfin < G) MSP_FAIL("asynchronous schedule cap reached before every block finished");
  return par_all([&](int k) { return async_finish(g->eng[k], &res[k], &runs[k]); });
}


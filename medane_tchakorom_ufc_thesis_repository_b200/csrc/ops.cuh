// ops.cuh — the operator surface on one block: rhs update, inner GMRES solve, exchange, A*S, minimisers (included by engine.cu).
#pragma once
// ------------------------------------------------------------------------------------------------
// operator surface
// ------------------------------------------------------------------------------------------------
static int op_update_rhs(msp_engine *e) {
  if (e->nbrow == 0) return 0;
  // rhs_K = b_K - A_KJ x_J: only the neighbours that belong to ANOTHER Jacobi block (with npb > 1 a neighbour strip of my own
  // block is part of A_KK)
  k_update_rhs<<<grid_for(e->nbrow), MSPK_THREADS, 0, e->st>>>(e->nbrow, e->brow, e->nb, e->W, e->H, e->ld, e->ecol, e->eval,
                                                               (e->has_nb[0] && !e->intra[0]) ? e->halo[0] : nullptr,
                                                               (e->has_nb[1] && !e->intra[1]) ? e->halo[1] : nullptr, e->b, e->rhs);
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
  } else if (e->npb > 1) {
    // A_KK of a block spread over several GPUs: the couplings to the strips of my own block count (x halos of the last exchange)
    a.lo = e->intra[0] ? e->halo[0] : nullptr; a.hi = e->intra[1] ? e->halo[1] : nullptr;
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

// ------------------------------------------------------------------------------------------------
// A Jacobi block spread over npb GPUs (the reference's -npb > 1: comm_jacobi_block of npb ranks, PETSc's MatMult_MPIAIJ
// scatter + VecMDot_MPI / VecNorm_MPI allreduces inside KSPSolve).  Each GPU owns one strip of the block.
//  * before every SpMV the boundary layers of the vector being multiplied travel to the block's neighbouring strips
//    (P2P stores into their windows + neighbour flags, like the x exchange; the layers are published already scaled);
//  * the MDot results and the sum of squares of the new basis vector are summed over the block's communicator, then a
//    one-thread kernel closes the step — every GPU of the block keeps an identical copy of the GMRES control state;
//  * everything is still decided on the device: the host only enqueues.
// ------------------------------------------------------------------------------------------------
static int exchange_intra(msp_engine *e, const double *v, const double *scale, int guard_it) {
  if (!e->intra[0] && !e->intra[1]) return 0;
  const int par = e->vpar;
  double *p_lo = e->intra[0] ? e->peer[0].vhalo(1, par) : nullptr; // my first layer is the lower strip's upper halo
  double *p_hi = e->intra[1] ? e->peer[1].vhalo(0, par) : nullptr;
  k_publish_boundary_scaled<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->nb, e->H, v, scale, e->ctl, guard_it, p_lo, p_hi);
  e->launches++;
  StreamWaitValue64Fn wait = e->bcomm->neighbour_flags() ? stream_wait_value64() : nullptr;
  if (wait) {
    const unsigned long long seq = ++e->vex_seq;
    unsigned long long *f_lo = e->intra[0] ? e->peer[0].flags() + 3 : nullptr;
    unsigned long long *f_hi = e->intra[1] ? e->peer[1].flags() + 2 : nullptr;
    k_signal_neighbours<<<1, 32, 0, e->st>>>(f_lo, f_hi, seq);
    e->launches++;
    for (int side = 0; side < 2; side++)
      if (e->intra[side] && wait(e->st, (unsigned long long)(uintptr_t)(e->win.flags() + 2 + side), seq, 0) != 0) MSP_FAIL("cuStreamWaitValue64 failed");
  } else {
    RC(e->bcomm->barrier(e->st));
  }
  e->vpar ^= 1;
  return 0;
}
// the layers exchange_intra just delivered
static const double *intra_lo(const msp_engine *e) { return e->intra[0] ? e->win.vhalo(0, e->vpar ^ 1) : nullptr; }
static const double *intra_hi(const msp_engine *e) { return e->intra[1] ? e->win.vhalo(1, e->vpar ^ 1) : nullptr; }

static int op_inner_solve_dist(msp_engine *e, const msp_ksp_opts *o, bool publish, int *its_out, int *reason_out, double *rnorm_out, bool defer) {
  if (o->restart < 1 || o->restart > e->prob.max_restart) MSP_FAIL("restart exceeds max_restart of the engine");
  if (o->mgs) MSP_FAIL("modified Gram-Schmidt is not available for a Jacobi block spread over several GPUs (npb > 1)");
  if (e->bcomm->nranks != e->npb) MSP_FAIL("npb > 1 but the block's communicator is not set up: create the engines through msp_group_create or call msp_comm_init");
  defer = defer && !its_out && !reason_out && !rnorm_out && o->max_it <= 4 * o->restart;
  const bool guess_zero = !o->guess_nonzero;
  double *bnorm_sq = nullptr;
  if (!guess_zero && !o->initial_rtol) {
    k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, 0.0, e->ws, 2, e->dsc + 8);
    e->launches++;
    RC(e->bcomm->allreduce_sum(e->dsc + 8, 1, e->st));
    bnorm_sq = e->dsc + 8;
  }
  k_ctl_begin<<<1, 32, 0, e->st>>>(e->ctl, o->restart, o->max_it, o->min_it, o->initial_rtol, guess_zero ? 1 : 0, o->cgs_refine, o->rtol,
                                   o->abstol, o->divtol, bnorm_sq);
  e->launches++;
  double *lhh = reinterpret_cast<double *>(reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, lhh));
  const double *invs = reinterpret_cast<const double *>(reinterpret_cast<const char *>(e->ctl) + offsetof(GmresCtl, inv_arr));
  int itcount = 0;
  bool first = true;
  struct { int its, it, reason, active; } hc{};
  while (true) {
    const int nsteps = std::min(o->restart, o->max_it - itcount);
    double *peer_lo = (publish && e->peer[0].base) ? e->peer[0].halo(1, e->par) : nullptr;
    double *peer_hi = (publish && e->peer[1].base) ? e->peer[1].halo(0, e->par) : nullptr;
    // ---- cycle prologue: vtilde_0 = rhs - A_KK x over the whole block
    if (first && guess_zero) {
      k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, e->V);
      k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, 0.0, e->ws, 2, e->dsc + 9);
      e->launches += 2;
    } else {
      RC(exchange_intra(e, e->x, nullptr, -1)); // the block's current iterate on the neighbouring strips
      SpmvArgs a = spmv_args(e, e->x, e->V);
      a.b = e->rhs; a.lo = intra_lo(e); a.hi = intra_hi(e);
      launch_spmv_w<1, true, false, true>(e, a, 0, nullptr);
      CK(cudaMemcpyAsync(e->dsc + 9, e->ws.partial + 0 * MSPK_MAX_PART + MSPK_MAX_PART - 1, sizeof(double), cudaMemcpyDeviceToDevice, e->st));
    }
    RC(e->bcomm->allreduce_sum(e->dsc + 9, 1, e->st));
    k_ctl_cycle_begin_from<<<1, 32, 0, e->st>>>(e->ctl, e->dsc + 9);
    e->launches++;
    for (int it = 0; it < nsteps; it++) {
      double *w = e->V + (long long)(it + 1) * e->ld;
      const double *vit = e->V + (long long)it * e->ld;
      RC(exchange_intra(e, vit, invs + it, it));
      SpmvArgs a = spmv_args(e, vit, w);
      a.guard_it = it; a.lo = intra_lo(e); a.hi = intra_hi(e);
      launch_spmv_w<1, false, true, false>(e, a, 0, nullptr);
      for (int pass = 0; pass < (o->cgs_refine ? 2 : 1); pass++) {
        launch_mdot(e, it + 1, e->V, e->ld, w, lhh, -1.0, it, pass, invs);
        RC(e->bcomm->allreduce_sum(lhh, it + 1, e->st));
        launch_maxpy<2>(e, it + 1, e->V, e->ld, lhh, w, e->dsc + 10, it, pass, pass, 0, invs);
        RC(e->bcomm->allreduce_sum(e->dsc + 10, 1, e->st));
        k_step_end<<<1, 32, 0, e->st>>>(e->ctl, e->dsc + 10, pass, it, pass);
        e->launches++;
      }
    }
    // ---- KSPGMRESBuildSoln + publication of the new iterate's layers for the exchange of the outer loop
    k_build_soln_coef<<<1, 32, 0, e->st>>>(e->ctl);
    UpdateXArgs u{};
    u.nb = e->nb; u.H = e->H; u.ld = e->ld; u.V = e->V; u.x = e->x; u.ctl = e->ctl; u.peer_lo = peer_lo; u.peer_hi = peer_hi;
    k_update_x<<<grid_for(e->nb, 8), MSPK_THREADS, 0, e->st>>>(u);
    e->launches += 2;
    CK(cudaMemcpyAsync(e->hsc + 32, reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, its), 16, cudaMemcpyDeviceToHost, e->st));
    first = false;
    if (defer) {
      itcount += nsteps;
      if (itcount >= o->max_it) return 0;
      continue;
    }
    CK(cudaStreamSynchronize(e->st));
    memcpy(&hc, e->hsc + 32, 16);
    itcount += hc.it;
    if (hc.reason) break;
    if (itcount >= o->max_it) { hc.reason = MSP_DIVERGED_ITS; break; }
    if (hc.it == 0) { hc.reason = MSP_DIVERGED_BREAKDOWN; break; }
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

// inner_solver utils.c:950-970 -> KSPSolve_GMRES.  One host synchronisation per
// restart cycle; inside a cycle every decision is taken on the device.
// `defer`: when the solve is at most four restart cycles (max_it <= 4 restart) and the caller does not ask for
// its / reason / rnorm, nothing is read back and the host does not wait: every cycle the solve can need is enqueued,
// the kernels decide everything on the device (a cycle that starts after the solve has ended — converged, broken down,
// max_it reached — turns all its launches into no-ops), the iteration count is added to ctl->its_total, and the
// caller's next stream synchronisation covers it.
static int op_inner_solve(msp_engine *e, const msp_ksp_opts *o, bool publish, int *its_out, int *reason_out, double *rnorm_out, bool defer = false) {
  if (e->npb > 1) return op_inner_solve_dist(e, o, publish, its_out, reason_out, rnorm_out, defer);
  defer = defer && !its_out && !reason_out && !rnorm_out && o->max_it <= 4 * o->restart;
  if (o->restart < 1 || o->restart > e->prob.max_restart) MSP_FAIL("restart exceeds max_restart of the engine");
  const bool guess_zero = !o->guess_nonzero;
  double *bnorm_sq = nullptr;
  if (!guess_zero && !o->initial_rtol) {
    k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, e->rhs, 0.0, e->ws, 2, e->dsc + 8);
    e->launches++;
    bnorm_sq = e->dsc + 8;
  }
  // PETSc ignores -ksp_gmres_cgs_refinement_type under modified Gram-Schmidt: the MGS step has no second pass
  const int cgs_refine = o->mgs ? 0 : o->cgs_refine;
  k_ctl_begin<<<1, 32, 0, e->st>>>(e->ctl, o->restart, o->max_it, o->min_it, o->initial_rtol, guess_zero ? 1 : 0, cgs_refine, o->rtol,
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
        if (cgs_refine) {
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
    bool cycle_enqueued = false;
    if (coop_eligible(e, o, cgs_refine, from_rhs)) {
      // small block: the whole cycle is one persistent cooperative kernel (cycle_coop.cuh), bit-identical to the launches below
      const int rc = launch_cycle_coop(e, nsteps, peer_lo, peer_hi);
      if (rc == 1) return 1;
      if (rc == 0) {
        CK(cudaMemcpyAsync(e->hsc + 32, reinterpret_cast<char *>(e->ctl) + offsetof(GmresCtl, its), 16, cudaMemcpyDeviceToHost, e->st));
        cycle_enqueued = true;
      }
    }
    if (!cycle_enqueued && e->use_graphs && !e->prof) {
      const msp_engine::CycleKey key(nsteps, cgs_refine + 4 * (o->mgs ? 1 : 0), from_rhs ? 1 : 0, (const void *)e->V, (const void *)peer_lo, (const void *)peer_hi);
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
    } else if (!cycle_enqueued) {
      RC(enqueue_cycle());
    }
    first = false;
    if (defer) {
      itcount += nsteps; // upper bound: the device stops earlier if it converges
      if (itcount >= o->max_it) return 0;
      continue;
    }
    CK(cudaStreamSynchronize(e->st));
    memcpy(&hc, e->hsc + 32, 16);
    itcount += hc.it;
    if (hc.reason == MSPK_COOP_REASON_ABORT) { RC(coop_check(e)); MSP_FAIL("the persistent restart-cycle kernel gave up"); }
    if (hc.reason) break;
    if (itcount >= o->max_it) { hc.reason = MSP_DIVERGED_ITS; break; }
    if (hc.it == 0) { hc.reason = MSP_DIVERGED_BREAKDOWN; break; } // a cycle that made no step can never end the loop
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
  if (diff_basis && (!kind_is_local(kind) || e->npb > 1)) { // (npb > 1: the layers of the block's own neighbouring strips belong to S_K too)
    if (e->has_nb[0]) { k_diff_basis<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, s, e->H, e->Slo); e->launches++; }
    if (e->has_nb[1]) { k_diff_basis<<<grid_for(e->H), MSPK_THREADS, 0, e->st>>>(e->H, s, e->H, e->Shi); e->launches++; }
  }
  const bool local = kind_is_local(kind);
  if (e->dmask) {
    // coded DIA view: s sweeps of the 17 B/row SpMV (85 n bytes for s = 5) beat one pass over the ELL matrix (64 n + 80 n,
    // gather-bound: 3.17 ms against 5 x 0.19 ms at 67 M rows); same fma chain per row, so the columns of R are unchanged
    for (int t = 0; t < s; t++) {
      SpmvArgs v = spmv_args(e, e->S + (long long)t * e->ld, e->R + (long long)t * e->ld);
      if (local && e->npb > 1) {
        // A_KK of a block spread over several GPUs: stored layers of the block's own neighbouring strips, nothing from other blocks
        v.lo = e->intra[0] ? e->Slo + (size_t)t * e->H : nullptr;
        v.hi = e->intra[1] ? e->Shi + (size_t)t * e->H : nullptr;
        launch_spmv_w<1, false, false, false>(e, v, 0, nullptr);
      } else if (local) launch_spmv_w<0, false, false, false>(e, v, 0, nullptr);
      else {
        v.lo = e->has_nb[0] ? e->Slo + (size_t)t * e->H : nullptr;
        v.hi = e->has_nb[1] ? e->Shi + (size_t)t * e->H : nullptr;
        launch_spmv_w<1, false, false, false>(e, v, 0, nullptr);
      }
    }
    return 0;
  }
  SpmmArgs a{};
  a.nb = e->nb; a.W = e->W; a.H = e->H; a.s = s; a.ld = e->ld; a.lds = e->ld; a.ecol = e->ecol; a.eval = e->eval;
  a.S = e->S; a.R = e->R;
  a.Slo = ((!local && e->has_nb[0]) || (local && e->intra[0])) ? e->Slo : nullptr;
  a.Shi = ((!local && e->has_nb[1]) || (local && e->intra[1])) ? e->Shi : nullptr;
  const int g = grid_for(e->nb, 8);
  for (int c0 = 0; c0 < s;) {
    int nc = std::min(8, s - c0);
    // chunk sizes 8,5,4,2,1 cover every s with few passes over the matrix
    int use = nc >= 8 ? 8 : nc >= 5 ? 5 : nc >= 4 ? 4 : nc >= 2 ? 2 : 1;
#define SPMM_CASE(N)                                                                                          \
  case N:                                                                                                     \
    if (local && e->npb == 1) k_spmm_ell<0, N><<<g, MSPK_THREADS, 0, e->st>>>(a, c0);                          \
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
  k<<<grid_for((long long)e->nb / 2, std::min(resident_blocks_per_sm(k), 4)), MSPK_THREADS, 0, e->st>>>(e->nb, e->ld, C, e->gram_partial, e->ws.counter + MSPK_GRAM_COUNTER, out_dev);
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

// Gram matrix of nc > 9 columns at C, panel by panel, upper triangle into out_dev (nc x nc column-major)
static void launch_gram_wide(msp_engine *e, int nc, const double *C, double *out_dev) {
  const int grid = grid_for((long long)e->nb / 2, std::min(resident_blocks_per_sm(k_gram_panel), 4));
  for (int j0 = 0; j0 < nc; j0 += 8)
    for (int i0 = 0; i0 <= j0; i0 += 8) {
      k_gram_panel<<<grid, MSPK_THREADS, 0, e->st>>>(e->nb, e->ld, C + (long long)i0 * e->ld, std::min(8, nc - i0), C + (long long)j0 * e->ld, std::min(8, nc - j0),
                                                     e->gram_partial, e->ws.counter + MSPK_GRAM_COUNTER, out_dev + (long long)j0 * nc + i0, nc);
      e->launches++;
    }
}
// C := C T for an upper-triangular T (device, nc x nc), in place, last panel first
static void launch_apply_upper(msp_engine *e, int nc, double *C, const double *T_dev) {
  const int grid = grid_for((long long)e->nb / 2, std::min(resident_blocks_per_sm(k_apply_upper), 4));
  const int last0 = ((nc - 1) / 8) * 8;
  for (int j0 = last0; j0 >= 0; j0 -= 8) {
    k_apply_upper<<<grid, MSPK_THREADS, sizeof(double) * nc * nc, e->st>>>(e->nb, e->ld, C, nc, j0, std::min(8, nc - j0), T_dev);
    e->launches++;
  }
}
// inverse of an upper-triangular matrix (column-major, host)
static void upper_inverse(int nc, const double *U, double *T) {
  std::fill(T, T + (size_t)nc * nc, 0.0);
  for (int j = 0; j < nc; j++) {
    T[(size_t)j * nc + j] = 1.0 / U[(size_t)j * nc + j];
    for (int i = j - 1; i >= 0; i--) {
      double t = 0.0;
      for (int k = i + 1; k <= j; k++) t += U[(size_t)k * nc + i] * T[(size_t)j * nc + k];
      T[(size_t)j * nc + i] = -t / U[(size_t)i * nc + i];
    }
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
  if (e->use_cholqr && nc >= 2 && nc <= MSP_MAX_S + 1) {
    // nc <= 9: one register-resident Gram kernel and a triangular solve; wider bases: the same in panels of 8 columns,
    // with the inverse factor applied as a product
    const bool wide = nc > 9;
    std::vector<double> T((size_t)nc * nc);
    U1.assign((size_t)nc * nc, 0.0); U2.assign((size_t)nc * nc, 0.0);
    double *Gdev = e->dfac + (size_t)e->prob.nblocks * nc * nc; // behind the G factor slots of the TSQR gather
    if (wide) launch_gram_wide(e, nc, e->R, Gdev); else launch_gram(e, nc, e->R, Gdev);
    CK(cudaMemcpyAsync(G.data(), Gdev, sizeof(double) * nc * nc, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    if (chol_upper(nc, G.data(), U1.data())) {
      if (wide) {
        upper_inverse(nc, U1.data(), T.data());
        CK(cudaMemcpyAsync(Gdev, T.data(), sizeof(double) * nc * nc, cudaMemcpyHostToDevice, e->st));
        launch_apply_upper(e, nc, e->R, Gdev);
      } else {
        CK(cudaMemcpyAsync(Gdev, U1.data(), sizeof(double) * nc * nc, cudaMemcpyHostToDevice, e->st));
        launch_trsolve(e, nc, e->R, Gdev);
      }
      have_u1 = true;
      if (wide) launch_gram_wide(e, nc, e->R, Gdev); else launch_gram(e, nc, e->R, Gdev);
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
  {
    // the block's copies of the neighbours' boundaries follow x_min = S alpha too (…-semi-local.c:335-338); the local variant
    // only rewrites its own block (…-local.c:256-260) — which, spread over several GPUs, includes the neighbouring strips of the block
    const bool local_kind = kind_is_local(kind);
    for (int side = 0; side < 2; side++)
      if (local_kind ? e->intra[side] : e->has_nb[side]) {
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

// The reference's two remaining outer solvers:
//  * `outer_solver` utils.c:972-996 with -outer_ksp_type cg (config/default_run_variables: cg, max_it 100000, rtol 1e-20):
//    MatTransposeMatMult(R,R) and MatMultTranspose(R,b) explicitly, then KSPCG on the s x s system.  Here: ONE Gram pass
//    over [R_K | rhs] (K9) (+ one allreduce of the (s+1)^2 entries when global), PETSc's CG recurrence (cg.c, PC none,
//    zero guess, initial-residual-norm test on the natural residual) on the host — the system is s x s.
//  * `outer_solver_cgne` utils.c:1020-1043 (KSPCGNE on R itself): CG on the normal equations WITHOUT forming R'R —
//    every iteration is w = R p (K10), z = R' r (K3) and r -= a w (K4+K5) on the device, s-vectors on the host.
// Both return ||rhs - R alpha|| (second pass / the maintained residual), the quantity the global driver stops on.
static int op_cg_normal(msp_engine *e, int kind, int s, bool global, bool cgne, int max_it, double rtol, double abstol, double *alpha,
                        double *rnorm_out, int *its_out) {
  const double *rhs_src = kind_is_local(kind) ? e->rhs : e->b;
  auto sum_blocks = [&](int first, int n) -> int { return global ? allreduce_host(e, first, n) : read_scalars(e, first, n); };
  for (int j = 0; j < s; j++) alpha[j] = 0.0;
  int its = 0;
  if (!cgne) {
    const int nc = s + 1;
    if (nc > 9) MSP_FAIL("the CG-on-normal-equations minimiser supports s <= 8");
    double *bcol = e->R + (long long)s * e->ld;
    k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, rhs_src, bcol);
    e->launches++;
    launch_gram(e, nc, e->R, e->dfac);
    if (global) RC(e->comm->allreduce_sum(e->dfac, nc * nc, e->st));
    std::vector<double> G((size_t)nc * nc);
    CK(cudaMemcpyAsync(G.data(), e->dfac, sizeof(double) * nc * nc, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    auto A = [&](int i, int j) { return i <= j ? G[(size_t)j * nc + i] : G[(size_t)i * nc + j]; }; // upper triangle stored
    std::vector<double> r(s), p(s, 0.0), w(s);
    for (int i = 0; i < s; i++) r[i] = A(i, s); // R'b; x = 0
    double beta = 0.0, betaold = 1.0;
    for (int i = 0; i < s; i++) beta += r[i] * r[i];
    double dp = std::sqrt(beta);
    const double ttol = std::max(rtol * dp, abstol);
    if (beta > 0.0 && dp > ttol) {
      for (int i = 0; i < max_it; i++) {
        if (i == 0) p = r;
        else { const double b = beta / betaold; for (int k = 0; k < s; k++) p[k] = r[k] + b * p[k]; }
        double dpi = 0.0;
        for (int k = 0; k < s; k++) { double t = 0.0; for (int l = 0; l < s; l++) t += A(k, l) * p[l]; w[k] = t; dpi += p[k] * t; }
        if (!(dpi > 0.0)) break; // KSP_DIVERGED_INDEFINITE_MAT (or NaN): keep the iterate
        const double a = beta / dpi;
        for (int k = 0; k < s; k++) { alpha[k] += a * p[k]; r[k] -= a * w[k]; }
        betaold = beta;
        beta = 0.0;
        for (int k = 0; k < s; k++) beta += r[k] * r[k];
        dp = std::sqrt(beta);
        its++;
        if (beta == 0.0 || dp <= ttol || std::isnan(dp)) break;
      }
    }
    if (rnorm_out) {
      for (int j = 0; j < s; j++) e->hsc[216 + j] = -alpha[j];
      CK(cudaMemcpyAsync(e->dsc + 216, e->hsc + 216, sizeof(double) * s, cudaMemcpyHostToDevice, e->st));
      launch_maxpy<0>(e, s, e->R, e->ld, e->dsc + 216, bcol, e->dsc + 6, -1, 0, 0, 3);
      RC(read_scalars(e, 6, 1));
      e->hsc[6] = e->hsc[6] * e->hsc[6];
      if (global) { CK(cudaMemcpyAsync(e->dsc + 6, e->hsc + 6, sizeof(double), cudaMemcpyHostToDevice, e->st)); RC(allreduce_host(e, 6, 1)); }
      *rnorm_out = std::sqrt(e->hsc[6]);
    }
    if (its_out) *its_out = its;
    return 0;
  }
  // ---- CGNE: r = rhs (x = 0), z = R' r, p = z; { w = R p; a = |z|^2 / |w|^2; x += a p; r -= a w; z = R' r; p = z + (|z|^2/|z_old|^2) p }
  double *Rv = e->Wb[0], *Wv = e->Wb[1];
  std::vector<double> z(s), p(s, 0.0);
  k_copy<<<grid_for(e->nb / 2), MSPK_THREADS, 0, e->st>>>(e->nb, rhs_src, Rv);
  k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, Rv, 0.0, e->ws, 2, e->dsc + 4);
  e->launches += 2;
  RC(sum_blocks(4, 1));
  double rn = std::sqrt(e->hsc[4]);
  launch_mdot(e, s, e->R, e->ld, Rv, e->dsc + 64, 1.0, -1, 0);
  RC(sum_blocks(64, s));
  double beta = 0.0, betaold = 1.0;
  for (int j = 0; j < s; j++) { z[j] = e->hsc[64 + j]; beta += z[j] * z[j]; }
  const double ttol = std::max(rtol * std::sqrt(beta), abstol);
  if (beta > 0.0 && std::sqrt(beta) > ttol) {
    for (int i = 0; i < max_it; i++) {
      if (i == 0) p = z;
      else { const double b = beta / betaold; for (int k = 0; k < s; k++) p[k] = z[k] + b * p[k]; }
      memcpy(e->hsc + 216, p.data(), sizeof(double) * s);
      CK(cudaMemcpyAsync(e->dsc + 216, e->hsc + 216, sizeof(double) * s, cudaMemcpyHostToDevice, e->st));
      k_lincomb<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, s, e->ld, e->R, e->dsc + 216, Wv); // w = R p
      k_sumsq<<<grid_for(e->nb), MSPK_THREADS, 0, e->st>>>(e->nb, Wv, 0.0, e->ws, 2, e->dsc + 5);
      e->launches += 2;
      RC(sum_blocks(5, 1));
      const double dpi = e->hsc[5];
      if (!(dpi > 0.0)) break;
      const double a = beta / dpi;
      for (int k = 0; k < s; k++) alpha[k] += a * p[k];
      e->hsc[5] = -a;
      CK(cudaMemcpyAsync(e->dsc + 5, e->hsc + 5, sizeof(double), cudaMemcpyHostToDevice, e->st));
      launch_maxpy<0>(e, 1, Wv, e->ld, e->dsc + 5, Rv, e->dsc + 6, -1, 0, 0, 3); // r -= a w, ||r||
      RC(read_scalars(e, 6, 1));
      e->hsc[6] = e->hsc[6] * e->hsc[6];
      if (global) { CK(cudaMemcpyAsync(e->dsc + 6, e->hsc + 6, sizeof(double), cudaMemcpyHostToDevice, e->st)); RC(allreduce_host(e, 6, 1)); }
      rn = std::sqrt(e->hsc[6]);
      launch_mdot(e, s, e->R, e->ld, Rv, e->dsc + 64, 1.0, -1, 0); // z = R' r
      RC(sum_blocks(64, s));
      betaold = beta;
      beta = 0.0;
      for (int j = 0; j < s; j++) { z[j] = e->hsc[64 + j]; beta += z[j] * z[j]; }
      its++;
      if (beta == 0.0 || std::sqrt(beta) <= ttol || std::isnan(beta)) break;
    }
  }
  if (rnorm_out) *rnorm_out = rn;
  if (its_out) *its_out = its;
  return 0;
}

// small dense least squares on the host: stack nfac upper factors [U_k | c_k; 0 rho_k] and solve by
// Householder QR.  This is the root of the TSQR tree (s <= 32: a few kflop).
// Rank handling: a basis column that is numerically dependent on the ones before it (iterates identical to rounding —
// this happens once a run has converged to machine precision) is dropped, alpha_k = 0, and does NOT consume a pivot
// row: the next column is reflected at the same row, so what the dropped column's row still holds of the right-hand
// side is minimised by the later columns and |diag[s]| stays the true residual norm of the returned alpha.  With no
// dependent column the arithmetic is the plain Householder QR, operation for operation.
static int tsqr_combine(int s, int nfac, const double *uall, double *alpha, double *resnorm) {
  const int nc = s + 1, rows = nfac * nc;
  std::vector<double> A((size_t)rows * nc, 0.0); // column-major rows x nc
  for (int f = 0; f < nfac; f++)
    for (int c = 0; c < nc; c++)
      for (int r = 0; r <= c; r++) A[(size_t)c * rows + f * nc + r] = uall[(size_t)f * nc * nc + (size_t)c * nc + r];
  std::vector<double> diag(nc, 0.0);
  std::vector<int> prow(nc, -1); // pivot row of every column, -1 = dropped
  int rk = 0;                    // next pivot row
  double dmax_so_far = 0.0;
  for (int k = 0; k < nc; k++) {
    double *a = &A[(size_t)k * rows];
    double nrm = 0.0;
    for (int r = rk; r < rows; r++) nrm += a[r] * a[r];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0 || (k < s && nrm <= 1e-14 * dmax_so_far)) { diag[k] = 0.0; continue; }
    double beta = (a[rk] >= 0.0) ? -nrm : nrm;
    a[rk] -= beta;
    double vtv = 0.0;
    for (int r = rk; r < rows; r++) vtv += a[r] * a[r];
    for (int j = k + 1; j < nc; j++) {
      double *aj = &A[(size_t)j * rows];
      double d = 0.0;
      for (int r = rk; r < rows; r++) d += a[r] * aj[r];
      d = 2.0 * d / vtv;
      for (int r = rk; r < rows; r++) aj[r] -= d * a[r];
    }
    diag[k] = beta;
    prow[k] = rk++;
    dmax_so_far = std::max(dmax_so_far, std::fabs(beta));
  }
  // back substitution on the pivot rows of the leading s columns against column s
  const double *cvec = &A[(size_t)s * rows];
  double dmax = 0.0;
  for (int k = 0; k < s; k++) dmax = std::max(dmax, std::fabs(diag[k]));
  for (int k = s - 1; k >= 0; k--) {
    if (prow[k] < 0) { alpha[k] = 0.0; continue; }
    double t = cvec[prow[k]];
    for (int j = k + 1; j < s; j++) t -= A[(size_t)j * rows + prow[k]] * alpha[j];
    // numerically dependent basis vector that still got a pivot (smaller than 1e-14 of a LATER column): drop it
    alpha[k] = (std::fabs(diag[k]) > 1e-14 * dmax) ? t / diag[k] : 0.0;
  }
  if (resnorm) *resnorm = std::fabs(diag[s]);
  return 0;
}


// pnp-b200 — right-preconditioned flexible GMRES with adaptive restart on the device
// (SURVEY §2.3 K15). Stands in for FASP's fasp_solver_dbsr_pvfgmres [EXT] reached through
// fasp_solver_dbsr_krylov_amg (benchmarks/pnp_exact_solution/linear_pnp.cpp:101): same
// restart schedule (restart_min 3, step 3, cr_max 0.99, cr_min 0.174), Givens recurrences,
// stopping rule ||r|| <= tol*||b|| and return convention (iterations >= 0, ERROR_SOLVER_MAXIT).
// Orthogonalisation is classical Gram-Schmidt with DGKS re-orthogonalisation: all
// projections of a step come from ONE fused pass over the basis (multi_dot) and one fused
// update (multi_axpy), i.e. 2 host synchronisations per iteration instead of FASP's j+1.
#include "internal.hpp"
#include <cmath>

namespace pnpb200 {

namespace {
struct PhaseTimer {
  Handle& h; double& acc; bool on;
  PhaseTimer(Handle& hh, double& a) : h(hh), acc(a), on(hh.timing_detail != 0)
  { if (on) cudaEventRecord(h.ev[4], h.stream); }
  ~PhaseTimer()
  {
    if (!on) return;
    cudaEventRecord(h.ev[5], h.stream); cudaEventSynchronize(h.ev[5]);
    float ms = 0; cudaEventElapsedTime(&ms, h.ev[4], h.ev[5]); acc += ms;
  }
};
}

namespace {
// the solver proper: b and x are in SWEEP order (internal.hpp), like every vector below
int fgmres_sweep(Handle& h, const SolverConfig& cfg, const double* b, double* x, bool zero_guess)
{
  cudaStream_t s = h.stream;
  const BsrT& A = h.hier.lv[0]->A;
  // The Krylov basis V is PACKED (KS = 3 doubles per row): an Arnoldi step streams the whole basis twice
  // (multi_dot, multi_axpy), padding it would cost 33 % more traffic there. The preconditioned vectors Z
  // and the iterate x are gathered by the matrix passes and therefore padded (VS doubles per row).
  constexpr int KS = 3;
  const size_t n = KS * (size_t)A.n;         // V: local vector length (owned + ghost rows) = basis stride
  const size_t m = KS * (size_t)owned_rows(*h.hier.lv[0]);   // reductions and updates run over the owned rows only
  const size_t nz = VS * (size_t)A.n, mz = VS * (size_t)owned_rows(*h.hier.lv[0]);     // the same for Z and x
  const bool dist = h.comm.active();
  Halo& halo = h.hier.lv[0]->halo;
  auto zero_ghost = [&](double* v) { if (n > m) PNP_CUDA(cudaMemsetAsync(v + m, 0, (n - m) * sizeof(double), s)); };
  const int MaxIt = cfg.maxit;
  const int restart_max = cfg.restart < 1 ? 1 : cfg.restart, restart_min = 3, dstep = 3;
  const double cr_max = 0.99, cr_min = 0.174, epsmac = 1e-20;
  Krylov& K = h.kry;
  // sweep-space vectors are padded to VS doubles per row; the pad entries must be (and stay) zero
  double* V; double* Z;
  if (h.share_scratch == 1) {
    // low-memory mode (internal.hpp): the basis lives in the assembly's record buffer, which is dead during the solve
    const size_t vlen = ((size_t)(restart_max + 1) * n + 15) & ~(size_t)15;      // Z stays 128-byte aligned
    const size_t total = vlen + (size_t)restart_max * nz;
    K.V.release(); K.Z.release();
    if (h.cellrec.n < total) h.cellrec.alloc(total, s);
    V = h.cellrec.p; Z = V + vlen;
    PNP_CUDA(cudaMemsetAsync(Z, 0, (size_t)restart_max * nz * sizeof(double), s));
  } else {
    K.V.alloc((size_t)(restart_max + 1) * n, s);
    { const double* old = K.Z.p; K.Z.alloc((size_t)restart_max * nz, s); if (K.Z.p != old) K.Z.zero(); }
    V = K.V.p; Z = K.Z.p;
  }
  auto Vi = [&](int i) { return V + (size_t)i * n; };
  auto Zi = [&](int i) { return Z + (size_t)i * nz; };

  std::vector<std::vector<double>> hh(restart_max + 1, std::vector<double>(restart_max, 0.0));
  std::vector<double> c(restart_max), sn(restart_max), rs(restart_max + 1), hcol(restart_max + 2), h2(restart_max + 2),
      neg(restart_max + 1);

  const double b_norm = vec_norm2(h, b, m);
  double r_norm, r_norm_old;
  if (zero_guess) {                 // r0 = b - A*0
    vec_copy(Vi(0), b, n, s);
    r_norm = b_norm;
  } else {
    if (dist) halo_exchange(h, halo, x, true);
    bsrt_residual(A, b, x, Vi(0), s, KS, KS);
    r_norm = vec_norm2(h, Vi(0), m);
  }
  const double den = b_norm > 0.0 ? b_norm : r_norm;
  const double epsilon = cfg.tol * den;
  h.stats.krylov_its = 0;
  // a NaN / Inf rhs or Jacobian (diverged Newton iterate) must not read as "converged in 0 iterations":
  // every comparison with NaN is false, so the finiteness test comes first
  if (!std::isfinite(r_norm) || !std::isfinite(b_norm)) return ERROR_SOLVER_EXIT;
  if (!(r_norm > epsilon)) return 0;

  int iter = 0, Restart = restart_max;
  double cr = 1.0;
  bool converged = false;
  while (iter < MaxIt) {
    rs[0] = r_norm; r_norm_old = r_norm;
    if (r_norm == 0.0) { h.stats.krylov_its = iter; return iter; }
    if (cr > cr_max || iter == 0) Restart = restart_max;
    else if (cr < cr_min) { /* keep */ }
    else { if (Restart - dstep > restart_min) Restart -= dstep; else Restart = restart_max; }
    vec_scale(1.0 / r_norm, Vi(0), m, s);
    int i = 0;
    while (i < Restart && iter < MaxIt) {
      ++i; ++iter;
      bool have_Az = false;      // the cycle's last backward sweep can deliver A z for half a matrix pass
      {
        PhaseTimer t(h, h.stats.ms_vcycle);
        zero_ghost(Vi(i - 1));     // the preconditioner is local to the rank (block Jacobi across GPUs)
        if (cfg.use_amg) have_Az = amg_vcycle(h, cfg, Vi(i - 1), Zi(i - 1), Vi(i), KS, KS);
        else if (cfg.use_ilu) ilu_apply(h, Vi(i - 1), KS, Zi(i - 1), VS, true);
        else vec_restride(Vi(i - 1), KS, Zi(i - 1), VS, (size_t)A.n, s);
      }
      if (!have_Az) { PhaseTimer t(h, h.stats.ms_spmv); if (dist) halo_exchange(h, halo, Zi(i - 1), true); bsrt_spmv(A, Zi(i - 1), Vi(i), s, KS); }
      double tnorm;
      {
        PhaseTimer t(h, h.stats.ms_ortho);
        multi_dot(h, V, n, i, Vi(i), m, hcol.data());
        const double before = std::sqrt(hcol[i]);
        for (int j = 0; j < i; ++j) neg[j] = -hcol[j];
        double after2 = 0.0;
        multi_axpy(h, V, n, i, neg.data(), Vi(i), m, &after2);
        if (std::sqrt(after2) < 0.1 * before) {   // DGKS: orthogonalise once more
          multi_dot(h, V, n, i, Vi(i), m, h2.data());
          for (int j = 0; j < i; ++j) { neg[j] = -h2[j]; hcol[j] += h2[j]; }
          multi_axpy(h, V, n, i, neg.data(), Vi(i), m, &after2);
        }
        tnorm = std::sqrt(after2);
        for (int j = 0; j < i; ++j) hh[j][i - 1] = hcol[j];
        hh[i][i - 1] = tnorm;
        if (tnorm != 0.0) vec_scale(1.0 / tnorm, Vi(i), m, s);
      }
      if (!std::isfinite(tnorm)) { h.stats.krylov_its = iter; return ERROR_SOLVER_EXIT; }
      for (int j = 1; j < i; ++j) {
        const double t = hh[j - 1][i - 1];
        hh[j - 1][i - 1] = sn[j - 1] * hh[j][i - 1] + c[j - 1] * t;
        hh[j][i - 1] = -sn[j - 1] * t + c[j - 1] * hh[j][i - 1];
      }
      double gamma = std::sqrt(hh[i][i - 1] * hh[i][i - 1] + hh[i - 1][i - 1] * hh[i - 1][i - 1]);
      if (gamma == 0.0) gamma = epsmac;
      c[i - 1] = hh[i - 1][i - 1] / gamma;
      sn[i - 1] = hh[i][i - 1] / gamma;
      rs[i] = -sn[i - 1] * rs[i - 1];
      rs[i - 1] = c[i - 1] * rs[i - 1];
      hh[i - 1][i - 1] = sn[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
      r_norm = std::fabs(rs[i]);
      if (cfg.print_level > 1) printf("  [pnpb200 fgmres] it %3d  relres %.6e\n", iter, r_norm / den);
      if (r_norm <= epsilon) break;
    }
    // solve the small triangular system, x += Z y
    rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
    for (int k = i - 2; k >= 0; --k) {
      double t = 0.0;
      for (int j = k + 1; j < i; ++j) t -= hh[k][j] * rs[j];
      t += rs[k];
      rs[k] = t / hh[k][k];
    }
    multi_axpy(h, Z, nz, i, rs.data(), x, mz, nullptr);
    // true residual for the convergence confirmation / next cycle
    const bool est_conv = r_norm <= epsilon;
    if (dist) halo_exchange(h, halo, x, true);
    bsrt_residual(A, b, x, Vi(0), s, KS, KS);
    r_norm = vec_norm2(h, Vi(0), m);
    if (est_conv && r_norm <= epsilon) { converged = true; break; }
    if (!std::isfinite(r_norm)) { h.stats.krylov_its = iter; return ERROR_SOLVER_EXIT; }
    if (r_norm <= epsilon) { converged = true; break; }
    cr = r_norm / r_norm_old;
  }
  h.stats.krylov_its = iter;
  if (!converged && iter >= MaxIt) return ERROR_SOLVER_MAXIT;
  return iter;
}
} // namespace

// b, x in row-id order (the C ABI / assembly side); the iteration runs in sweep order
int fgmres_solve(Handle& h, const SolverConfig& cfg, const double* b, double* x, bool zero_guess)
{
  cudaStream_t s = h.stream;
  const BsrT& A = h.hier.lv[0]->A;
  const size_t n = VS * (size_t)A.n;
  Krylov& K = h.kry;
  K.bs.alloc(3 * (size_t)A.n, s); K.xs.alloc(n, s);       // rhs packed like the Krylov basis, iterate padded
  vec_to_sweep(A, b, K.bs.p, s, 3);
  if (zero_guess) vec_set(K.xs.p, 0.0, n, s); else vec_to_sweep(A, x, K.xs.p, s);
  const int rc = fgmres_sweep(h, cfg, K.bs.p, K.xs.p, zero_guess);
  vec_from_sweep(A, K.xs.p, x, s);      // the last iterate, also on failure
  return rc;
}

} // namespace pnpb200

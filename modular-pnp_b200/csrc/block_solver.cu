// pnp-b200 — PNP-Stokes 2x2 block solver. Stands in for fasp4ns'
// fasp_solver_bdcsr_krylov_pnp_stokes [EXT, source not in the reference tree] as it is called at
// benchmarks/pnp_stokes/linear_pnp_ns.cpp:183-193 on the block matrix built at :150-157:
//
//        [ A_pp  A_ps ] [x_p]   [b_p]      p = PNP dofs (3 per vertex, interleaved, :75-79)
//        [ A_sp  A_ss ] [x_s] = [b_s]      s = Stokes dofs: n_vel velocity dofs, then n_pres pressure dofs
//
// Outer iteration: right-preconditioned flexible GMRES (bcsr.dat: solver_type 6, restart 10) with
// the block UPPER-triangular preconditioner of bcsr.dat's precond_type 33:
//        z_s = S(r_s),   z_p = P(r_p - A_ps z_s)
// P = the BSR(3) AMG-FGMRES solve of the PNP block (the hot path of this library, unchanged),
// S = FGMRES on the Stokes block (ns.dat: solver_type 5) preconditioned by the block lower-
// triangular factor  z_u = AMG-cycle(A_uu) r_u,  z_q = Sd^{-1}(r_q - A_qu z_u)  with
// Sd = diag(A_qq - A_qu diag(A_uu)^{-1} A_uq), the diagonal of the SIMPLE Schur complement (for the
// reference's RT1 x DG0 pair the exact Schur complement is spectrally equivalent to the diagonal
// DG0 mass matrix, which Sd reproduces up to a constant). The velocity block runs through the same
// BSR(3) unsmoothed-aggregation hierarchy as the PNP block, its scalar dofs grouped in threes.
// Every stage is a device kernel; the host only sequences launches and owns the small Hessenberg
// recurrences, exactly as in fgmres.cu. Parity: unpinned (no source, no golden data for fasp4ns);
// the tests check the solution against a sparse direct solve of the same block system.
#include "internal.hpp"
#include <cmath>
#include <functional>

namespace pnpb200 {

namespace {

// G lanes per row, rows [r0, r1), columns restricted to [c0, c1)
template <int G>
__global__ void __launch_bounds__(256)
csr_spmv_kernel(int r0, int r1, int c0, int c1, const int* __restrict__ ia, const int* __restrict__ ja,
                const double* __restrict__ val, double alpha, const double* __restrict__ x, double beta, double* __restrict__ y)
{
  const int t = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) / G);
  const int lane = threadIdx.x % G;
  const int row = r0 + t;
  double a = 0.0;
  if (row < r1) {
    const int e = ia[row + 1];
    for (int p = ia[row] + lane; p < e; p += G) {
      const int j = ja[p];
      if (j >= c0 && j < c1) a += val[p] * x[j - c0];
    }
  }
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0 && row < r1) y[t] = beta == 0.0 ? alpha * a : beta * y[t] + alpha * a;
}

// inverse diagonal of the SIMPLE Schur complement of the Stokes block A = [A_uu A_uq; A_qu A_qq]
// (velocity dofs first): sinv[i] = 1 / (a_qq(i,i) - sum_j a_qu(i,j) a_uq(j,i) / a_uu(j,j)), one
// thread per pressure row (rows of an RT1 x DG0 divergence have <= 4 entries)
__global__ void schur_diag_kernel(int nu, int npres, const int* __restrict__ ia, const int* __restrict__ ja,
                                  const double* __restrict__ val, double* __restrict__ sinv, int* __restrict__ bad)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npres) return;
  const int row = nu + i;
  double sd = 0.0;
  for (int p = ia[row]; p < ia[row + 1]; ++p) {
    const int j = ja[p];
    if (j == row) { sd += val[p]; continue; }
    if (j >= nu) continue;
    double ajj = 0.0, aji = 0.0;
    for (int q = ia[j]; q < ia[j + 1]; ++q) {
      const int c = ja[q];
      if (c == j) ajj = val[q]; else if (c == row) aji = val[q];
    }
    if (ajj != 0.0) sd -= val[p] * aji / ajj;
  }
  if (sd == 0.0 || !isfinite(sd)) { atomicExch(bad, 1); sinv[i] = 0.0; } else sinv[i] = 1.0 / sd;
}

__global__ void scale_by_kernel(int n, const double* __restrict__ d, const double* __restrict__ in, double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = d[i] * in[i];
}

// ordering between the streams of the handles involved: work queued on `to` after this call sees
// everything queued on `from` before it
struct StreamLink {
  cudaEvent_t ev = nullptr;
  StreamLink() { PNP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
  ~StreamLink() { if (ev) cudaEventDestroy(ev); }
  void pass(cudaStream_t from, cudaStream_t to)
  {
    if (from == to) return;
    PNP_CUDA(cudaEventRecord(ev, from));
    PNP_CUDA(cudaStreamWaitEvent(to, ev, 0));
  }
};

using Op = std::function<void(const double*, double*)>;

// Flexible GMRES on plain device vectors with operator / preconditioner callbacks: the control flow
// of fgmres.cu (FASP's pvfgmres: adaptive restart, Givens recurrences, ||r|| <= tol ||b||,
// true residual at every restart), fused classical Gram-Schmidt with DGKS re-orthogonalisation.
// Returns the iteration count or ERROR_SOLVER_MAXIT / ERROR_SOLVER_EXIT; relres_out = final
// ||b - A x|| / ||b|| (true residual).
int fgmres_ops(Handle& h, size_t n, const Op& A, const Op& M, const double* b, double* x, bool zero_guess,
               const SolverConfig& cfg, const char* tag, double* relres_out)
{
  cudaStream_t s = h.stream;
  const int MaxIt = cfg.maxit > 0 ? cfg.maxit : 100;
  int restart_max = cfg.restart < 1 ? 1 : (cfg.restart > 31 ? 31 : cfg.restart);
  if (restart_max > MaxIt) restart_max = MaxIt;
  const int restart_min = 3, dstep = 3;
  const double cr_max = 0.99, cr_min = 0.174, epsmac = 1e-20;
  DevBuf<double> Vb, Zb;
  Vb.alloc((size_t)(restart_max + 1) * n, s); Zb.alloc((size_t)restart_max * n, s);
  double* V = Vb.p; double* Z = Zb.p;
  auto Vi = [&](int i) { return V + (size_t)i * n; };
  auto Zi = [&](int i) { return Z + (size_t)i * n; };
  std::vector<std::vector<double>> hh(restart_max + 1, std::vector<double>(restart_max, 0.0));
  std::vector<double> c(restart_max), sn(restart_max), rs(restart_max + 1), hcol(restart_max + 2), h2(restart_max + 2),
      neg(restart_max + 1);
  auto residual = [&](double* r) {            // r = b - A x
    A(x, r);
    vec_scale(-1.0, r, n, s);
    vec_axpy(1.0, b, r, n, s);
  };
  const double b_norm = vec_norm2(h, b, n);
  double r_norm, r_norm_old;
  if (zero_guess) { vec_copy(Vi(0), b, n, s); r_norm = b_norm; }
  else { residual(Vi(0)); r_norm = vec_norm2(h, Vi(0), n); }
  const double den = b_norm > 0.0 ? b_norm : r_norm;
  const double epsilon = cfg.tol * den;
  if (relres_out) *relres_out = den > 0.0 ? r_norm / den : 0.0;
  if (!std::isfinite(r_norm) || !std::isfinite(b_norm)) return ERROR_SOLVER_EXIT;   // before the NaN-blind comparison
  if (!(r_norm > epsilon)) return 0;
  int iter = 0, Restart = restart_max;
  double cr = 1.0;
  bool converged = false;
  while (iter < MaxIt) {
    rs[0] = r_norm; r_norm_old = r_norm;
    if (r_norm == 0.0) break;
    if (cr > cr_max || iter == 0) Restart = restart_max;
    else if (cr < cr_min) { /* keep */ }
    else { if (Restart - dstep > restart_min) Restart -= dstep; else Restart = restart_max; }
    vec_scale(1.0 / r_norm, Vi(0), n, s);
    int i = 0;
    while (i < Restart && iter < MaxIt) {
      ++i; ++iter;
      M(Vi(i - 1), Zi(i - 1));
      A(Zi(i - 1), Vi(i));
      multi_dot(h, V, n, i, Vi(i), n, hcol.data());
      const double before = std::sqrt(hcol[i]);
      for (int j = 0; j < i; ++j) neg[j] = -hcol[j];
      double after2 = 0.0;
      multi_axpy(h, V, n, i, neg.data(), Vi(i), n, &after2);
      if (std::sqrt(after2) < 0.1 * before) {
        multi_dot(h, V, n, i, Vi(i), n, h2.data());
        for (int j = 0; j < i; ++j) { neg[j] = -h2[j]; hcol[j] += h2[j]; }
        multi_axpy(h, V, n, i, neg.data(), Vi(i), n, &after2);
      }
      const double tnorm = std::sqrt(after2);
      if (!std::isfinite(tnorm)) return ERROR_SOLVER_EXIT;
      for (int j = 0; j < i; ++j) hh[j][i - 1] = hcol[j];
      hh[i][i - 1] = tnorm;
      if (tnorm != 0.0) vec_scale(1.0 / tnorm, Vi(i), n, s);
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
      if (cfg.print_level > 1) printf("  [pnpb200 %s] it %3d  relres %.6e\n", tag, iter, r_norm / den);
      if (r_norm <= epsilon) break;
    }
    rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
    for (int k = i - 2; k >= 0; --k) {
      double t = 0.0;
      for (int j = k + 1; j < i; ++j) t -= hh[k][j] * rs[j];
      t += rs[k];
      rs[k] = t / hh[k][k];
    }
    multi_axpy(h, Z, n, i, rs.data(), x, n, nullptr);
    residual(Vi(0));
    r_norm = vec_norm2(h, Vi(0), n);
    if (relres_out) *relres_out = r_norm / den;
    if (!std::isfinite(r_norm)) return ERROR_SOLVER_EXIT;
    if (r_norm <= epsilon) { converged = true; break; }
    cr = r_norm / r_norm_old;
  }
  if (!converged && iter >= MaxIt) return ERROR_SOLVER_MAXIT;
  return iter;
}

} // namespace

void csr_upload(const dCSRmat* A, DevCsr& D, cudaStream_t s)
{
  if (!A || A->row < 0 || A->col < 0 || (A->row > 0 && !A->IA)) throw Error(ERROR_INPUT_PAR, "csr_upload: bad dCSRmat");
  D.nrow = A->row; D.ncol = A->col;
  D.nnz = A->row > 0 ? A->IA[A->row] - A->IA[0] : 0;
  if (D.nnz < 0 || (D.nnz > 0 && (!A->JA || !A->val))) throw Error(ERROR_INPUT_PAR, "csr_upload: bad dCSRmat arrays");
  D.ia.alloc((size_t)D.nrow + 1, s);
  if (A->row > 0 && A->IA[0] != 0) {           // 1-based or offset IA is not what the reference produces
    throw Error(ERROR_DATA_STRUCTURE, "csr_upload: IA[0] must be 0");
  }
  if (A->row > 0) D.ia.upload(A->IA, (size_t)D.nrow + 1); else D.ia.zero();
  D.ja.alloc(D.nnz ? D.nnz : 1, s); D.val.alloc(D.nnz ? D.nnz : 1, s);
  if (D.nnz > 0) { D.ja.upload(A->JA, D.nnz); D.val.upload(A->val, D.nnz); }
}

void csr_spmv(const DevCsr& A, int r0, int r1, int c0, int c1, double alpha, const double* x, double beta, double* y,
              cudaStream_t s)
{
  const int rows = r1 - r0;
  if (rows <= 0) return;
  const double avg = A.nrow > 0 ? (double)A.nnz / A.nrow : 0.0;
  if (avg > 24.0) csr_spmv_kernel<32><<<cdiv((size_t)rows * 32, 256), 256, 0, s>>>(r0, r1, c0, c1, A.ia.p, A.ja.p, A.val.p, alpha, x, beta, y);
  else if (avg > 6.0) csr_spmv_kernel<8><<<cdiv((size_t)rows * 8, 256), 256, 0, s>>>(r0, r1, c0, c1, A.ia.p, A.ja.p, A.val.p, alpha, x, beta, y);
  else csr_spmv_kernel<2><<<cdiv((size_t)rows * 2, 256), 256, 0, s>>>(r0, r1, c0, c1, A.ia.p, A.ja.p, A.val.p, alpha, x, beta, y);
  PNP_COUNT_LAUNCH();
}

int block_pnp_stokes_solve(BlockSystem& S, const double* b, double* x, BlockStats* st)
{
  Handle& ho = *S.outer; Handle& hp = *S.pnp; Handle& hv = *S.vel;
  cudaStream_t so = ho.stream, sp = hp.stream, sv = hv.stream;
  const int np = S.np, ns = S.ns, nu = S.nu, nq = S.npres;
  const size_t n = (size_t)np + ns;
  StreamLink link;
  BlockStats stats;

  // hierarchies of the two diagonal blocks that have one (each on its handle's stream)
  hp.last_cfg = S.cfg_pnp; hv.last_cfg = S.cfg_vel;
  if (S.cfg_pnp.use_amg) amg_setup(hp, S.cfg_pnp);
  amg_setup(hv, S.cfg_vel);
  // diagonal of the SIMPLE Schur complement
  S.sinv.alloc(nq ? nq : 1, so);
  {
    DevBuf<int> bad; bad.alloc(1, so); bad.zero();
    if (nq > 0) { schur_diag_kernel<<<cdiv(nq, 256), 256, 0, so>>>(nu, nq, S.Ass.ia.p, S.Ass.ja.p, S.Ass.val.p, S.sinv.p, bad.p); PNP_COUNT_LAUNCH(); }
    int hb = 0; bad.download(&hb, 1);
    if (hb) throw Error(ERROR_DATA_STRUCTURE, "Stokes block: zero diagonal in the pressure Schur complement (need A_qq or A_qu diag(A_uu)^-1 A_uq)");
  }
  link.pass(sp, so); link.pass(sv, so);

  const BsrT& Av = hv.hier.lv[0]->A;
  const size_t nvp = 3 * (size_t)Av.n;          // padded velocity length
  DevBuf<double> vr, vz, vrs, vzs, tq, tp, zp;
  vr.alloc(nvp, sv); vz.alloc(nvp, sv); vrs.alloc(VS * (size_t)Av.n, sv); vzs.alloc(VS * (size_t)Av.n, sv);   // vrs, vzs: sweep space
  vzs.zero();
  vr.zero();                                     // the padding stays zero
  tq.alloc(nq ? nq : 1, so); tp.alloc(np, so); zp.alloc(np, sp);
  link.pass(sv, so);

  // ---- Stokes block: operator and block lower-triangular preconditioner
  Op A_ss = [&](const double* xs, double* ys) { csr_spmv(S.Ass, 0, ns, 0, ns, 1.0, xs, 0.0, ys, so); };
  Op M_ss = [&](const double* r, double* z) {
    // z_u = one AMG cycle of the velocity hierarchy (sweep space inside), on the velocity handle's stream
    link.pass(so, sv);
    PNP_CUDA(cudaMemcpyAsync(vr.p, r, (size_t)nu * sizeof(double), cudaMemcpyDeviceToDevice, sv));
    if (S.vel_krylov) {
      // ns.dat: solver_type_v = vFGMRES, precond_type_v = AMG, itsolver_tol_v — the div-div part of the RT1 velocity
      // block has a large near-kernel that one cycle of an aggregation hierarchy does not resolve; with an accurate
      // velocity solve the outer Stokes iteration count is mesh independent (profiles/r02_stokes_precond.txt)
      const int vits = fgmres_solve(hv, S.cfg_vel, vr.p, vz.p, true);
      ++stats.vel_solves;
      if (vits > 0) stats.vel_its += vits;
      else if (vits == ERROR_SOLVER_MAXIT) stats.vel_its += S.cfg_vel.maxit;
      else if (vits < 0) throw Error(vits, "velocity block solve failed");
    } else {
      vec_to_sweep(Av, vr.p, vrs.p, sv);
      amg_vcycle(hv, S.cfg_vel, vrs.p, vzs.p);
      vec_from_sweep(Av, vzs.p, vz.p, sv);
    }
    PNP_CUDA(cudaMemcpyAsync(z, vz.p, (size_t)nu * sizeof(double), cudaMemcpyDeviceToDevice, sv));
    link.pass(sv, so);
    if (nq > 0) {
      // z_q = Sd^{-1} (r_q - A_qu z_u)
      PNP_CUDA(cudaMemcpyAsync(tq.p, r + nu, (size_t)nq * sizeof(double), cudaMemcpyDeviceToDevice, so));
      csr_spmv(S.Ass, nu, ns, 0, nu, -1.0, z, 1.0, tq.p, so);
      scale_by_kernel<<<cdiv(nq, 256), 256, 0, so>>>(nq, S.sinv.p, tq.p, z + nu); PNP_COUNT_LAUNCH();
    }
  };
  auto stokes_solve = [&](const double* r, double* z) {
    vec_set(z, 0.0, ns, so);
    double rr = 0.0;
    const int its = fgmres_ops(ho, ns, A_ss, M_ss, r, z, true, S.cfg_ns, "stokes", &rr);
    ++stats.stokes_solves;
    if (its > 0) stats.stokes_its += its;
    else if (its == ERROR_SOLVER_MAXIT) stats.stokes_its += S.cfg_ns.maxit;      // inexact solve: keep the last iterate
    else if (its < 0) throw Error(its, "Stokes block solve failed");
  };
  // ---- PNP block: the library's BSR(3) AMG-FGMRES (row-id vectors on the PNP handle's stream)
  auto pnp_solve = [&](const double* r, double* z) {
    link.pass(so, sp);
    const int its = fgmres_solve(hp, S.cfg_pnp, r, zp.p, true);
    PNP_CUDA(cudaMemcpyAsync(z, zp.p, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice, sp));
    link.pass(sp, so);
    ++stats.pnp_solves;
    if (its > 0) stats.pnp_its += its;
    else if (its == ERROR_SOLVER_MAXIT) stats.pnp_its += S.cfg_pnp.maxit;
    else if (its < 0) throw Error(its, "PNP block solve failed");
  };
  // ---- the coupled system
  Op A_full = [&](const double* xx, double* yy) {
    csr_spmv(S.App, 0, np, 0, np, 1.0, xx, 0.0, yy, so);
    csr_spmv(S.Aps, 0, np, 0, ns, 1.0, xx + np, 1.0, yy, so);
    csr_spmv(S.Asp, 0, ns, 0, np, 1.0, xx, 0.0, yy + np, so);
    csr_spmv(S.Ass, 0, ns, 0, ns, 1.0, xx + np, 1.0, yy + np, so);
  };
  Op M_upper = [&](const double* r, double* z) {
    stokes_solve(r + np, z + np);
    PNP_CUDA(cudaMemcpyAsync(tp.p, r, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice, so));
    csr_spmv(S.Aps, 0, np, 0, ns, -1.0, z + np, 1.0, tp.p, so);
    pnp_solve(tp.p, z);
  };
  double relres = 0.0;
  const int its = fgmres_ops(ho, n, A_full, M_upper, b, x, false, S.cfg_outer, "pnp-stokes", &relres);
  stats.outer_its = its > 0 ? its : (its == ERROR_SOLVER_MAXIT ? S.cfg_outer.maxit : 0);
  stats.relres = relres;
  if (st) *st = stats;
  PNP_CUDA(cudaStreamSynchronize(so));
  PNP_CUDA(cudaStreamSynchronize(sp));
  PNP_CUDA(cudaStreamSynchronize(sv));
  return its;
}

} // namespace pnpb200

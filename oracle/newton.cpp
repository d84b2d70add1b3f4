// TEST INFRASTRUCTURE ONLY (oracle/) — see pnp_oracle.h.
// Newton driver mirroring benchmarks/pnp_exact_solution/main.cpp:293-356 with
// Newton_Status (src/newton_status.cpp:11-74) and Linear_PNP::fasp_solve
// (benchmarks/pnp_exact_solution/linear_pnp.cpp:96-132). Like the reference it rebuilds the
// matrix, the BSR copy and the AMG hierarchy every step and assembles the residual three
// times per step (src/pde.cpp:546, main.cpp:342-343).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "pnp_oracle.h"

namespace {
struct NewtonStatus {   // src/newton_status.cpp
  size_t max_iterations, iteration = 1;
  double initial_residual, rel_tol, max_tol, relative_residual = 1.0, max_residual, residual = 0;
  NewtonStatus(size_t mx, double r0, double rt, double mt)
    : max_iterations(mx), initial_residual(r0), rel_tol(rt), max_tol(mt), max_residual(mt + 1) {}
  bool needs_to_iterate() const {
    return iteration < max_iterations + 1 && (relative_residual > rel_tol && max_residual > max_tol);
  }
  bool converged() const {
    return !(iteration > max_iterations) && (relative_residual < rel_tol || max_residual < max_tol);
  }
};
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" int orc_newton_exact(int nx, int ny, int nz, double lx, double ly, double lz,
                                const orc_solver_param* p, int max_newton, double rel_tol,
                                double max_tol, int use_eafe, double* uu_out, orc_newton_result* res)
{
  int nv, nc;
  orc_box_mesh_counts(nx, ny, nz, &nv, &nc);
  const size_t N = 3 * (size_t)nv;
  std::vector<double> xyz(3 * (size_t)nv);
  std::vector<int> cells(4 * (size_t)nc);
  orc_box_mesh(nx, ny, nz, lx, ly, lz, xyz.data(), cells.data());
  std::vector<double> eps(nv), fixed(nv), diff(N), val(N), reac(N), uu(N);
  orc_exact_problem(nv, xyz.data(), eps.data(), fixed.data(), diff.data(), val.data(), reac.data(), uu.data(), nullptr);
  std::vector<unsigned char> vflag(nv), bc(N);
  orc_dirichlet_vertices(nv, xyz.data(), nc, cells.data(), 0, vflag.data());
  for (int v = 0; v < nv; ++v) bc[3 * v] = bc[3 * v + 1] = bc[3 * v + 2] = vflag[v];   // main.cpp:293 components {0,0,0}
  std::vector<int> IA(nv + 1);
  const int nnzb = orc_block_pattern_count(nv, nc, cells.data(), IA.data());
  std::vector<int> JA(nnzb);
  orc_block_pattern_fill(nv, nc, cells.data(), IA.data(), JA.data());
  std::vector<double> A(9 * (size_t)nnzb), b(N), dx(N);

  memset(res, 0, sizeof(*res));
  double l2, linf, t0;
  t0 = now();
  orc_residual(nv, xyz.data(), nc, cells.data(), uu.data(), eps.data(), fixed.data(), diff.data(), val.data(), reac.data(), bc.data(), b.data(), &l2, nullptr);
  orc_residual(nv, xyz.data(), nc, cells.data(), uu.data(), eps.data(), fixed.data(), diff.data(), val.data(), reac.data(), bc.data(), b.data(), nullptr, &linf);
  res->t_residual += now() - t0;
  res->initial_residual = l2; res->initial_max_residual = linf;
  NewtonStatus newton(max_newton, l2, rel_tol, max_tol);
  newton.max_residual = linf;
  int step = 0;
  while (newton.needs_to_iterate()) {
    t0 = now();
    orc_assemble(nv, xyz.data(), nc, cells.data(), IA.data(), JA.data(), uu.data(), eps.data(), fixed.data(),
                 diff.data(), val.data(), reac.data(), bc.data(), use_eafe, A.data(), b.data());
    res->t_assemble += now() - t0;
    t0 = now();
    std::fill(dx.begin(), dx.end(), 0.0);                                   // linear_pnp.cpp:93
    const int status = orc_solver_dbsr_krylov_amg(nv, IA.data(), JA.data(), A.data(), b.data(), dx.data(), p);
    res->t_solve += now() - t0;
    if (status >= 0) for (size_t i = 0; i < N; ++i) uu[i] += dx[i];        // linear_pnp.cpp:109-127
    t0 = now();
    orc_residual(nv, xyz.data(), nc, cells.data(), uu.data(), eps.data(), fixed.data(), diff.data(), val.data(), reac.data(), bc.data(), b.data(), &l2, nullptr);
    orc_residual(nv, xyz.data(), nc, cells.data(), uu.data(), eps.data(), fixed.data(), diff.data(), val.data(), reac.data(), bc.data(), b.data(), nullptr, &linf);
    res->t_residual += now() - t0;
    newton.residual = l2; newton.relative_residual = l2 / newton.initial_residual; newton.max_residual = linf;
    newton.iteration++;
    if (step < 32) { res->rel_res[step] = newton.relative_residual; res->max_res[step] = linf; res->krylov_its[step] = status; }
    ++step;
    if (p->print_level > 0)
      printf("[oracle newton] it %d krylov %d relres %.5e maxres %.5e\n", step, status, newton.relative_residual, linf);
  }
  res->newton_its = step;
  res->converged = newton.converged() ? 1 : 0;
  if (uu_out) memcpy(uu_out, uu.data(), N * sizeof(double));
  return 0;
}

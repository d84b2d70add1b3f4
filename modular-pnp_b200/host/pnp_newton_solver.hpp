// pnp-b200 host side — the diode benchmark's damped Newton loop
// (benchmarks/pnp_diode/pnp_newton_solver.h:171-284) on the device-resident fast path of the C ABI.
// Every iterate the reference loop evaluates is base + alpha * update (the full step :200, the
// growth-limited step :229-252, the <= 50 backtracking steps :256-268), so the solution never returns
// to the host: pnpb200_solve_update / pnpb200_update_ranges / pnpb200_trial_residual.
#pragma once
#include <cmath>
#include <cstdio>
#include <vector>
#include "pnpb200/pnpb200.h"
#include "newton_status.hpp"

struct Newton_Step_Record { int krylov_status; double alpha, relative_residual, max_residual; int backtracks; };

// `h` carries mesh, coefficients (pnpb200_set_form_variant for the diode forms), Dirichlet dofs and the
// initial guess. Returns true when Newton_Status::converged().
inline bool pnp_damped_newton(pnpb200_handle* h, itsolver_param* itsolver, AMG_param* amg, std::size_t max_newton,
                              double relative_residual_tol, double max_residual_tol,
                              std::vector<Newton_Step_Record>* history = nullptr)
{
  const double dof_size = 3.0 * pnpb200_num_vertices(h);
  double l2 = 0.0, mx = 0.0;
  if (pnpb200_residual(h, &l2, &mx) != FASP_SUCCESS) return false;
  Newton_Status newton(max_newton, l2 / dof_size, relative_residual_tol, max_residual_tol);
  newton.update_max_residual(mx);
  const double increase_tolerance = 1e-1;
  unsigned fasp_fail_count = 0;
  while (newton.needs_to_iterate()) {
    if (pnpb200_residual(h, &l2, &mx) != FASP_SUCCESS) return false;
    const double previous_residual = l2 / dof_size;
    int status = 0;
    if (pnpb200_solve_update(h, itsolver, amg, &status) != FASP_SUCCESS) return false;
    if (status < 0 && ++fasp_fail_count > 10) { printf("Linear solver has failed more than 10 times...\n"); break; }
    double r[6];
    if (pnpb200_update_ranges(h, r) != FASP_SUCCESS) return false;
    double alpha = 1.0;
    pnpb200_trial_residual(h, 1.0, &l2, &mx);                       // the reference applies the full update first
    double residual_check = l2 / dof_size;
    const double prev_range = r[1] - r[0] + 1e-12;
    if ((r[5] - r[4]) > (1.0 + increase_tolerance) * prev_range) {   // bounded growth of the solution range
      alpha = increase_tolerance * prev_range / (r[3] - r[2] + 1e-12);
      if (alpha * r[3] > alpha * r[2] + increase_tolerance * prev_range) alpha *= (alpha * r[3] - alpha * r[2]) * increase_tolerance * 1e-4;
    }
    int backtrack_count = 0;
    while (backtrack_count < 50 && (residual_check > 1.00001 * previous_residual || std::isnan(residual_check))) {
      pnpb200_trial_residual(h, alpha * std::pow(0.5, ++backtrack_count), &l2, &mx);
      residual_check = l2 / dof_size;
    }
    if (pnpb200_residual(h, &l2, &mx) != FASP_SUCCESS) return false;
    newton.update_residuals(l2 / dof_size, mx);
    newton.update_iteration();
    if (history) history->push_back({status, backtrack_count ? alpha * std::pow(0.5, backtrack_count) : 1.0, newton.relative_residual, mx, backtrack_count});
  }
  return newton.converged();
}

// pnp-b200 host side — Newton convergence bookkeeping with the semantics of the reference's
// Newton_Status (src/newton_status.cpp:11-74): iteration starts at 1; iterate while BOTH the
// relative and the max residual exceed their tolerances; converged iff not too many iterations
// and EITHER residual is below its tolerance. print_status keeps the reference's output line for line, including its
// quirk (src/newton_status.cpp:89-93): the "max residual too large" line is printed under the
// RELATIVE-residual condition. Pinned against the reference's own src/newton_status.cpp compiled where it lies
// (oracle/_ref/libpnpref_host.so; tests/test_host_logic.py, tests/golden/host_logic.json).
#pragma once
#include <cstddef>
#include <cstdio>

class Newton_Status {
public:
  Newton_Status(std::size_t max_iterations_in, double initial_residual_in, double rel_residual_tol_in,
                double max_residual_tol_in)
    : max_iterations(max_iterations_in), iteration(1), initial_residual(initial_residual_in),
      residual(initial_residual_in), relative_residual(1.0), max_residual(max_residual_tol_in + 1),
      rel_residual_tol(rel_residual_tol_in), max_residual_tol(max_residual_tol_in) {}

  void update_iteration() { ++iteration; }
  void update_residuals(double residual_in, double max_residual_in)
  { update_rel_residual(residual_in); update_max_residual(max_residual_in); }
  void update_max_residual(double m) { max_residual = m; }
  void update_rel_residual(double r) { residual = r; relative_residual = r / initial_residual; }

  bool needs_to_iterate() const
  {
    return iteration < max_iterations + 1 && (relative_residual > rel_residual_tol && max_residual > max_residual_tol);
  }
  bool converged() const
  {
    return !(iteration > max_iterations) && (relative_residual < rel_residual_tol || max_residual < max_residual_tol);
  }
  void print_status() const
  {
    printf("Newton solver status:\n");
    if (iteration > max_iterations + 1) printf("\ttoo many iterations\n");
    if (!(relative_residual < rel_residual_tol)) printf("\trelative residual too large : %e >= %e\n", relative_residual, rel_residual_tol);
    if (!(relative_residual < rel_residual_tol)) printf("\tmax residual too large : %e >= %e\n", max_residual, max_residual_tol);   // sic
    printf(converged() ? "\tconverged\n" : "\tnot converged\n");
  }

  std::size_t max_iterations, iteration;
  double initial_residual, residual, relative_residual, max_residual, rel_residual_tol, max_residual_tol;
};

// pnp-b200 host side — driver with the control flow of the reference's
// benchmarks/pnp_exact_solution/main.cpp:138-391 (parameters, problem set-up, Newton loop,
// final status), written against host/linear_pnp.hpp. No DOLFIN, no VTK output; the nodal error
// against the interpolated manufactured solution is printed instead of the L2/H1 functionals
// (out of scope).
//   usage: pnp_exact_solution [domain.dat [bsr.dat]] [--grid nx ny nz] [--lengths lx ly lz]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include "domain.hpp"
#include "linear_pnp.hpp"
#include "newton_status.hpp"

static const double pi = 3.14159265359;          // main.cpp:24
static const double permittivity_const = 1.0, electric_strength = 1.0, ref_concentration = 1.0;

static void exact_solution(double x, double* s)
{
  s[0] = x * electric_strength + std::sin(pi * x);
  s[1] = std::log(ref_concentration) - x * x * electric_strength;
  s[2] = std::log(ref_concentration) + x * x * electric_strength;
}
static void exact_derivative(double x, double* s)
{ s[0] = electric_strength + pi * std::cos(pi * x); s[1] = -2.0 * x * electric_strength; s[2] = 2.0 * x * electric_strength; }
static void exact_second(double x, double* s)
{ s[0] = -(pi * pi) * std::sin(pi * x); s[1] = -2.0 * electric_strength; s[2] = 2.0 * electric_strength; }

int main(int argc, char** argv)
{
  char exe[4096] = {0};
  std::string here = ".";
  if (readlink("/proc/self/exe", exe, sizeof exe - 1) > 0) { here = exe; here = here.substr(0, here.find_last_of('/')); }
  std::string domain_file = here + "/domain.dat", fasp_file = here + "/bsr.dat";
  int grid[3] = {0, 0, 0}; double len[3] = {0, 0, 0};
  int pos = 0;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "--grid") && a + 3 < argc) { for (int d = 0; d < 3; ++d) grid[d] = atoi(argv[++a]); }
    else if (!strcmp(argv[a], "--lengths") && a + 3 < argc) { for (int d = 0; d < 3; ++d) len[d] = atof(argv[++a]); }
    else if (pos == 0) { domain_file = argv[a]; ++pos; }
    else if (pos == 1) { fasp_file = argv[a]; ++pos; }
  }
  printf("\n----------------------------------------------------\n Setting up the linearized PNP problem\n----------------------------------------------------\n\n");
  domain_param domain;
  if (domain_param_input(domain_file.c_str(), &domain) != FASP_SUCCESS) return 1;
  if (grid[0] > 0) { domain.grid_x = grid[0]; domain.grid_y = grid[1]; domain.grid_z = grid[2]; }
  if (len[0] > 0) { domain.length_x = len[0]; domain.length_y = len[1]; domain.length_z = len[2]; }
  pnpb200_handle* mesh = domain_build(domain);
  if (!mesh) { printf("### ERROR: could not build the mesh (no CUDA device?)\n"); return 2; }

  input_param input; itsolver_param itsolver; AMG_param amg; ILU_param ilu;
  fasp_param_input(fasp_file.c_str(), &input);
  fasp_param_init(&input, &itsolver, &amg, &ilu, NULL);

  printf("Construct vector PNP problem\n");
  Linear_PNP pnp_problem(mesh, itsolver, amg);
  const std::vector<double>& xyz = pnp_problem.coordinates();
  const size_t nv = xyz.size() / 3, N = 3 * nv;
  std::vector<double> permittivity(nv), charges(nv), diffusivity(N), reaction(N), valency(N), exact(N);
  for (size_t v = 0; v < nv; ++v) {
    const double x = xyz[3 * v];
    double s[3], ds[3], dds[3];
    exact_solution(x, s); exact_derivative(x, ds); exact_second(x, dds);
    const double diff[3] = {0.0, 1.0, 1.0};                                                   // main.cpp:53-59
    permittivity[v] = permittivity_const;
    charges[v] = -permittivity_const * dds[0] - (std::exp(s[1]) - std::exp(s[2]));           // main.cpp:61-66
    for (int k = 0; k < 3; ++k) { diffusivity[3 * v + k] = diff[k]; exact[3 * v + k] = s[k]; }
    valency[3 * v] = 0.0; valency[3 * v + 1] = 1.0; valency[3 * v + 2] = -1.0;               // main.cpp:128-134
    reaction[3 * v] = 0.0;                                                                   // main.cpp:68-78
    reaction[3 * v + 1] = -diff[1] * std::exp(s[1]) * (ds[1] * (ds[1] + ds[0]) + (dds[1] + dds[0]));
    reaction[3 * v + 2] = -diff[2] * std::exp(s[2]) * (ds[2] * (ds[2] - ds[0]) + (dds[2] - dds[0]));
  }
  pnp_problem.set_coefficients(permittivity, charges, diffusivity, valency, reaction);

  printf("Record interpolant for given Dirichlet BCs (initial guess for solution)\n");
  double left[3], right[3];
  exact_solution(-1.0, left); exact_solution(+1.0, right);
  pnp_problem.set_DirichletBC({0, 0, 0}, {{left[0], right[0]}, {left[1], right[1]}, {left[2], right[2]}});

  printf("Initializing nonlinear solver\n");
  const std::size_t max_newton = 15;
  const double max_residual_tol = 1.0e-10, relative_residual_tol = 1.0e-8;
  const double initial_residual = pnp_problem.compute_residual("l2");
  const double initial_max_residual = pnp_problem.compute_residual("max");
  Newton_Status newton(max_newton, initial_residual, relative_residual_tol, max_residual_tol);
  printf("\tinitial residual :     %10.5e\n\tinitial max residual : %10.5e\n\n", newton.initial_residual, initial_max_residual);
  newton.update_max_residual(initial_max_residual);
  while (newton.needs_to_iterate()) {
    printf("Solving for Newton iterate %lu \n", (unsigned long)newton.iteration);
    pnp_problem.use_eafe();
    pnp_problem.fasp_solve();
    printf("Newton measurements for iteration :\n");
    newton.update_residuals(pnp_problem.last_residual("l2"), pnp_problem.last_residual("max"));
    newton.update_iteration();
    printf("\tkrylov iterations : %d\n\tmaximum residual :  %10.5e\n\trelative residual : %10.5e\n\n", pnp_problem.last_status(),
           newton.max_residual, newton.relative_residual);
  }
  if (newton.converged()) printf("Solver succeeded!\n"); else newton.print_status();
  std::vector<double> uu = pnp_problem.get_solution();
  double err = 0.0;
  for (size_t i = 0; i < N; ++i) err = std::fmax(err, std::fabs(uu[i] - exact[i]));
  double l2_error = 0.0, h1_error = 0.0;
  pnp_problem.compute_error(exact, l2_error, h1_error);     // main.cpp:368-390
  printf("\nMeasuring error of computed solution wrt interpolant\n\tL2 error: %e\n\tH1 error: %e\n\tmax nodal error: %e\n",
         l2_error, h1_error, err);
  printf("NEWTON_ITERATIONS %lu CONVERGED %d RELRES %.6e\n", (unsigned long)(newton.iteration - 1), newton.converged() ? 1 : 0,
         newton.relative_residual);
  printf("\nSolver exiting\n");
  return newton.converged() ? 0 : 3;
}

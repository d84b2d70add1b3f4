// pnp-b200 host side — driver with the control flow of the reference's benchmarks/pnp_stokes/main.cpp:34-241
// (domain, three FASP parameter files, problem data, Newton loop with Newton_Status), written against
// host/linear_pnp_ns.hpp with the mixed P1^3 x RT1 x DG0 system assembled on the device (pnpb200/pnpb200_ns.h).
// No DOLFIN, no VTK output. Boundary conditions: the PNP components on the two x faces (linear_pnp_ns.cpp:287-304);
// the reference's velocity / pressure conditions select no RT1 / DG0 dof (a corner point, a facet marker on DG0), which
// leaves constant velocities in the kernel of the Newton matrix — this driver closes the system with u.n = 0 on the
// boundary facets and a pinned pressure in cell 0, and starts from zero velocity (README "PNP-Stokes").
//   usage: pnp_stokes [domain.dat] [--grid nx ny nz] [--lengths lx ly lz] [bcsr.dat bsr.dat ns.dat]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include "domain.hpp"
#include "linear_pnp_ns.hpp"
#include "newton_status.hpp"

int main(int argc, char** argv)
{
  char exe[4096] = {0};
  std::string here = ".";
  if (readlink("/proc/self/exe", exe, sizeof exe - 1) > 0) { here = exe; here = here.substr(0, here.find_last_of('/')); }
  std::string files[4] = {here + "/domain_stokes.dat", here + "/bcsr.dat", here + "/bsr.dat", here + "/ns.dat"};
  int grid[3] = {0, 0, 0}; double len[3] = {0, 0, 0};
  int pos = 0;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "--grid") && a + 3 < argc) { for (int d = 0; d < 3; ++d) grid[d] = atoi(argv[++a]); }
    else if (!strcmp(argv[a], "--lengths") && a + 3 < argc) { for (int d = 0; d < 3; ++d) len[d] = atof(argv[++a]); }
    else if (pos < 4) files[pos++] = argv[a];
  }
  printf("\n----------------------------------------------------\n Setting up the PNP-Stokes problem\n----------------------------------------------------\n\n");
  domain_param domain;
  if (domain_param_input(files[0].c_str(), &domain) != FASP_SUCCESS) return 1;
  if (grid[0] > 0) { domain.grid_x = grid[0]; domain.grid_y = grid[1]; domain.grid_z = grid[2]; }
  if (len[0] > 0) { domain.length_x = len[0]; domain.length_y = len[1]; domain.length_z = len[2]; }
  pnpb200_handle* mesh = domain_build(domain);
  if (!mesh) { printf("### ERROR: could not build the mesh (no CUDA device?)\n"); return 2; }
  const int nv = pnpb200_num_vertices(mesh), nc = pnpb200_num_cells(mesh);
  std::vector<double> xyz(3 * (size_t)nv); std::vector<int> cells(4 * (size_t)nc);
  pnpb200_get_mesh(mesh, xyz.data(), cells.data());
  pnpb200_destroy(mesh);

  // main.cpp:55-86: outer iteration (bcsr.dat), PNP block (bsr.dat), Stokes block (ns.dat)
  input_param inpar; itsolver_param itpar; AMG_param amgpar; ILU_param ilupar;
  fasp_param_input(files[1].c_str(), &inpar);
  fasp_param_init(&inpar, &itpar, &amgpar, &ilupar, NULL);
  input_param pnp_inpar; itsolver_param pnp_itpar; AMG_param pnp_amgpar; ILU_param pnp_ilupar; Schwarz_param pnp_schpar;
  fasp_param_input(files[2].c_str(), &pnp_inpar);
  fasp_param_init(&pnp_inpar, &pnp_itpar, &pnp_amgpar, &pnp_ilupar, &pnp_schpar);
  input_ns_param ns_inpar; itsolver_ns_param ns_itpar; AMG_ns_param ns_amgpar; ILU_param ns_ilupar; Schwarz_param ns_schpar;
  fasp_ns_param_input(files[3].c_str(), &ns_inpar);
  fasp_ns_param_init(&ns_inpar, &ns_itpar, &ns_amgpar, &ns_ilupar, &ns_schpar);

  printf("Construct vector PNP-Stokes problem\n");
  pnpb200_ns_system* sys = nullptr;
  if (pnpb200_ns_create(&sys, nv, xyz.data(), nc, cells.data()) != FASP_SUCCESS) return 2;
  int nf = 0, nif = 0, ndof = 0, nnz = 0;
  pnpb200_ns_sizes(sys, &nf, &nif, &ndof, &nnz);
  printf("\tvertices %d, cells %d, facets %d (%d interior), dofs %d, nonzeros %d\n", nv, nc, nf, nif, ndof, nnz);
  const double par[8] = {1.0, 1.0, 1.0, 1.0, -1.0, 0.1, 1.0, 1.0};            // main.cpp:123-132
  pnpb200_ns_set_parameters(sys, par);
  Linear_PNP_NS pnp_ns_problem(itpar, pnp_itpar, pnp_amgpar, ns_itpar, ns_amgpar);
  pnp_ns_problem.get_dofs_fasp(sys, nv);

  printf("Initialize Dirichlet BCs & Initial Guess\n");
  std::vector<int> facet_cells(2 * (size_t)nf);
  pnpb200_ns_topology(sys, nullptr, nullptr, facet_cells.data());
  std::vector<int> bc;
  const double Lx = domain.length_x;
  for (int v = 0; v < nv; ++v)
    if (std::fabs(std::fabs(xyz[3 * (size_t)v]) - Lx / 2) < 1e-5 * Lx) for (int k = 0; k < 3; ++k) bc.push_back(3 * v + k);
  for (int f = 0; f < nf; ++f) if (facet_cells[2 * (size_t)f + 1] < 0) bc.push_back(3 * nv + f);
  bc.push_back(3 * nv + nf);
  pnpb200_ns_set_dirichlet(sys, (int)bc.size(), bc.data());
  // main.cpp:146-153: Linear_Function PNP(0, -5, 5, {0, -2.30258509299, 1}, {-2.30258509299, 0, -1}); pressure 0
  const double lo[3] = {0.0, -2.30258509299, 1.0}, hi[3] = {-2.30258509299, 0.0, -1.0};
  std::vector<double> cc(3 * (size_t)nv), uu(nf, 0.0), pp(nc, 0.0);
  for (int v = 0; v < nv; ++v) {
    const double t = (xyz[3 * (size_t)v] + Lx / 2) / Lx;
    for (int k = 0; k < 3; ++k) cc[3 * (size_t)v + k] = lo[k] * (1 - t) + hi[k] * t;
  }
  pnpb200_ns_set_solution(sys, cc.data(), uu.data(), pp.data());

  printf("Initializing nonlinear solver\n");
  const std::size_t max_newton = 20;                                          // main.cpp:169-178
  const double max_residual_tol = 1.0e-10, relative_residual_tol = 1.0e-4;
  const double initial_residual = pnp_ns_problem.compute_residual(sys, "l2");
  Newton_Status newton(max_newton, initial_residual, relative_residual_tol, max_residual_tol);
  printf("\tinitial residual : %10.5e\n\n", newton.initial_residual);
  int last_status = 0;
  while (newton.needs_to_iterate()) {
    printf("Solving for Newton iterate %lu \n", (unsigned long)newton.iteration);
    last_status = pnp_ns_problem.fasp_solve(sys);
    printf("Newton measurements for iteration :\n");
    const double residual = pnp_ns_problem.compute_residual(sys, "l2");
    const double max_residual = pnp_ns_problem.compute_residual(sys, "max");
    newton.update_residuals(residual, max_residual);
    newton.update_iteration();
    printf("\touter iterations :  %d\n\tmaximum residual :  %10.5e\n\trelative residual : %10.5e\n\n", last_status, newton.max_residual,
           newton.relative_residual);
  }
  if (newton.converged()) printf("Solver succeeded!\n"); else newton.print_status();
  pnpb200_ns_get_solution(sys, cc.data(), uu.data(), pp.data());
  double umax = 0.0, pmin = 1e300, pmax = -1e300;
  for (double u : uu) umax = std::fmax(umax, std::fabs(u));
  for (double p : pp) { pmin = std::fmin(pmin, p); pmax = std::fmax(pmax, p); }
  printf("NEWTON_ITERATIONS %lu CONVERGED %d RELRES %.6e MAXFLUX %.6e PRESSURE %.6e %.6e\n", (unsigned long)(newton.iteration - 1),
         newton.converged() ? 1 : 0, newton.relative_residual, umax, pmin, pmax);
  pnpb200_ns_destroy(sys);
  printf("Solver exiting\n");
  return newton.converged() ? 0 : 3;
}

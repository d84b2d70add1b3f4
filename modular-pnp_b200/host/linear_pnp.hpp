// pnp-b200 host side — `Linear_PNP` with the method surface of the reference's problem class
// (benchmarks/pnp_exact_solution/linear_pnp.h:20-101 on top of PDE, include/pde.h:16-221), but on
// flat arrays instead of DOLFIN objects and with every method body a call into the C ABI
// (include/pnpb200/pnpb200.h). Same names, argument meaning and failure behaviour:
//   use_eafe / no_eafe                 linear_pnp.cpp:187-192
//   setup_fasp_linear_algebra          linear_pnp.cpp:73-94
//   fasp_solve                         linear_pnp.cpp:96-132 (update skipped when the solver fails)
//   fasp_test_solver                   linear_pnp.cpp:134-175
//   compute_residual("l2"|"max")       src/pde.cpp:377-397
//   set_coefficients / set_DirichletBC / set_solution / get_solution   src/pde.cpp:250-483
#pragma once
#include <cmath>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "pnpb200/pnpb200.h"

class Linear_PNP {
public:
  // takes ownership of a handle built by domain_build / pnpb200_create
  Linear_PNP(pnpb200_handle* mesh_handle, const itsolver_param& itsolver, const AMG_param& amg)
    : _h(mesh_handle), _itsolver(itsolver), _amg(amg)
  {
    if (!_h) throw std::runtime_error("Linear_PNP: null mesh handle");
    _nv = pnpb200_num_vertices(_h);
    _xyz.resize(3 * (size_t)_nv);
    pnpb200_get_mesh(_h, _xyz.data(), nullptr);
    update_mesh();
  }
  ~Linear_PNP() { pnpb200_destroy(_h); }
  Linear_PNP(const Linear_PNP&) = delete;

  std::size_t get_solution_dimension() const { return 3; }
  std::size_t num_dofs() const { return 3 * (size_t)_nv; }
  const std::vector<double>& coordinates() const { return _xyz; }

  void use_eafe() { _use_eafe = true; pnpb200_use_eafe(_h, 1); }
  void no_eafe() { _use_eafe = false; pnpb200_use_eafe(_h, 0); }

  // nodal coefficient fields (what main.cpp:226-260 interpolates from Expressions)
  void set_coefficients(const std::vector<double>& permittivity, const std::vector<double>& fixed_charge,
                        const std::vector<double>& diffusivity, const std::vector<double>& valency,
                        const std::vector<double>& reaction)
  {
    check(pnpb200_set_coefficients(_h, permittivity.data(), fixed_charge.data(), diffusivity.data(), valency.data(),
                                   reaction.data()), "set_coefficients");
  }

  // PDE::set_DirichletBC (src/pde.cpp:463-483): component i gets homogeneous Dirichlet rows on the
  // two faces of coordinate `component[i]`, and the solution becomes the linear interpolant of the
  // boundary values along that coordinate (Linear_Function::eval, src/dirichlet.cpp:73-90).
  void set_DirichletBC(const std::vector<std::size_t>& component, const std::vector<std::vector<double>>& boundary_value)
  {
    std::vector<int> bc;
    std::vector<double> uu(num_dofs(), 0.0);
    for (std::size_t k = 0; k < component.size() && k < 3; ++k) {
      const std::size_t d = component[k];
      const double mn = _mesh_min[d], mx = _mesh_max[d], dist = 1.0 / (mx - mn);
      for (int v = 0; v < _nv; ++v) {
        const double x = _xyz[3 * (size_t)v + d];
        if (x < mn + _mesh_epsilon || x > mx - _mesh_epsilon) bc.push_back(3 * v + (int)k);   // src/dirichlet.cpp:24-37
        double val = boundary_value[k][0] * (mx - x) * dist;
        val += boundary_value[k][1] * (x - mn) * dist;
        uu[3 * (size_t)v + k] = val;
      }
    }
    check(pnpb200_set_dirichlet(_h, (int)bc.size(), bc.data()), "set_dirichlet");
    set_solution(uu);
  }

  void set_solution(const std::vector<double>& uu) { check(pnpb200_set_solution(_h, uu.data()), "set_solution"); }
  std::vector<double> get_solution()
  {
    std::vector<double> uu(num_dofs());
    check(pnpb200_get_solution(_h, uu.data()), "get_solution");
    return uu;
  }

  double compute_residual(const std::string& norm_type)
  {
    double l2 = 0, mx = 0;
    check(pnpb200_residual(_h, &l2, &mx), "compute_residual");
    return (norm_type == "max" || norm_type == "infinity") ? mx : l2;
  }

  void setup_fasp_linear_algebra() { check(pnpb200_assemble(_h), "setup_fasp_linear_algebra"); }

  // returns the new solution; on solver failure prints the reference's warning and keeps the old one
  std::vector<double> fasp_solve()
  {
    int status = 0; double l2 = 0, mx = 0;
    printf("Solving linear system using FASP solver...\n"); fflush(stdout);
    check(pnpb200_newton_step(_h, &_itsolver, &_amg, &status, &l2, &mx), "fasp_solve");
    _last_l2 = l2; _last_max = mx; _last_status = status;
    if (status < 0) printf("\n### WARNING: FASP solver failed! Exit status = %d.\n", status);
    else printf("Successfully solved the linear system\n");
    fflush(stdout);
    return get_solution();
  }
  // residual norms at the iterate fasp_solve just produced (fused on the device)
  double last_residual(const std::string& norm_type) const
  { return (norm_type == "max" || norm_type == "infinity") ? _last_max : _last_l2; }
  int last_status() const { return _last_status; }

  std::vector<double> fasp_test_solver(const std::vector<double>& target_vector)
  {
    std::vector<double> x(num_dofs());
    int status = 0;
    check(pnpb200_test_solver(_h, target_vector.data(), x.data(), &_itsolver, &_amg, &status), "fasp_test_solver");
    if (status < 0) printf("\n### WARNING: FASP solver failed! Exit status = %d.\n", status);
    return x;
  }

  // Error::compute_l2_error / compute_h1_error (src/error.cpp:58-103) against a nodal exact solution
  void compute_error(const std::vector<double>& exact, double& l2_error, double& h1_error)
  {
    check(pnpb200_error_norms(_h, exact.data(), &l2_error, &h1_error), "compute_error");
  }
  // Linear_PNP::get_total_charge (linear_pnp.cpp:322-351)
  std::vector<double> get_total_charge()
  {
    std::vector<double> q((size_t)_nv);
    check(pnpb200_total_charge(_h, q.data()), "get_total_charge");
    return q;
  }

  // Mesh_Refiner::mark_for_refinement (src/mesh_refiner.cpp:187-209) and
  // mark_for_refinement_with_target_size (:155-185, target_size > 0) with the indicator of
  // compute_entropy_error_vector (:212-271); returns the number of marked cells
  long mark_for_refinement(double entropy_tolerance, std::vector<unsigned char>& cell_marker, long target_size = 0)
  {
    cell_marker.assign((size_t)pnpb200_num_cells(_h), 0);
    long marked = 0;
    check(pnpb200_entropy_indicator(_h, nullptr), "compute_entropy_error_vector");
    check(pnpb200_mark_cells(_h, entropy_tolerance, target_size, cell_marker.data(), &marked, nullptr), "mark_for_refinement");
    return marked;
  }

  pnpb200_handle* handle() { return _h; }

private:
  void check(int rc, const char* what) const
  {
    if (rc != FASP_SUCCESS) throw std::runtime_error(std::string(what) + " failed with FASP status " + std::to_string(rc));
  }
  // PDE::update_mesh (src/pde.cpp:67-87): bounding box and epsilon = 0.1 * rmin. On the uniform
  // Kuhn boxes every cell has the same inradius; it is evaluated from the first cube.
  void update_mesh()
  {
    for (int d = 0; d < 3; ++d) { _mesh_min[d] = 1e20; _mesh_max[d] = -1e20; }
    for (int v = 0; v < _nv; ++v)
      for (int d = 0; d < 3; ++d) {
        const double c = _xyz[3 * (size_t)v + d];
        if (c > _mesh_max[d]) _mesh_max[d] = c;
        if (c < _mesh_min[d]) _mesh_min[d] = c;
      }
    double hmin = 1e20;
    for (int d = 0; d < 3; ++d) {
      double best = 1e20;
      for (int v = 1; v < _nv && v < 4096; ++v) {
        const double dx = std::fabs(_xyz[3 * (size_t)v + d] - _xyz[d]);
        if (dx > 0 && dx < best) best = dx;
      }
      if (best < hmin) hmin = best;
    }
    _mesh_epsilon = 0.1 * (hmin / (3.0 + std::sqrt(3.0)));   // lower bound of the Kuhn-tet inradius
  }

  pnpb200_handle* _h;
  itsolver_param _itsolver;
  AMG_param _amg;
  int _nv = 0;
  bool _use_eafe = true;
  std::vector<double> _xyz;
  double _mesh_min[3], _mesh_max[3], _mesh_epsilon = 0.0;
  double _last_l2 = 0, _last_max = 0;
  int _last_status = 0;
};

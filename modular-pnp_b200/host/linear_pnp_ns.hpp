// pnp-b200 host side — the linear-algebra half of the reference's `Linear_PNP_NS`
// (benchmarks/pnp_stokes/linear_pnp_ns.h, linear_pnp_ns.cpp:23-235) on flat arrays: the caller
// (DOLFIN side) owns the assembled mixed-space matrix / rhs and the per-component dof lists, this
// class owns the FASP-side state and calls the C ABI (include/pnpb200/fasp_compat.h). Same names,
// argument meaning and failure behaviour:
//   get_dofs_fasp                      linear_pnp_ns.cpp:69-107 (PNP dofs interleaved d*i+k; velocity, then pressure)
//   EigenVector_to_dvector_block       linear_pnp_ns.cpp:111-130
//   setup_fasp_linear_algebra          linear_pnp_ns.cpp:132-176 (four fasp_dcsr_getblk cuts, zero initial guess)
//   fasp_solve                         linear_pnp_ns.cpp:178-235 (status < 0: warning, the update is still scattered)
// Assembly: either the caller hands in the assembled mixed-space matrix (a DOLFIN build), or the system is
// assembled on the device from the mesh and the current iterate (pnpb200/pnpb200_ns.h: the P1^3 x RT1 x DG0 forms of
// vector_linear_pnp_ns_forms.ufl incl. the interior-facet terms) — fasp_solve(pnpb200_ns_system*), compute_residual.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>
#include "pnpb200/pnpb200.h"
#include "pnpb200/pnpb200_ns.h"
#include <string>

class Linear_PNP_NS {
public:
  Linear_PNP_NS(const itsolver_param& itsolver, const itsolver_param& pnpitsolver, const AMG_param& pnpamg,
                const itsolver_ns_param& nsitsolver, const AMG_ns_param& nsamg)
    : _itsolver(itsolver), _pnpitsolver(pnpitsolver), _pnpamg(pnpamg), _nsitsolver(nsitsolver), _nsamg(nsamg)
  {
    _fasp_block_matrix.brow = 2; _fasp_block_matrix.bcol = 2;
    _fasp_block_matrix.blocks = (dCSRmat**)calloc(4, sizeof(dCSRmat*));
    fasp_mem_check((void*)_fasp_block_matrix.blocks, "block matrix:cannot allocate memory!\n", ERROR_ALLOC_MEM);
    for (int i = 0; i < 4; ++i) _fasp_block_matrix.blocks[i] = (dCSRmat*)fasp_mem_calloc(1, sizeof(dCSRmat));
    _pnp_dofs.row = _stokes_dofs.row = _velocity_dofs.row = _pressure_dofs.row = 0;
    _pnp_dofs.val = _stokes_dofs.val = _velocity_dofs.val = _pressure_dofs.val = nullptr;
    _fasp_vector.row = _fasp_soln.row = 0; _fasp_vector.val = _fasp_soln.val = nullptr;
  }
  ~Linear_PNP_NS()
  {
    free_blocks();
    for (int i = 0; i < 4; ++i) free(_fasp_block_matrix.blocks[i]);
    free(_fasp_block_matrix.blocks);
    fasp_ivec_free(&_pnp_dofs); fasp_ivec_free(&_stokes_dofs); fasp_ivec_free(&_velocity_dofs); fasp_ivec_free(&_pressure_dofs);
    fasp_dvec_free(&_fasp_vector); fasp_dvec_free(&_fasp_soln);
  }
  Linear_PNP_NS(const Linear_PNP_NS&) = delete;

  // dof_map[c] = sorted mixed-space dofs of solution component c (PDE::get_dofs, src/pde.cpp:404-431);
  // pnp_dimensions = the three PNP components, ns_dimensions = {velocity, pressure}
  void get_dofs_fasp(const std::vector<std::vector<int>>& dof_map, const std::vector<std::size_t>& pnp_dimensions,
                     const std::vector<std::size_t>& ns_dimensions)
  {
    if (ns_dimensions.size() != 2) throw std::runtime_error("get_dofs_fasp: ns_dimensions = {velocity, pressure}");
    int d = 0;
    for (std::size_t c : pnp_dimensions) d += (int)dof_map[c].size();
    fasp_ivec_free(&_pnp_dofs); fasp_ivec_alloc(d, &_pnp_dofs);
    d = 0;
    for (std::size_t c : ns_dimensions) d += (int)dof_map[c].size();
    fasp_ivec_free(&_stokes_dofs); fasp_ivec_alloc(d, &_stokes_dofs);
    const int npd = (int)pnp_dimensions.size();
    for (int k = 0; k < npd; ++k)
      for (std::size_t i = 0; i < dof_map[pnp_dimensions[k]].size(); ++i) _pnp_dofs.val[npd * i + k] = dof_map[pnp_dimensions[k]][i];
    const std::vector<int>& vel = dof_map[ns_dimensions[0]];
    const std::vector<int>& pre = dof_map[ns_dimensions[1]];
    fasp_ivec_free(&_velocity_dofs); fasp_ivec_alloc((int)vel.size(), &_velocity_dofs);
    fasp_ivec_free(&_pressure_dofs); fasp_ivec_alloc((int)pre.size(), &_pressure_dofs);
    int count = 0;
    for (std::size_t i = 0; i < vel.size(); ++i) { _stokes_dofs.val[count++] = vel[i]; _velocity_dofs.val[i] = vel[i]; }
    for (std::size_t i = 0; i < pre.size(); ++i) { _stokes_dofs.val[count++] = pre[i]; _pressure_dofs.val[i] = pre[i]; }
  }

  // matrix / rhs: the assembled mixed-space system (what EigenMatrix_to_dCSRmat aliases, src/pde.cpp:577-630)
  void setup_fasp_linear_algebra(dCSRmat* fasp_matrix, const double* rhs)
  {
    if (_pnp_dofs.row + _stokes_dofs.row != fasp_matrix->row) fasp_chkerr(ERROR_INPUT_PAR, "setup_fasp_linear_algebra");
    free_blocks();
    fasp_dcsr_getblk(fasp_matrix, _pnp_dofs.val, _pnp_dofs.val, _pnp_dofs.row, _pnp_dofs.row, _fasp_block_matrix.blocks[0]);
    fasp_dcsr_getblk(fasp_matrix, _pnp_dofs.val, _stokes_dofs.val, _pnp_dofs.row, _stokes_dofs.row, _fasp_block_matrix.blocks[1]);
    fasp_dcsr_getblk(fasp_matrix, _stokes_dofs.val, _pnp_dofs.val, _stokes_dofs.row, _pnp_dofs.row, _fasp_block_matrix.blocks[2]);
    fasp_dcsr_getblk(fasp_matrix, _stokes_dofs.val, _stokes_dofs.val, _stokes_dofs.row, _stokes_dofs.row, _fasp_block_matrix.blocks[3]);
    const int row = fasp_matrix->row;
    if (_fasp_soln.row != row) { fasp_dvec_free(&_fasp_soln); fasp_dvec_alloc(row, &_fasp_soln); }
    if (_fasp_vector.row != row) { fasp_dvec_free(&_fasp_vector); fasp_dvec_alloc(row, &_fasp_vector); }
    for (int i = 0; i < _pnp_dofs.row; ++i) _fasp_vector.val[i] = rhs[_pnp_dofs.val[i]];                    // EigenVector_to_dvector_block
    for (int i = 0; i < _stokes_dofs.row; ++i) _fasp_vector.val[_pnp_dofs.row + i] = rhs[_stokes_dofs.val[i]];
    fasp_dvec_set(_fasp_vector.row, &_fasp_soln, 0.0);
  }

  // solves the block system on the GPU and scatters the update back to the mixed numbering:
  // update[mixed dof]. Returns the FASP status (outer iterations, or < 0 with a warning like the reference).
  INT fasp_solve(dCSRmat* fasp_matrix, const double* rhs, std::vector<double>& update)
  {
    setup_fasp_linear_algebra(fasp_matrix, rhs);
    printf("Solving linear system using FASP solver...\n"); fflush(stdout);
    const INT status = fasp_solver_bdcsr_krylov_pnp_stokes(&_fasp_block_matrix, &_fasp_vector, &_fasp_soln, &_itsolver,
                                                           &_pnpitsolver, &_pnpamg, &_nsitsolver, &_nsamg,
                                                           _velocity_dofs.row, _pressure_dofs.row);
    if (status < 0) printf("\n### WARNING: FASP solver failed! Exit status = %d.\n", status);
    else printf("Successfully solved the linear system\n");
    fflush(stdout);
    update.assign((std::size_t)fasp_matrix->row, 0.0);
    for (int i = 0; i < _pnp_dofs.row; ++i) update[_pnp_dofs.val[i]] = _fasp_soln.val[i];
    for (int i = 0; i < _stokes_dofs.row; ++i) update[_stokes_dofs.val[i]] = _fasp_soln.val[_pnp_dofs.row + i];
    return status;
  }

  // Device-assembled variants (the system object owns mesh, dofmap, Dirichlet dofs and the current iterate):
  // get_dofs_fasp for this library's numbering 3 v + k | 3 nv + facet | 3 nv + n_facets + cell
  void get_dofs_fasp(pnpb200_ns_system* sys, int n_vertices)
  {
    int nf = 0, nif = 0, nd = 0, nnz = 0;
    pnpb200_ns_sizes(sys, &nf, &nif, &nd, &nnz);
    const int nc = nd - 3 * n_vertices - nf;
    std::vector<std::vector<int>> dof_map(5);
    for (int v = 0; v < n_vertices; ++v) for (int k = 0; k < 3; ++k) dof_map[k].push_back(3 * v + k);
    for (int f = 0; f < nf; ++f) dof_map[3].push_back(3 * n_vertices + f);
    for (int c = 0; c < nc; ++c) dof_map[4].push_back(3 * n_vertices + nf + c);
    get_dofs_fasp(dof_map, {0, 1, 2}, {3, 4});
  }
  // Linear_PNP_NS::fasp_solve (linear_pnp_ns.cpp:178-235): setup_linear_algebra on the device, the four getblk cuts,
  // the block solve, and the update added to the iterate (:216-232)
  INT fasp_solve(pnpb200_ns_system* sys)
  {
    int nf = 0, nif = 0, nd = 0, nnz = 0;
    pnpb200_ns_sizes(sys, &nf, &nif, &nd, &nnz);
    if (pnpb200_ns_assemble(sys) != FASP_SUCCESS) return ERROR_UNKNOWN;
    std::vector<INT> IA(nd + 1), JA(nnz); std::vector<REAL> val(nnz), rhs(nd);
    dCSRmat A; A.row = A.col = nd; A.nnz = nnz; A.IA = IA.data(); A.JA = JA.data(); A.val = val.data();
    dvector b; b.row = nd; b.val = rhs.data();
    INT status = pnpb200_ns_get_system(sys, &A, &b);
    std::vector<double> update;
    if (status == FASP_SUCCESS) {
      status = fasp_solve(&A, b.val, update);
      pnpb200_ns_add_update(sys, update.data(), 1.0);       // like the reference: the update is applied whatever the status
    }
    return status;
  }
  // PDE::compute_residual (src/pde.cpp:377-397): "l2" or "max" norm of the residual vector
  double compute_residual(pnpb200_ns_system* sys, const std::string& norm_type)
  {
    double l2 = -1.0, linf = -1.0;
    if (pnpb200_ns_residual(sys, &l2, &linf) != FASP_SUCCESS) return -1.0;
    if (norm_type == "l2") return l2;
    if (norm_type == "max") return linf;
    return -1.0;
  }

  const ivector& pnp_dofs() const { return _pnp_dofs; }
  const ivector& stokes_dofs() const { return _stokes_dofs; }

private:
  void free_blocks() { for (int i = 0; i < 4; ++i) if (_fasp_block_matrix.blocks[i]->IA) fasp_dcsr_free(_fasp_block_matrix.blocks[i]); }
  itsolver_param _itsolver, _pnpitsolver;
  AMG_param _pnpamg;
  itsolver_ns_param _nsitsolver;
  AMG_ns_param _nsamg;
  block_dCSRmat _fasp_block_matrix;
  ivector _pnp_dofs, _stokes_dofs, _velocity_dofs, _pressure_dofs;
  dvector _fasp_vector, _fasp_soln;
};

/* pnp-b200 — assembly C ABI of the PNP-Stokes Newton system.
 *
 * Replaces the assembly half of the reference's `Linear_PNP_NS` for benchmarks/pnp_stokes:
 *   PDE::setup_linear_algebra       src/pde.cpp:542-555  (dolfin::assemble of a and L + DirichletBC::apply) with the
 *                                   forms of benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.ufl:45-74, i.e. the
 *                                   generated kernels vector_linear_pnp_ns_forms.h:16355-23898 (cell integrals and
 *                                   interior-facet integrals of the bilinear and the linear form)
 *   PDE::compute_residual           src/pde.cpp:377-397
 *   the solution update             benchmarks/pnp_stokes/linear_pnp_ns.cpp:216-232
 * What it hands back — the mixed-space dCSRmat and right-hand side — is what
 * Linear_PNP_NS::setup_fasp_linear_algebra (linear_pnp_ns.cpp:137-176) cuts into the 2 x 2 block system with
 * fasp_dcsr_getblk and gives to fasp_solver_bdcsr_krylov_pnp_stokes (fasp_compat.h).
 *
 * Mixed space P1^3 x RT1 x DG0 on a conforming tetrahedral mesh. Dof numbering of this library:
 *     3 * vertex + k         k = 0 cation, 1 anion, 2 potential (log densities; the `cc` coefficient)
 *     3 * n_vertices + facet             facet-normal velocity dofs (lowest-order Raviart-Thomas; `uu`)
 *     3 * n_vertices + n_facets + cell   pressure (`pp`)
 * Facets are numbered by this library (pnpb200_ns_topology returns the numbering); cell vertices are sorted
 * ascending like UFC's mesh ordering, facet i of a cell is opposite its i-th vertex, and the velocity basis function
 * of a facet is the FIAT/FFC one: phi_i(x) = s_i (x - v_i) / detJ, s = (-1, +1, -1, +1).
 * Everything runs on the GPU (fp64); host arrays are borrowed during a call. Returns 0 or a negative FASP-style status.
 */
#ifndef PNPB200_NS_H
#define PNPB200_NS_H
#include "fasp_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pnpb200_ns_system pnpb200_ns_system;

/* xyz[3 * n_vertices], cells[4 * n_cells] (any vertex order inside a cell). Builds the facet numbering, the adjacency
 * lists and the sparsity pattern of the mixed space on the host (once per mesh: Mesh::init(2) + the pattern builder of
 * dolfin::assemble) and uploads them to the current CUDA device. */
int  pnpb200_ns_create(pnpb200_ns_system** s, int n_vertices, const double* xyz, int n_cells, const int* cells);
void pnpb200_ns_destroy(pnpb200_ns_system* s);
int  pnpb200_ns_sizes(const pnpb200_ns_system* s, int* n_facets, int* n_interior_facets, int* n_dofs, int* nnz);
/* cells_sorted[4 nc], cell_facets[4 nc] (facet opposite the i-th sorted vertex), facet_cells[2 nf] (lower cell first,
 * -1 on the boundary); any pointer may be NULL */
int  pnpb200_ns_topology(const pnpb200_ns_system* s, int* cells_sorted, int* cell_facets, int* facet_cells);

/* par8 = {permittivity, diffusivity0, valency0, diffusivity1, valency1, mu, penalty1, penalty2}
 * (benchmarks/pnp_stokes/main.cpp:123-132; default = those values) */
int  pnpb200_ns_set_parameters(pnpb200_ns_system* s, const double par8[8]);
/* homogeneous Dirichlet dofs of the Newton update, in the mixed numbering (Linear_PNP_NS::init_BC,
 * linear_pnp_ns.cpp:287-319): their rows become identity rows, their right-hand side 0 */
int  pnpb200_ns_set_dirichlet(pnpb200_ns_system* s, int n_bc, const int* bc_dofs);
/* the previous iterates: cc[3 nv] (dof 3 v + k), uu[n_facets], pp[n_cells] */
int  pnpb200_ns_set_solution(pnpb200_ns_system* s, const double* cc, const double* uu, const double* pp);
int  pnpb200_ns_get_solution(pnpb200_ns_system* s, double* cc, double* uu, double* pp);

/* matrix and right-hand side of the Newton system at the current iterate (device resident) */
int  pnpb200_ns_assemble(pnpb200_ns_system* s);
/* copies them out: A with row = col = n_dofs, nnz as pnpb200_ns_sizes reports, IA / JA / val allocated by the caller;
 * b with n_dofs entries. Columns ascending inside a row. Either may be NULL. */
int  pnpb200_ns_get_system(pnpb200_ns_system* s, dCSRmat* A, dvector* b);
/* l2 and max norm of the residual vector (L assembled at the current iterate, Dirichlet rows zero) */
int  pnpb200_ns_residual(pnpb200_ns_system* s, double* l2, double* linf);
/* iterate += alpha * du, du[n_dofs] in the mixed numbering */
int  pnpb200_ns_add_update(pnpb200_ns_system* s, const double* du, double alpha);

#ifdef __cplusplus
}
#endif
#endif

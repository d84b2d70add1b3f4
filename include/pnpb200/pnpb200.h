/* pnp-b200 — assembly / Newton-step C ABI (drop-in boundary B.2 of SURVEY.md §8b).
 *
 * Replaces the BODIES of the reference's hot-path methods; every entry point cites the
 * reference code it stands in for (paths relative to the reference root):
 *   PDE::setup_linear_algebra            src/pde.cpp:542-555
 *   PDE::compute_residual                src/pde.cpp:377-397
 *   Linear_PNP::setup_fasp_linear_algebra benchmarks/pnp_exact_solution/linear_pnp.cpp:73-94
 *   Linear_PNP::apply_eafe               benchmarks/pnp_exact_solution/linear_pnp.cpp:195-285
 *   Linear_PNP::fasp_solve               benchmarks/pnp_exact_solution/linear_pnp.cpp:96-132
 * The data those bodies have in hand are flat DOLFIN arrays (mesh.coordinates(), mesh.cells(),
 * Function::vector()->get_local(), DirichletBC::get_boundary_values), which is what these
 * functions take. Everything runs on the GPU (sm_100a, fp64); host arrays are borrowed for the
 * duration of a call, the handle owns all device memory. Return value: 0 (FASP_SUCCESS) or a
 * negative FASP-style status (fasp_compat.h); never abort().
 *
 * Dof numbering: blocked, dof = 3*vertex + component (component 0 = potential, 1..2 = log
 * densities), the numbering fasp_format_dcsr_dbsr(...,3) needs for dense 3x3 blocks.
 */
#ifndef PNPB200_H
#define PNPB200_H
#include "fasp_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pnpb200_handle pnpb200_handle;

/* library / device */
const char* pnpb200_version(void);
int  pnpb200_device_count(void);

/* create from a P1 tetrahedral mesh: xyz[3*n_vertices], cells[4*n_cells] (vertex ids).
 * Builds on the device: vertex->cell adjacency, the BSR(3) sparsity pattern (all vertex pairs
 * sharing a cell — what dolfin::assemble's pattern builder produces, src/pde.cpp:545), the EAFE
 * edge table omega_ij, and the fine-level colouring. device = CUDA ordinal. */
int  pnpb200_create(pnpb200_handle** h, int n_vertices, const double* xyz, int n_cells,
                    const int* cells, int device);
/* domain_build (src/domain.cpp:290-311): the centred dolfin::BoxMesh(-L/2, +L/2, nx, ny, nz) every
 * BASELINE config runs on, generated directly in device memory (no host mesh), then as create. */
int  pnpb200_create_box(pnpb200_handle** h, int nx, int ny, int nz, double lx, double ly, double lz,
                        int device);
/* copy the mesh back (xyz[3*Nv], cells[4*Nc]; either may be NULL) */
int  pnpb200_get_mesh(pnpb200_handle* h, double* xyz, int* cells);
int  pnpb200_num_vertices(const pnpb200_handle* h);
int  pnpb200_num_cells(const pnpb200_handle* h);
void pnpb200_destroy(pnpb200_handle* h);
/* the CUDA stream (cudaStream_t as void*) all work of this handle is enqueued on; callers that
 * time with their own events must record them on this stream */
void* pnpb200_get_stream(pnpb200_handle* h);

/* PDE::set_coefficients (src/pde.cpp:355-375): nodal coefficient fields.
 * permittivity[Nv], fixed_charge[Nv]; diffusivity/valency/reaction [3*Nv] in dof order. */
int  pnpb200_set_coefficients(pnpb200_handle* h, const double* permittivity,
                              const double* fixed_charge, const double* diffusivity,
                              const double* valency, const double* reaction);
/* The caller's dof numbering (PDE::get_dofs src/pde.cpp:404-431; the component-major cell dofs of
 * vector_linear_pnp_forms.h:7706-7726 after DOLFIN's renumbering). Internally dof = 3*vertex + component.
 * vertex_to_dof[3*v + k] = the caller's index of (vertex v, component k) — exactly dolfin::vertex_to_dof_map(V)
 * of the mixed space; vertex_to_scalar_dof[v] the same for the P1 coefficient spaces (permittivity, fixed
 * charge; NULL = vertex order). From then on EVERY dof-indexed host array of this header (solution,
 * coefficients, Dirichlet dofs, rhs, update, exact solution, solver-test vectors, total charge) is in the
 * caller's numbering; matrices (pnpb200_get_matrix) stay in vertex-block order. NULL, NULL resets. Call
 * before pnpb200_set_coefficients / pnpb200_set_dirichlet. */
int  pnpb200_set_dof_map(pnpb200_handle* h, const int* vertex_to_dof, const int* vertex_to_scalar_dof);
/* PDE::set_DirichletBC / remove_Dirichlet_dof (src/pde.cpp:433-483): homogeneous Dirichlet
 * rows for the Newton update on the listed scalar dofs. */
int  pnpb200_set_dirichlet(pnpb200_handle* h, int n_bc, const int* bc_dofs);
/* PDE::set_solution / get_solution (src/pde.cpp:250-275) */
int  pnpb200_set_solution(pnpb200_handle* h, const double* uu);
int  pnpb200_get_solution(pnpb200_handle* h, double* uu);
/* same, with buffers already resident on this handle's device */
int  pnpb200_set_solution_device(pnpb200_handle* h, const double* d_uu);
int  pnpb200_get_solution_device(pnpb200_handle* h, double* d_uu);
/* Linear_PNP::use_eafe / no_eafe (linear_pnp.cpp:187-192) */
int  pnpb200_use_eafe(pnpb200_handle* h, int on);

/* Form variant of benchmarks/pnp_diode/vector_linear_pnp_forms.{ufl,h}: `poisson_scale` (a UFL
 * Constant there) multiplies the charge terms of the Poisson row in a and L
 * (forms.h:8721-8724, L I[9]); with conc_cutoff != 0 the Jacobian's diffusion terms use
 * w > -50 ? exp(w) : 0 (forms.h:8606-8608 — the generated header, which is what runs, not the
 * rational tail of the .ufl). Default (1.0, 0) = the pnp_exact_solution forms. */
int  pnpb200_set_form_variant(pnpb200_handle* h, double poisson_scale, int conc_cutoff);

/* sizes */
int  pnpb200_num_dofs(const pnpb200_handle* h);
int  pnpb200_num_block_rows(const pnpb200_handle* h);
int  pnpb200_num_blocks(const pnpb200_handle* h);

/* Linear_PNP::setup_fasp_linear_algebra: assemble Jacobian (Galerkin + EAFE overwrite +
 * Dirichlet rows + zero-row fix, straight into BSR(3)) and the rhs on the device. */
int  pnpb200_assemble(pnpb200_handle* h);
/* copy the assembled system to host FASP structs. A_out arrays are allocated with the FASP
 * allocator (free with fasp_dbsr_free); b_out must have row == num_dofs. Either may be NULL. */
int  pnpb200_get_matrix(pnpb200_handle* h, dBSRmat* A_out);
int  pnpb200_get_rhs(pnpb200_handle* h, dvector* b_out);
/* PDE::compute_residual: assemble L at the current solution, zero Dirichlet rows, return
 * the l2 and max norms (fused; either pointer may be NULL). */
int  pnpb200_residual(pnpb200_handle* h, double* l2, double* linf);

/* Linear_PNP::fasp_solve without the assembly: solve J du = b for the system currently on
 * the device (AMG setup + FGMRES). status_its gets iterations >= 0 or a negative FASP status. */
int  pnpb200_solve(pnpb200_handle* h, const itsolver_param* itparam, const AMG_param* amgparam,
                   int* status_its);
/* copy the last Newton update du to the host */
int  pnpb200_get_update(pnpb200_handle* h, double* du);
/* Linear_PNP::fasp_test_solver (linear_pnp.cpp:134-175): rhs = J*target, solve, return x */
int  pnpb200_test_solver(pnpb200_handle* h, const double* target, double* x,
                         const itsolver_param* itparam, const AMG_param* amgparam, int* status_its);

/* One full Newton step of benchmarks/pnp_exact_solution/main.cpp:334-345, nothing leaving
 * the device: assemble, AMG setup, FGMRES, u += du (skipped if the solver failed,
 * linear_pnp.cpp:109-127), residual norms at the new iterate. */
int  pnpb200_newton_step(pnpb200_handle* h, const itsolver_param* itparam, const AMG_param* amgparam,
                         int* krylov_status_its, double* l2, double* linf);

/* Damped Newton of the diode benchmark (benchmarks/pnp_diode/pnp_newton_solver.h:195-284)
 * without host round trips of the solution: every iterate that loop evaluates is
 * base + alpha * update (full step, growth-limited step, the <= 50 backtracking steps).
 *   pnpb200_solve_update   Linear_PNP::fasp_solve without the update: assembles J and the residual at the
 *                          current solution, solves J du = b (du stays on the device), remembers the solution
 *                          as `base`. *krylov_status as in pnpb200_solve.
 *   pnpb200_update_ranges  ranges6 = {min, max} of base, of du and of base + du (:220-237, all ranks)
 *   pnpb200_trial_residual solution <- base + alpha * du, then PDE::compute_residual there (:217, :258-268):
 *                          one residual-only pass, the l2 and max norms in one sweep */
int  pnpb200_solve_update(pnpb200_handle* h, const itsolver_param* itparam, const AMG_param* amgparam, int* krylov_status_its);
int  pnpb200_update_ranges(pnpb200_handle* h, double* ranges6);
int  pnpb200_trial_residual(pnpb200_handle* h, double alpha, double* l2, double* linf);

/* Post-processing on the device. Error::compute_l2_error / compute_h1_error
 * (src/error.cpp:58-103) against the NODAL interpolant `exact[3*Nv]` of the exact solution, summed
 * over the three components: l2 = sqrt(sum_k int e_k^2), h1 = sqrt(l2^2 + sum_k int |grad e_k|^2).
 * (single-GPU handles; on a partitioned handle ghost cells would be counted twice). */
int  pnpb200_error_norms(pnpb200_handle* h, const double* exact, double* l2_error, double* h1_error);
/* Linear_PNP::get_total_charge (linear_pnp.cpp:322-351): nodal fixed_charge + sum_k q_k exp(u_k), charge[Nv] */
int  pnpb200_total_charge(pnpb200_handle* h, double* charge);

/* Adaptive-refinement indicator on the device.
 * pnpb200_entropy_indicator = Mesh_Refiner::compute_entropy_error_vector (src/mesh_refiner.cpp:212-271;
 * form ufl/poisson_cell_marker.ufl:6-15, kernels include/poisson_cell_marker.h:5203-5231,5255-5422) with
 * the bindings of benchmarks/pnp_refine_exact/main.cpp:260-266: per ion species k the entropy potential
 * u_k + q_k phi, log weight u_k, diffusivity D_k; the mass-lumped DG0 values of the species are summed.
 * eta[n_cells] may be NULL (the values stay on the device for pnpb200_mark_cells).
 * pnpb200_mark_cells = Mesh_Refiner::mark_for_refinement (:187-209, target_size = 0: marker = eta >
 * entropy_tolerance) and mark_for_refinement_with_target_size (:155-185, target_size > 0: tolerance =
 * max(sorted_eta[n - min(n, permissible)], entropy_tolerance), permissible = n > target ? 1 : target - n).
 * marker[n_cells] receives 0/1, *n_marked the count, *tolerance_used (may be NULL) the tolerance applied.
 * Refining the mesh itself stays with the caller (dolfin::adapt): pnpb200_create on the new (xyz, cells)
 * rebuilds pattern, edge table, colouring and hierarchy on the device. */
int  pnpb200_entropy_indicator(pnpb200_handle* h, double* eta);
int  pnpb200_mark_cells(pnpb200_handle* h, double entropy_tolerance, long target_size, unsigned char* marker,
                        long* n_marked, double* tolerance_used);

/* hierarchy policy: 0 = rebuild aggregates/patterns/colourings every solve (the reference's
 * behaviour, default), 1 = keep the symbolic hierarchy of the previous solve and only
 * recompute numerical values (Galerkin products, block inverses). */
int  pnpb200_set_hierarchy_reuse(pnpb200_handle* h, int reuse);
/* sweep shortcuts: 1 (default) = the cycle exploits the [L | D | U] row storage: forward sweep from
 * the zero initial guess reads L only, the residual after it U only, and A*z after the last backward
 * sweep is b + L*(z_new - z_old); 0 = every sweep / residual / matvec reads full rows. Same
 * arithmetic up to rounding (tests/test_gpu_solver.py), 2.5 instead of 4 matrix passes per
 * Krylov iteration on every level. */
int  pnpb200_set_sweep_shortcuts(pnpb200_handle* h, int on);

/* Krylov preconditioner of pnpb200_solve / _newton_step / _solve_update: kind 0 = BSR UA-AMG cycle (bsr.dat of
 * pnp_exact_solution, default), 1 = block ILU(k) with k = ilu_lfil, the diode configuration
 * (benchmarks/pnp_diode/linear_pnp.cpp:103-109, bsr.dat:34-38: ILUk, lfil 2). Symbolic factorisation once per
 * pattern (host), numeric factorisation and the triangular solves on the device (csrc/ilu.cu). */
int  pnpb200_set_preconditioner(pnpb200_handle* h, int kind, int ilu_lfil);
/* the ILU factors of the last solve, FASP-style arrays in vertex order (parity tests): n_blocks first, then the
 * arrays (any may be NULL); val holds L (strict lower, unit diagonal implied) and U, dinv the inverse U_ii */
int  pnpb200_get_ilu_factors(pnpb200_handle* h, int* n_blocks, int* IA, int* JA, double* val, double* dinv);
/* z = U^{-1} L^{-1} r with those factors (host vectors, vertex order) */
int  pnpb200_apply_ilu(pnpb200_handle* h, const double* r, double* z);

/* ---- introspection for the parity tests and the byte model of bench.py ---- */
typedef struct pnpb200_stats {
  int    n_levels;
  int    level_rows[32];
  int    level_nnzb[32];
  int    level_colors[32];
  int    krylov_its;          /* last solve */
  int    launches;            /* kernels launched by the last call */
  double ms_assemble, ms_setup, ms_krylov, ms_residual;   /* CUDA-event times of the last call */
  double ms_spmv, ms_vcycle, ms_ortho;   /* accumulated inside the last Krylov solve (when timing detail is on) */
} pnpb200_stats;
int  pnpb200_get_stats(const pnpb200_handle* h, pnpb200_stats* s);
int  pnpb200_set_timing_detail(pnpb200_handle* h, int on);
/* counters of the calling thread's last fasp_solver_bdcsr_krylov_pnp_stokes call:
 * out[0] outer iterations, out[1] PNP block solves, out[2] their Krylov iterations in total,
 * out[3] Stokes block solves, out[4] their Krylov iterations in total; *relres = final true
 * relative residual of the coupled system */
int  pnpb200_last_block_solve_stats(int out[5], double* relres);
/* inner velocity-block solves of the same call (ns.dat's _v set): {solves, Krylov iterations} */
int  pnpb200_last_block_velocity_stats(int out[2]);
/* copy level-l hierarchy pieces to the host (FASP layout): matrix, aggregate map (length
 * level_rows[l], -1 = isolated), GS row order (rows sorted by colour). */
int  pnpb200_get_level_matrix(pnpb200_handle* h, int level, dBSRmat* A_out);
int  pnpb200_get_level_aggregates(pnpb200_handle* h, int level, int* agg, int* n_agg);
int  pnpb200_get_level_order(pnpb200_handle* h, int level, int* order);
/* apply one V-cycle of the current hierarchy: z = M^{-1} r (host vectors) */
int  pnpb200_apply_vcycle(pnpb200_handle* h, const double* r, double* z);
/* roofline probe: run `reps` back-to-back launches of one hot kernel on the fine level, timed
 * with CUDA events on the handle's stream; returns the mean ms per launch and the ALGORITHMIC
 * bytes one launch moves (SURVEY §8d model). which: 0 = BSR SpMV y=Ax, 1 = residual r=b-Ax,
 * 2 = one multicolour Gauss-Seidel sweep (all colours = one matrix pass),
 * 3 = fused multi-dot over 16 basis vectors, 4 = residual cell kernel, 5 = Jacobian cell kernel,
 * 6 = block assembly kernel. Needs an assembled system (and a solve for 2/3). */
int  pnpb200_probe_kernel(pnpb200_handle* h, int which, int reps, double* ms_per_launch,
                          double* algorithmic_bytes);
/* values per 3x3 block the matrix passes of the last solve read on `level`: 9 (FASP's dense blocks) or 7 (PNP blocks have
 * no cation-anion coupling: entries (1,2) and (2,1) are zero on every level; the solve then reads a 7-value copy) */
int  pnpb200_get_value_format(pnpb200_handle* h, int level);
/* fasp_solver_dbsr_krylov_amg / _ilu (fasp_compat.h) keep the device copy of the LAST host matrix — stream, colouring,
 * [L | D | U] layout, level buffers, ILU pattern — and re-use it when the next call brings the same sparsity pattern (one
 * Newton step after the other, linear_pnp.cpp:101). This frees it (device memory back to the pool). */
void pnpb200_release_host_cache(void);
/* development aid: run matrix pass `which` (probe numbering: 0, 1, 2, 7..10) of the fine level reps + 1 times from the
 * same input and compare the results bit by bit; out[18]: [0] repetitions that differed, [1] differing entries of
 * the first one, [2..16] their sweep-space positions, [17] its index */
int  pnpb200_debug_repeat_pass(pnpb200_handle* h, int which, int reps, int* out);
/* y = J x with the assembled fine-level matrix (host vectors) */
int  pnpb200_apply_matrix(pnpb200_handle* h, const double* x, double* y);

/* ---- multi-GPU: one process per GPU, block rows partitioned by vertex ----
 * Each rank creates its handle from its LOCAL mesh: owned vertices first (local ids
 * 0..n_owned-1), then the ghost vertices its cells touch, grouped by owning neighbour rank;
 * cells = every cell that touches an owned vertex. Ghost dofs are identity rows of the local
 * matrix; their values are refreshed by halo exchange (grouped ncclSend/ncclRecv) before every
 * global SpMV / residual / assembly, and every dot product / norm is an fp64 ncclAllReduce.
 * pnpb200_comm_unique_id: rank 0 creates the 128-byte NCCL id, the caller's own bootstrap hands
 * it to every rank (bench.py uses torch.distributed), then every rank calls pnpb200_comm_init. */
int  pnpb200_comm_unique_id(unsigned char id[128]);
int  pnpb200_comm_init(pnpb200_handle* h, const unsigned char id[128], int rank, int n_ranks);
/* halo exchanges issued so far: out[0] through the peer-memory kernel (one launch: pack, NVLink stores into the
 * neighbours' receive slots, flags, unpack), out[1] through NCCL send / recv (fallback, PNPB200_HALO_IPC=0) */
int  pnpb200_comm_halo_counts(const pnpb200_handle* h, long long out[2]);
/* neighbour k: send the owned vertices send_idx[send_ptr[k] .. send_ptr[k+1]) (in that order) and
 * receive its values into the local ghost vertices [recv_ptr[k], recv_ptr[k+1]). Call before
 * pnpb200_set_dirichlet / after create. */
int  pnpb200_set_partition(pnpb200_handle* h, int n_owned, int n_neighbors, const int* neighbor_ranks,
                           const int* send_ptr, const int* send_idx, const int* recv_ptr);
/* A proper colouring of the local fine-level graph known to the caller (colour[v] in [0, n_colors),
 * no two vertices of a cell share a colour), e.g. (i + j + k) mod 4 of the GLOBAL lattice position
 * of a BoxMesh brick (pnpb200_partition_box): the multicolour Gauss-Seidel sweeps of the fine level
 * then take n_colors launches instead of the ~10 of the graph colouring. Checked against the
 * pattern on the device; ERROR_INPUT_PAR if it is not proper. Call after pnpb200_set_partition
 * (or right after create on one GPU). */
int  pnpb200_set_coloring(pnpb200_handle* h, int n_colors, const int* color);

/* ---- mesh partitioner (host, csrc/partition.cpp): the vertex partition + halo lists that
 * pnpb200_create / pnpb200_set_partition consume, for rank `rank`.
 *   pnpb200_partition_box : dolfin::BoxMesh (src/domain.cpp:306-308) cut into px x py x pz bricks of
 *     vertex planes (rank = rx + px * (ry + py * rz)), generated in closed form — only the rank's own
 *     brick and its ghost layer are ever built. (1, 1, N) gives z-slabs.
 *   pnpb200_partition_mesh: recursive coordinate bisection of ANY conforming tetrahedral mesh given by
 *     its global arrays (the refined meshes of src/mesh_refiner.cpp:73-153); deterministic, so every
 *     rank computes the same owner map.
 * Local vertices: owned (ascending global id), then ghosts grouped by owner rank (ascending rank,
 * ascending global id); local cells: every cell touching an owned vertex, vertices ascending in
 * GLOBAL id (UFC order). The arrays stay owned by the partition object (pnpb200_partition_free). */
typedef struct pnpb200_partition pnpb200_partition;
int  pnpb200_partition_box(int nx, int ny, int nz, double lx, double ly, double lz, int px, int py, int pz, int rank,
                           pnpb200_partition** out);
int  pnpb200_partition_mesh(int n_vertices, const double* xyz, int n_cells, const int* cells, int n_ranks, int rank,
                            pnpb200_partition** out);
int  pnpb200_partition_sizes(const pnpb200_partition* p, int* n_owned, int* n_local, int* n_cells, int* n_neighbors,
                             int* n_send, int* n_bc);
/* gids: global vertex id of every local vertex; color: (i + j + k) mod 4 for box partitions, NULL otherwise;
 * bc_vertices: local ids of the vertices on the two x-faces (box partitions; the benchmarks' Dirichlet set) */
int  pnpb200_partition_arrays(const pnpb200_partition* p, const double** xyz, const int** cells, const long long** gids,
                              const int** neighbor_ranks, const int** send_ptr, const int** send_idx, const int** recv_ptr,
                              const int** color, const int** bc_vertices);
void pnpb200_partition_free(pnpb200_partition* p);

#ifdef __cplusplus
}
#endif
#endif

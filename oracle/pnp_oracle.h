/* TEST INFRASTRUCTURE ONLY — CPU oracle for the pnp-b200 hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libpnpb200.so) never does.
 *
 * What it restates (citations relative to /root/reference):
 *   - element kernels: include/EAFE.h:4712-4910,
 *     benchmarks/pnp_exact_solution/vector_linear_pnp_forms.h:7917-8360 (a), :8385-8644 (L)
 *   - assembly orchestration: src/pde.cpp:377-397,404-483,542-630,
 *     benchmarks/pnp_exact_solution/linear_pnp.cpp:73-132,195-285
 *   - Newton control: src/newton_status.cpp:11-74, benchmarks/pnp_exact_solution/main.cpp:293-356
 *   - DOLFIN BoxMesh / P1^3 dofmap / DirichletBC and FASP BSR-UA-AMG + pvfgmres: these live in
 *     un-vendored third-party code (DOLFIN 2016.1/2016.2, FASP ~1.8-1.9; no version pinned in
 *     CMakeLists.txt:14,40-48). Restated from their published algorithms.
 *
 * PARITY PINNING: the element kernels are pinned against the reference's OWN compiled
 * kernels (oracle/_ref/libpnpref.so, built by oracle/Makefile from /root/reference) and
 * committed fixtures generated from them (tests/golden). The DOLFIN/FASP restatement has no
 * golden vectors in the reference tree => "parity unpinned" for global assembly order,
 * Krylov iteration counts and residual histories (pinned only by the reference's
 * known-answer problems, tests/eafe_tests/test_eafe.cpp:195,272 etc.).
 */
#ifndef PNP_ORACLE_H
#define PNP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---- element kernels (restated) ---- */
void orc_eafe_element(double* A16, const double* eta4, const double* alpha4,
                      const double* beta4, const double* gamma4, const double* xyz12);
void orc_jac_element(double* A144, const double* uu12, const double* eps4,
                     const double* diff12, const double* val12, const double* xyz12);
void orc_res_element(double* b12, const double* uu12, const double* eps4, const double* fc4,
                     const double* diff12, const double* val12, const double* reac12,
                     const double* xyz12);
/* choose which element kernels the assembler drives: 0 = restated (default),
 * 1 = the reference's own (needs oracle/_ref/libpnpref.so; returns -1 if absent) */
int  orc_use_ref_kernels(int on, const char* libpath);
/* form variant of benchmarks/pnp_diode/vector_linear_pnp_forms.{ufl,h}: poisson_scale on the charge
 * terms of the Poisson row, exp cutoff (w > -50 ? exp(w) : 0) in the Jacobian's diffusion terms.
 * (1.0, 0) = pnp_exact_solution forms (default). */
void orc_set_form_variant(double poisson_scale, int conc_cutoff);

/* ---- mesh / dofs / BC ---- */
void orc_box_mesh_counts(int nx, int ny, int nz, int* nv, int* nc);
void orc_box_mesh(int nx, int ny, int nz, double lx, double ly, double lz, double* xyz, int* cells);
/* Dirichlet vertices as src/pde.cpp:433-461 + src/dirichlet.cpp:24-37 select them on axis
 * `coord`; returns count, fills flags[nv] (0/1) */
int  orc_dirichlet_vertices(int nv, const double* xyz, int nc, const int* cells, int coord,
                            unsigned char* flags);

/* ---- block pattern (vertex graph incl. diagonal, sorted) ---- */
int  orc_block_pattern_count(int nv, int nc, const int* cells, int* IA /*nv+1*/);
void orc_block_pattern_fill(int nv, int nc, const int* cells, const int* IA, int* JA);

/* ---- assembly: mirrors Linear_PNP::setup_fasp_linear_algebra (linear_pnp.cpp:73-94) ----
 * outputs BSR(3) with FASP layout (row-major 3x3 blocks) on the block pattern and rhs b.
 * dof order is blocked: dof = 3*vertex + component. bcflag[N] marks Dirichlet dofs.
 * coefficients: eps[nv], fixed[nv], diff[N], val[N], reac[N]. */
void orc_assemble(int nv, const double* xyz, int nc, const int* cells,
                  const int* IA, const int* JA,
                  const double* uu, const double* eps, const double* fixed,
                  const double* diff, const double* val, const double* reac,
                  const unsigned char* bcflag, int use_eafe,
                  double* bsr_val /*9*nnzb*/, double* b /*N*/);
/* residual vector with BC rows zeroed, src/pde.cpp:377-397 */
void orc_residual(int nv, const double* xyz, int nc, const int* cells,
                  const double* uu, const double* eps, const double* fixed,
                  const double* diff, const double* val, const double* reac,
                  const unsigned char* bcflag, double* b, double* l2, double* linf);
/* scalar EAFE matrix on the vertex pattern (include/EAFE.h via dolfin::assemble, linear_pnp.cpp:256) */
void orc_eafe_matrix(int nv, const double* xyz, int nc, const int* cells,
                     const int* IA, const int* JA,
                     const double* eta, const double* alpha, const double* beta, const double* gamma,
                     double* val /*nnz of vertex pattern*/);

/* L2 / H1 error against a nodal exact solution, src/error.cpp:58-103 */
/* Mesh_Refiner::compute_entropy_error_vector (src/mesh_refiner.cpp:212-271), eta[nc] */
double orc_marker_element(const double* mu4, const double* logw4, const double* diff4, const double* xyz12);
void orc_entropy_indicator(int nv, const double* xyz, int nc, const int* cells, const double* uu,
                           const double* diff, const double* val, double* eta);
void orc_error_norms(int nv, const double* xyz, int nc, const int* cells, const double* uu,
                     const double* exact, double* l2, double* h1);

/* ---- PNP-Stokes element kernels (benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.ufl:45-74; oracle/ns_elements.cpp):
 * P1^3 x RT1 x DG0, 17 dofs per cell, macro element of an interior facet = 34. par8 = {permittivity, diffusivity0,
 * valency0, diffusivity1, valency1, mu, penalty1, penalty2}. Pinned against oracle/_ref/libpnpref_ns.so. ---- */
void orc_ns_cell_a(double* A289, const double* cc12, const double* uu4, const double* par8, const double* xyz12);
void orc_ns_cell_L(double* b17, const double* cc12, const double* uu4, const double* pp1, const double* par8, const double* xyz12);
void orc_ns_facet_a(double* A1156, const double* par8, const double* xyz12_0, const double* xyz12_1, int facet_0, int facet_1);
void orc_ns_facet_L(double* b34, const double* uu8, const double* par8, const double* xyz12_0, const double* xyz12_1,
                    int facet_0, int facet_1);

/* dolfin::assemble of the mixed PNP-Stokes system restated (cell loop + interior-facet loop) as COO triplets
 * (289 nc + 1156 n_interior entries, zeros included); ref_lib = path of oracle/_ref/libpnpref_ns.so to drive the
 * reference's own kernels, NULL = the restatement. Returns the number of triplets, -1 if ref_lib cannot be loaded. */
long orc_ns_assemble(int nv, const double* xyz, int nc, const int* cells_sorted, const int* cell_facets, int nf,
                     const int* facet_cells, const double* cc, const double* uu, const double* pp, const double* par8,
                     const char* ref_lib, int* coo_row, int* coo_col, double* coo_val, double* rhs);

/* ---- FASP-restated linear algebra (BSR nb=3) ---- */
void orc_csr_to_bsr3_count(int n, const int* IA, const int* JA, int* bIA);
void orc_csr_to_bsr3_fill(int n, const int* IA, const int* JA, const double* val,
                          const int* bIA, int* bJA, double* bval);
void orc_bsr_spmv(int nrow, const int* IA, const int* JA, const double* val,
                  double alpha, const double* x, double beta, double* y);
void orc_bsr_diaginv(int nrow, const int* IA, const int* JA, const double* val, double* dinv);
/* one block-GS sweep in the given row order (order==NULL: ascending; FASP
 * fasp_smoother_dbsr_gs_ascend / _order1 [EXT]) */
void orc_bsr_gs(int nrow, const int* IA, const int* JA, const double* val, const double* dinv,
                const double* b, double* x, const int* order, int norder);
/* FASP VMB aggregation on the L-inf condensed matrix [EXT]; returns #aggregates, agg[i] in
 * [-1 (isolated), nagg) */
int  orc_vmb_aggregate(int nrow, const int* IA, const int* JA, const double* val,
                       double strong_coupled, int max_aggregation, int* agg);
/* Galerkin RAP with boolean block prolongator given by agg[] */
int  orc_rap_agg_count(int nrow, const int* IA, const int* JA, const int* agg, int nagg, int* cIA);
void orc_rap_agg_fill(int nrow, const int* IA, const int* JA, const double* val,
                      const int* agg, int nagg, const int* cIA, int* cJA, double* cval);

/* solver parameter block (subset of FASP itsolver_param + AMG_param actually read) */
typedef struct {
  double tol;            /* itsolver_tol */
  int    maxit;          /* itsolver_maxit */
  int    restart;        /* itsolver_restart */
  int    max_levels;     /* AMG_levels */
  int    coarse_dof;     /* AMG_coarse_dof */
  double strong_coupled; /* AMG_strong_coupled */
  int    max_aggregation;/* AMG_max_aggregation */
  double tentative_smooth;/* AMG_tentative_smooth (>0 => threshold halves per level) */
  double amg_tol;        /* AMG_tol (coarse solve) */
  int    presmooth, postsmooth;
  int    print_level;
  /* mirror mode (to replay the GPU's hierarchy on the CPU): */
  int    coarse_dense;   /* 1: dense LU on the coarsest level instead of GMRES(25) */
  int    ortho_cgs;      /* 1: classical Gram-Schmidt with DGKS re-orthogonalisation (GPU Krylov) */
  int    cycle_type;     /* 1 V-cycle (bsr.dat); else K-cycle (nonlinear AMLI) on coarse levels 1..kcycle_levels */
  int    kcycle_levels;
} orc_solver_param;
void orc_default_param(orc_solver_param* p); /* benchmarks/pnp_exact_solution/bsr.dat */

typedef struct orc_amg orc_amg;
/* FASP fasp_amg_setup_ua_bsr restated. ext_nlevels>0 replays an external hierarchy:
 * ext_agg[l] (length ROW(l)) and ext_order[l] (GS row order, may be NULL). */
orc_amg* orc_amg_setup(int nrow, const int* IA, const int* JA, const double* val,
                       const orc_solver_param* p,
                       int ext_nlevels, const int* const* ext_agg, const int* ext_nagg,
                       const int* const* ext_order);
int  orc_amg_nlevels(const orc_amg* h);
void orc_amg_level_info(const orc_amg* h, int l, int* nrow, int* nnzb, int* nagg);
void orc_amg_level_matrix(const orc_amg* h, int l, int* IA, int* JA, double* val);
void orc_amg_level_agg(const orc_amg* h, int l, int* agg);
void orc_amg_vcycle(orc_amg* h, const double* r, double* z); /* z = M^{-1} r, one V(1,1) */
void orc_amg_free(orc_amg* h);
/* fasp_solver_dbsr_pvfgmres restated; returns iterations >=0 or -1 (ERROR_SOLVER_MAXIT) etc.
 * relres_hist (may be NULL) gets min(its,hist_len) entries */
int  orc_pvfgmres(int nrow, const int* IA, const int* JA, const double* val,
                  const double* b, double* x, orc_amg* precond, const orc_solver_param* p,
                  double* relres_hist, int hist_len);
/* fasp_solver_dbsr_krylov_amg restated: setup + pvfgmres + free */
int  orc_solver_dbsr_krylov_amg(int nrow, const int* IA, const int* JA, const double* val,
                                const double* b, double* x, const orc_solver_param* p);

/* ---- block ILU(k), fasp_solver_dbsr_krylov_ilu restated (benchmarks/pnp_diode/linear_pnp.cpp:103-109,
 * bsr.dat:34-38: ILU_type 2 = ILUk, ILU_lfil 2); oracle/ilu_restated.cpp ---- */
typedef struct orc_ilu orc_ilu;
int      orc_ilu_symbolic_count(int n, const int* IA, const int* JA, int lfil, int* fIA /*n+1*/);
orc_ilu* orc_ilu_setup(int n, const int* IA, const int* JA, const double* val, int lfil);
int      orc_ilu_nnz(const orc_ilu* F);
void     orc_ilu_factors(const orc_ilu* F, int* IA, int* JA, double* val, double* dinv);
void     orc_ilu_apply(const orc_ilu* F, const double* r, double* z);
void     orc_ilu_free(orc_ilu* F);
int      orc_solver_dbsr_krylov_ilu(int nrow, const int* IA, const int* JA, const double* val, const double* b, double* x,
                                    const orc_solver_param* p, int lfil);

/* ---- Newton driver mirroring benchmarks/pnp_exact_solution/main.cpp:138-356 ---- */
typedef struct {
  int    newton_its;       /* Newton_Status::iteration - 1 at exit */
  int    converged;
  double initial_residual, initial_max_residual;
  double rel_res[32], max_res[32];   /* after each Newton step */
  int    krylov_its[32];
  double t_assemble, t_solve, t_residual; /* seconds, summed */
} orc_newton_result;
/* coefficient / initial-guess generator for the pnp_exact_solution physics (main.cpp:24-135) */
void orc_exact_problem(int nv, const double* xyz, double* eps, double* fixed, double* diff,
                       double* val, double* reac, double* uu_init, double* uu_exact);
int  orc_newton_exact(int nx, int ny, int nz, double lx, double ly, double lz,
                      const orc_solver_param* p, int max_newton, double rel_tol, double max_tol,
                      int use_eafe, double* uu_out /*N or NULL*/, orc_newton_result* res);

#ifdef __cplusplus
}
#endif
#endif

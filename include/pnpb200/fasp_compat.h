/* pnp-b200 — FASP-compatible C surface (drop-in boundary B.1 of SURVEY.md §8b).
 *
 * The reference includes FASP through
 *     extern "C" { #include "fasp.h"  #include "fasp_functs.h" }
 * (include/pde.h:11-14, benchmarks/pnp_exact_solution/linear_pnp.cpp:10-13). This header
 * declares the subset of that surface the reference's hot path touches, with the same type
 * names, field names and calling conventions, so linear_pnp.cpp / pde.cpp compile against it
 * unchanged and link libpnpb200.so instead of libfasp.a. Every solver entry point below runs
 * on the GPU (sm_100a); there is no CPU fallback — without a usable CUDA device they return
 * ERROR_DEVICE.
 *
 * Plain C ABI: pointers and sizes only. INT is 32-bit like upstream FASP.
 */
#ifndef PNPB200_FASP_COMPAT_H
#define PNPB200_FASP_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int    INT;
typedef short  SHORT;
typedef double REAL;

/* ---- status codes (FASP fasp_const.h names; used at src/domain.cpp:19,63,96, src/pde.cpp:586) */
#define FASP_SUCCESS         0
#define ERROR_OPEN_FILE    -10
#define ERROR_WRONG_FILE   -11
#define ERROR_INPUT_PAR    -13
#define ERROR_ALLOC_MEM    -15
#define ERROR_DATA_STRUCTURE -16
#define ERROR_AMG_COARSEING -23
#define ERROR_SOLVER_TYPE  -40
#define ERROR_SOLVER_MAXIT -48
#define ERROR_SOLVER_EXIT  -49
#define ERROR_SOLVER_MISC  -45   /* FASP: "unknown solver runtime error" — singular coarsest level */
#define ERROR_UNKNOWN      -99
#define ERROR_DEVICE      -120   /* pnp-b200: CUDA device / kernel failure (no CPU fallback) */

/* ---- enumerations the .dat files use (benchmarks/pnp_exact_solution/bsr.dat) */
#define SOLVER_CG 1
#define SOLVER_BiCGstab 2
#define SOLVER_MinRes 3
#define SOLVER_GMRES 4
#define SOLVER_VGMRES 5
#define SOLVER_VFGMRES 6
#define SOLVER_GCG 7
#define PREC_NULL 0
#define PREC_DIAG 1
#define PREC_AMG 2
#define PREC_FMG 3
#define PREC_ILU 4
#define PREC_SCHWARZ 5
#define STOP_REL_RES 1
#define STOP_REL_PRECRES 2
#define STOP_MOD_REL_RES 3
#define CLASSIC_AMG 1
#define SA_AMG 2
#define UA_AMG 3
#define V_CYCLE 1
#define W_CYCLE 2
#define AMLI_CYCLE 3
#define NL_AMLI_CYCLE 4
#define SMOOTHER_JACOBI 1
#define SMOOTHER_GS 2
#define SMOOTHER_SGS 3
#define SMOOTHER_CG 4
#define SMOOTHER_SOR 5
#define SMOOTHER_SSOR 6
#define SMOOTHER_GSOR 7
#define SMOOTHER_SGSOR 8
#define SMOOTHER_POLY 9
#define SMOOTHER_L1DIAG 10
#define PAIRWISE 1
#define VMB 2
#define ON 1
#define OFF 0

/* ---- matrices / vectors: layouts of upstream FASP (fields used at src/pde.cpp:624-629,
 *      :675-676; benchmarks/pnp_stokes/linear_pnp_ns.cpp:60-109) */
typedef struct dCSRmat { INT row, col, nnz; INT *IA, *JA; REAL *val; } dCSRmat;
typedef struct dBSRmat {
  INT ROW, COL, NNZ, nb;
  INT storage_manner;           /* 0: row-major blocks */
  REAL *val; INT *IA, *JA;
} dBSRmat;
typedef struct block_dCSRmat { INT brow, bcol; dCSRmat **blocks; } block_dCSRmat;
typedef struct dvector { INT row; REAL *val; } dvector;
typedef struct ivector { INT row; INT  *val; } ivector;

/* ---- parameter blocks. The reference treats these as opaque values it copies
 *      (linear_pnp.cpp:61-62); only input_param.print_level is ever touched
 *      (tests/eafe_tests/test_eafe.cpp:129). Field names follow FASP. */
typedef struct input_param {
  SHORT print_level, output_type;
  char  inifile[256], workdir[256];
  INT   problem_num;
  SHORT solver_type, precond_type, stop_type;
  REAL  itsolver_tol;
  INT   itsolver_maxit, restart;
  SHORT ILU_type; INT ILU_lfil; REAL ILU_droptol, ILU_relax, ILU_permtol;
  INT   Schwarz_mmsize, Schwarz_maxlvl, Schwarz_type, Schwarz_blksolver;
  SHORT AMG_type, AMG_levels, AMG_cycle_type, AMG_smoother, AMG_smooth_order;
  REAL  AMG_relaxation;
  SHORT AMG_polynomial_degree, AMG_presmooth_iter, AMG_postsmooth_iter;
  INT   AMG_coarse_dof;
  REAL  AMG_tol;
  INT   AMG_maxit;
  SHORT AMG_ILU_levels, AMG_coarse_solver, AMG_coarse_scaling, AMG_amli_degree, AMG_nl_amli_krylov_type;
  INT   AMG_Schwarz_levels;
  SHORT AMG_coarsening_type, AMG_aggregation_type, AMG_interpolation_type;
  REAL  AMG_strong_threshold, AMG_truncation_threshold, AMG_max_row_sum;
  INT   AMG_aggressive_level, AMG_aggressive_path, AMG_pair_number;
  REAL  AMG_quality_bound;
  REAL  AMG_strong_coupled;
  INT   AMG_max_aggregation;
  REAL  AMG_tentative_smooth;
  SHORT AMG_smooth_filter;
} input_param;

typedef struct itsolver_param {
  SHORT itsolver_type, precond_type, stop_type;
  INT   maxit;
  REAL  tol;
  INT   restart;
  SHORT print_level;
} itsolver_param;

typedef struct AMG_param {
  SHORT AMG_type, print_level;
  INT   maxit;
  REAL  tol;
  SHORT max_levels;
  INT   coarse_dof;
  SHORT cycle_type;
  REAL  quality_bound;
  SHORT smoother, smooth_order, presmooth_iter, postsmooth_iter;
  REAL  relaxation;
  SHORT polynomial_degree, coarse_solver, coarse_scaling, amli_degree;
  REAL *amli_coef;
  SHORT nl_amli_krylov_type, coarsening_type, aggregation_type, interpolation_type;
  REAL  strong_threshold, max_row_sum, truncation_threshold;
  INT   aggressive_level, aggressive_path, pair_number;
  REAL  strong_coupled;
  INT   max_aggregation;
  REAL  tentative_smooth;
  SHORT smooth_filter;
  SHORT ILU_levels, ILU_type; INT ILU_lfil; REAL ILU_droptol, ILU_relax, ILU_permtol;
  INT   Schwarz_levels, Schwarz_mmsize, Schwarz_maxlvl, Schwarz_type, Schwarz_blksolver;
} AMG_param;

typedef struct ILU_param {
  SHORT print_level, ILU_type; INT ILU_lfil; REAL ILU_droptol, ILU_relax, ILU_permtol;
} ILU_param;

typedef struct Schwarz_param {
  SHORT print_level, Schwarz_type; INT Schwarz_maxlvl, Schwarz_mmsize, Schwarz_blksolver;
} Schwarz_param;

/* ---- fasp4ns parameter blocks (benchmarks/pnp_stokes/main.cpp:78-86). fasp4ns is not in the
 *      reference tree; the reference only declares these, fills them with fasp_ns_param_input /
 *      fasp_ns_param_init and passes them on by value (linear_pnp_ns.cpp:52-56,183-193), so their
 *      layout is private to this library: three FASP parameter sets — the outer Stokes iteration,
 *      the velocity block (keys ending in _v in ns.dat) and the pressure block (_p). */
typedef struct input_ns_param { input_param ns, v, p; } input_ns_param;
typedef struct itsolver_ns_param {
  SHORT itsolver_type, precond_type, stop_type;
  INT   maxit;
  REAL  tol;
  INT   restart;
  SHORT print_level;
  SHORT itsolver_type_v, precond_type_v; INT pre_maxit_v; REAL pre_tol_v; INT pre_restart_v;
  SHORT itsolver_type_p, precond_type_p; INT pre_maxit_p; REAL pre_tol_p; INT pre_restart_p;
} itsolver_ns_param;
typedef struct AMG_ns_param { AMG_param param_v, param_p; } AMG_ns_param;

/* ---- parameter input (benchmarks/pnp_exact_solution/main.cpp:170-177) */
void  fasp_param_input(const char *filenm, input_param *inparam);
void  fasp_param_input_init(input_param *inparam);
void  fasp_param_init(const input_param *inparam, itsolver_param *itsparam, AMG_param *amgparam,
                      ILU_param *iluparam, Schwarz_param *schparam);
void  fasp_param_solver_init(itsolver_param *itsparam);
void  fasp_param_amg_init(AMG_param *amgparam);
/* benchmarks/pnp_stokes/main.cpp:85-86 (fasp4ns): ns.dat = outer / _v / _p parameter sets */
void  fasp_ns_param_input(const char *filenm, input_ns_param *inparam);
void  fasp_ns_param_init(input_ns_param *inparam, itsolver_ns_param *itsparam, AMG_ns_param *amgparam,
                         ILU_param *iluparam, Schwarz_param *schparam);

/* ---- format conversion (linear_pnp.cpp:85): CSR -> BSR(nb), device kernels, returns by value */
dBSRmat fasp_format_dcsr_dbsr(const dCSRmat *A, const INT nb);

/* ---- solvers */
/* linear_pnp.cpp:101,150 — BSR unsmoothed-aggregation AMG + variable-restart FGMRES, all on
 * the GPU. A, b read-only; x in/out (initial guess). Returns iterations >= 0 or ERROR_* < 0;
 * on failure x holds the last iterate. */
INT fasp_solver_dbsr_krylov_amg(dBSRmat *A, dvector *b, dvector *x, itsolver_param *itparam,
                                AMG_param *amgparam);
/* tests/eafe_tests/test_eafe.cpp:170 — unpreconditioned Krylov (FGMRES) on a CSR matrix, GPU */
INT fasp_solver_dcsr_krylov(dCSRmat *A, dvector *b, dvector *x, itsolver_param *itparam);
/* benchmarks/pnp_diode/linear_pnp.cpp:103-109 — block ILU(k)-preconditioned FGMRES on the GPU, k = ILU_lfil
 * (pnp_diode/bsr.dat:34-38: ILU_type 2 = ILUk, lfil 2); same return convention. csrc/ilu.cu */
INT fasp_solver_dbsr_krylov_ilu(dBSRmat *A, dvector *b, dvector *x, itsolver_param *itparam,
                                ILU_param *iluparam);

/* benchmarks/pnp_stokes/linear_pnp_ns.cpp:183-193 (fasp4ns) — the coupled PNP-Stokes system as a
 * 2x2 block_dCSRmat {A_pp, A_ps, A_sp, A_ss} (PNP dofs interleaved 3 per vertex; Stokes dofs =
 * num_velocity velocity dofs followed by num_pressure pressure dofs, linear_pnp_ns.cpp:69-107).
 * All on the GPU: outer FGMRES (itparam) with the block upper-triangular preconditioner
 * (bcsr.dat precond_type 33): Stokes block by FGMRES (itparam_stokes) preconditioned with an AMG
 * cycle on the velocity block and the diagonal SIMPLE Schur complement on the pressure, then the
 * PNP block by the BSR(3) AMG-FGMRES solve of fasp_solver_dbsr_krylov_amg (itparam_pnp,
 * amgparam_pnp). b, x ordered [PNP; velocity; pressure]; x in/out (initial guess). Returns the
 * outer iteration count >= 0 or ERROR_* < 0 (x = last iterate). DESIGN.md §2 f-2. */
INT fasp_solver_bdcsr_krylov_pnp_stokes(block_dCSRmat *A, dvector *b, dvector *x, itsolver_param *itparam,
                                        itsolver_param *itparam_pnp, AMG_param *amgparam_pnp,
                                        itsolver_ns_param *itparam_stokes, AMG_ns_param *amgparam_stokes,
                                        const INT num_velocity, const INT num_pressure);

/* ---- BLAS-level entry used by Linear_PNP::fasp_test_solver's callers (GPU) */
void fasp_blas_dbsr_mxv(const dBSRmat *A, const REAL *x, REAL *y);
void fasp_blas_dbsr_aAxpy(const REAL alpha, const dBSRmat *A, const REAL *x, REAL *y);

/* ---- host-side memory / vector helpers (linear_pnp.cpp:88,93,178-181; linear_pnp_ns.cpp:63-105) */
void  fasp_dvec_alloc(const INT m, dvector *u);
void  fasp_dvec_set(INT n, dvector *x, REAL val);
void  fasp_dvec_free(dvector *u);
void  fasp_ivec_alloc(const INT m, ivector *u);
void  fasp_ivec_free(ivector *u);
void  fasp_dcsr_free(dCSRmat *A);
void  fasp_dbsr_free(dBSRmat *A);
void  fasp_bdcsr_free(block_dCSRmat *A);
void *fasp_mem_calloc(unsigned int size, unsigned int type);
void  fasp_mem_free(void *mem);
void  fasp_mem_check(void *ptr, const char *message, INT ERR);
void  fasp_chkerr(const SHORT status, const char *fctname);
/* benchmarks/pnp_stokes/linear_pnp_ns.cpp:150-157: sub-matrix by index lists (host utility) */
SHORT fasp_dcsr_getblk(dCSRmat *A, INT *Is, INT *Js, INT m, INT n, dCSRmat *B);
/* debug writers (only in comments in the reference) */
void  fasp_dvec_write(const char *filename, dvector *vec);
void  fasp_ivec_write(const char *filename, ivector *vec);
void  fasp_dcoo_write(const char *filename, dCSRmat *A);

#ifdef __cplusplus
}
#endif
#endif

// pnp-b200 — host-side FASP-compatible utilities (include/pnpb200/fasp_compat.h): the
// `key = value` parameter reader used at benchmarks/pnp_exact_solution/main.cpp:170-177
// (same grammar as the in-tree clone src/domain.cpp:108-113: '%', '[' and '|' start comments),
// parameter defaults / copies, and the small memory + vector helpers of
// benchmarks/pnp_exact_solution/linear_pnp.cpp:88,93,178-181. No numerics live here.
#include "pnpb200/fasp_compat.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

void fasp_chkerr(const SHORT status, const char* fctname)
{
  if (status >= 0) return;
  fprintf(stderr, "### ERROR: %s failed with status %d\n", fctname ? fctname : "?", (int)status);
  exit(status);      // upstream FASP behaviour (src/pde.cpp:585-587 relies on it)
}

void* fasp_mem_calloc(unsigned int size, unsigned int type)
{
  const size_t tot = (size_t)size * type;
  if (tot == 0) return nullptr;
  void* p = calloc(size, type);
  if (!p) { fprintf(stderr, "### ERROR: fasp_mem_calloc failed for %zu bytes\n", tot); }
  return p;
}
void fasp_mem_free(void* mem) { free(mem); }
void fasp_mem_check(void* ptr, const char* message, INT ERR)
{
  if (ptr == nullptr) { fprintf(stderr, "### ERROR: %s\n", message ? message : "null pointer"); exit(ERR); }
}

void fasp_dvec_alloc(const INT m, dvector* u) { u->row = m; u->val = (REAL*)fasp_mem_calloc(m, sizeof(REAL)); }
void fasp_dvec_set(INT n, dvector* x, REAL val)
{
  if (n > 0) x->row = n; else n = x->row;
  if (val == 0.0) memset(x->val, 0, sizeof(REAL) * (size_t)n);
  else for (INT i = 0; i < n; ++i) x->val[i] = val;
}
void fasp_dvec_free(dvector* u) { if (!u) return; free(u->val); u->val = nullptr; u->row = 0; }
void fasp_ivec_alloc(const INT m, ivector* u) { u->row = m; u->val = (INT*)fasp_mem_calloc(m, sizeof(INT)); }
void fasp_ivec_free(ivector* u) { if (!u) return; free(u->val); u->val = nullptr; u->row = 0; }
void fasp_dcsr_free(dCSRmat* A)
{
  if (!A) return;
  free(A->IA); free(A->JA); free(A->val);
  A->IA = A->JA = nullptr; A->val = nullptr; A->row = A->col = A->nnz = 0;
}
void fasp_dbsr_free(dBSRmat* A)
{
  if (!A) return;
  free(A->IA); free(A->JA); free(A->val);
  A->IA = A->JA = nullptr; A->val = nullptr; A->ROW = A->COL = A->NNZ = A->nb = 0;
}
void fasp_bdcsr_free(block_dCSRmat* A)
{
  if (!A || !A->blocks) return;
  for (INT i = 0; i < A->brow * A->bcol; ++i) { fasp_dcsr_free(A->blocks[i]); A->blocks[i] = nullptr; }
  free(A->blocks); A->blocks = nullptr;
}

// sub-matrix B = A(Is, Js), benchmarks/pnp_stokes/linear_pnp_ns.cpp:150-157 [EXT fasp_dcsr_getblk]
SHORT fasp_dcsr_getblk(dCSRmat* A, INT* Is, INT* Js, INT m, INT n, dCSRmat* B)
{
  std::vector<INT> col_flag(A->col, 0);
  for (INT j = 0; j < n; ++j) col_flag[Js[j]] = j + 1;
  B->row = m; B->col = n;
  B->IA = (INT*)fasp_mem_calloc(m + 1, sizeof(INT));
  INT nnz = 0;
  for (INT i = 0; i < m; ++i)
    for (INT k = A->IA[Is[i]]; k < A->IA[Is[i] + 1]; ++k) if (col_flag[A->JA[k]] > 0) ++nnz;
  B->JA = (INT*)fasp_mem_calloc(nnz ? nnz : 1, sizeof(INT));
  B->val = (REAL*)fasp_mem_calloc(nnz ? nnz : 1, sizeof(REAL));
  nnz = 0;
  for (INT i = 0; i < m; ++i) {
    for (INT k = A->IA[Is[i]]; k < A->IA[Is[i] + 1]; ++k) {
      const INT j = A->JA[k];
      if (col_flag[j] > 0) { B->JA[nnz] = col_flag[j] - 1; B->val[nnz] = A->val[k]; ++nnz; }
    }
    B->IA[i + 1] = nnz;
  }
  B->nnz = nnz;
  return FASP_SUCCESS;
}

void fasp_dvec_write(const char* filename, dvector* vec)
{
  FILE* fp = fopen(filename, "w");
  if (!fp) { fasp_chkerr(ERROR_OPEN_FILE, "fasp_dvec_write"); return; }
  fprintf(fp, "%d\n", vec->row);
  for (INT i = 0; i < vec->row; ++i) fprintf(fp, "%0.15e\n", vec->val[i]);
  fclose(fp);
}
void fasp_ivec_write(const char* filename, ivector* vec)
{
  FILE* fp = fopen(filename, "w");
  if (!fp) { fasp_chkerr(ERROR_OPEN_FILE, "fasp_ivec_write"); return; }
  fprintf(fp, "%d\n", vec->row);
  for (INT i = 0; i < vec->row; ++i) fprintf(fp, "%d %d\n", i, vec->val[i]);
  fclose(fp);
}
void fasp_dcoo_write(const char* filename, dCSRmat* A)
{
  FILE* fp = fopen(filename, "w");
  if (!fp) { fasp_chkerr(ERROR_OPEN_FILE, "fasp_dcoo_write"); return; }
  fprintf(fp, "%d  %d  %d\n", A->row, A->col, A->nnz);
  for (INT i = 0; i < A->row; ++i)
    for (INT j = A->IA[i]; j < A->IA[i + 1]; ++j) fprintf(fp, "%d  %d  %0.15e\n", i, A->JA[j], A->val[j]);
  fclose(fp);
}

// ---- parameters ------------------------------------------------------------------------------
void fasp_param_input_init(input_param* in)
{
  memset(in, 0, sizeof(*in));
  strcpy(in->workdir, "../data/");
  in->print_level = 0; in->output_type = 0;
  in->problem_num = 10; in->solver_type = SOLVER_CG; in->precond_type = PREC_AMG; in->stop_type = STOP_REL_RES;
  in->itsolver_tol = 1e-6; in->itsolver_maxit = 500; in->restart = 25;
  in->ILU_type = 1; in->ILU_lfil = 0; in->ILU_droptol = 0.001; in->ILU_relax = 0; in->ILU_permtol = 0.0;
  in->Schwarz_mmsize = 200; in->Schwarz_maxlvl = 2; in->Schwarz_type = 1; in->Schwarz_blksolver = 0;
  in->AMG_type = CLASSIC_AMG; in->AMG_levels = 20; in->AMG_cycle_type = V_CYCLE; in->AMG_smoother = SMOOTHER_GS;
  in->AMG_smooth_order = 1; in->AMG_presmooth_iter = 1; in->AMG_postsmooth_iter = 1; in->AMG_relaxation = 1.0;
  in->AMG_coarse_dof = 500; in->AMG_coarse_solver = 0; in->AMG_tol = 1e-6; in->AMG_maxit = 1;
  in->AMG_ILU_levels = 0; in->AMG_Schwarz_levels = 0; in->AMG_coarse_scaling = OFF; in->AMG_amli_degree = 1;
  in->AMG_nl_amli_krylov_type = 2; in->AMG_polynomial_degree = 3;
  in->AMG_coarsening_type = 1; in->AMG_interpolation_type = 1; in->AMG_max_row_sum = 0.9;
  in->AMG_strong_threshold = 0.3; in->AMG_truncation_threshold = 0.2; in->AMG_aggressive_level = 0; in->AMG_aggressive_path = 1;
  in->AMG_aggregation_type = PAIRWISE; in->AMG_quality_bound = 8.0; in->AMG_pair_number = 2;
  in->AMG_strong_coupled = 0.25; in->AMG_max_aggregation = 9; in->AMG_tentative_smooth = 0.67; in->AMG_smooth_filter = ON;
}

static bool on_off(const std::string& v) { return v == "ON" || v == "on" || v == "On"; }

} // extern "C"

// one `key = value` pair into an input_param: 0 = taken, 1 = bad value, 2 = unknown key
static int set_key(input_param* in, const char* key, const char* val)
{
  const std::string k(key), v(val);
  bool bad = false;
  auto I = [&]() { return atoi(val); };
  auto R = [&]() { return atof(val); };
    if (k == "workdir") strncpy(in->workdir, val, sizeof(in->workdir) - 1);
    else if (k == "problem_num") in->problem_num = I();
    else if (k == "print_level") in->print_level = (SHORT)I();
    else if (k == "output_type") in->output_type = (SHORT)I();
    else if (k == "solver_type") in->solver_type = (SHORT)I();
    else if (k == "stop_type") in->stop_type = (SHORT)I();
    else if (k == "precond_type") in->precond_type = (SHORT)I();
    else if (k == "itsolver_tol") in->itsolver_tol = R();
    else if (k == "itsolver_maxit") in->itsolver_maxit = I();
    else if (k == "itsolver_restart") in->restart = I();
    else if (k == "ILU_type") in->ILU_type = (SHORT)I();
    else if (k == "ILU_lfil") in->ILU_lfil = I();
    else if (k == "ILU_droptol") in->ILU_droptol = R();
    else if (k == "ILU_relax") in->ILU_relax = R();
    else if (k == "ILU_permtol") in->ILU_permtol = R();
    else if (k == "Schwarz_mmsize") in->Schwarz_mmsize = I();
    else if (k == "Schwarz_maxlvl") in->Schwarz_maxlvl = I();
    else if (k == "Schwarz_type") in->Schwarz_type = I();
    else if (k == "Schwarz_blksolver") in->Schwarz_blksolver = I();
    else if (k == "AMG_type") {
      if (v == "C" || v == "c") in->AMG_type = CLASSIC_AMG;
      else if (v == "SA" || v == "sa") in->AMG_type = SA_AMG;
      else if (v == "UA" || v == "ua") in->AMG_type = UA_AMG;
      else bad = true;
    }
    else if (k == "AMG_cycle_type") {
      if (v == "V" || v == "v") in->AMG_cycle_type = V_CYCLE;
      else if (v == "W" || v == "w") in->AMG_cycle_type = W_CYCLE;
      else if (v == "A" || v == "a") in->AMG_cycle_type = AMLI_CYCLE;
      else if (v == "NA" || v == "na") in->AMG_cycle_type = NL_AMLI_CYCLE;
      else bad = true;
    }
    else if (k == "AMG_smoother") {
      if (v == "JACOBI" || v == "jacobi") in->AMG_smoother = SMOOTHER_JACOBI;
      else if (v == "GS" || v == "gs") in->AMG_smoother = SMOOTHER_GS;
      else if (v == "SGS" || v == "sgs") in->AMG_smoother = SMOOTHER_SGS;
      else if (v == "CG" || v == "cg") in->AMG_smoother = SMOOTHER_CG;
      else if (v == "SOR" || v == "sor") in->AMG_smoother = SMOOTHER_SOR;
      else if (v == "SSOR" || v == "ssor") in->AMG_smoother = SMOOTHER_SSOR;
      else if (v == "GSOR" || v == "gsor") in->AMG_smoother = SMOOTHER_GSOR;
      else if (v == "SGSOR" || v == "sgsor") in->AMG_smoother = SMOOTHER_SGSOR;
      else if (v == "POLY" || v == "poly") in->AMG_smoother = SMOOTHER_POLY;
      else if (v == "L1DIAG" || v == "l1diag") in->AMG_smoother = SMOOTHER_L1DIAG;
      else bad = true;
    }
    else if (k == "AMG_smooth_order") in->AMG_smooth_order = (v == "NO" || v == "no") ? 0 : 1;
    else if (k == "AMG_coarse_scaling") in->AMG_coarse_scaling = on_off(v) ? ON : OFF;
    else if (k == "AMG_smooth_filter") in->AMG_smooth_filter = on_off(v) ? ON : OFF;
    else if (k == "AMG_levels") in->AMG_levels = (SHORT)I();
    else if (k == "AMG_tol") in->AMG_tol = R();
    else if (k == "AMG_maxit") in->AMG_maxit = I();
    else if (k == "AMG_coarse_dof") in->AMG_coarse_dof = I();
    else if (k == "AMG_coarse_solver") in->AMG_coarse_solver = (SHORT)I();
    else if (k == "AMG_amli_degree") in->AMG_amli_degree = (SHORT)I();
    else if (k == "AMG_nl_amli_krylov_type") in->AMG_nl_amli_krylov_type = (SHORT)I();
    else if (k == "AMG_ILU_levels") in->AMG_ILU_levels = (SHORT)I();
    else if (k == "AMG_Schwarz_levels" || k == "AMG_schwarz_levels") in->AMG_Schwarz_levels = I();
    else if (k == "AMG_relaxation") in->AMG_relaxation = R();
    else if (k == "AMG_polynomial_degree") in->AMG_polynomial_degree = (SHORT)I();
    else if (k == "AMG_presmooth_iter") in->AMG_presmooth_iter = (SHORT)I();
    else if (k == "AMG_postsmooth_iter") in->AMG_postsmooth_iter = (SHORT)I();
    else if (k == "AMG_coarsening_type") in->AMG_coarsening_type = (SHORT)I();
    else if (k == "AMG_interpolation_type") in->AMG_interpolation_type = (SHORT)I();
    else if (k == "AMG_strong_threshold") in->AMG_strong_threshold = R();
    else if (k == "AMG_truncation_threshold") in->AMG_truncation_threshold = R();
    else if (k == "AMG_max_row_sum") in->AMG_max_row_sum = R();
    else if (k == "AMG_aggressive_level") in->AMG_aggressive_level = I();
    else if (k == "AMG_aggressive_path") in->AMG_aggressive_path = I();
    else if (k == "AMG_aggregation_type") in->AMG_aggregation_type = (SHORT)I();
    else if (k == "AMG_pair_number") in->AMG_pair_number = I();
    else if (k == "AMG_quality_bound") in->AMG_quality_bound = R();
    else if (k == "AMG_strong_coupled") in->AMG_strong_coupled = R();
    else if (k == "AMG_max_aggregation") in->AMG_max_aggregation = I();
    else if (k == "AMG_tentative_smooth") in->AMG_tentative_smooth = R();
    else return 2;
  return bad ? 1 : 0;
}

// shared reader: every non-comment `key = value` line goes to sink(key, value); false on a malformed line
template <typename F>
static bool read_dat(const char* filenm, const char* who, F&& sink)
{
  FILE* fp = fopen(filenm, "r");
  if (!fp) { fprintf(stderr, "### ERROR: %s cannot be opened\n", filenm); fasp_chkerr(ERROR_OPEN_FILE, who); return false; }
  char line[1024];
  bool ok = true;
  while (fgets(line, sizeof line, fp)) {
    char key[512], eq[16], val[512];
    if (sscanf(line, "%511s", key) != 1) continue;
    if (key[0] == '%' || key[0] == '[' || key[0] == '|' || key[0] == '#') continue;
    if (sscanf(line, "%511s %15s %511s", key, eq, val) != 3 || strcmp(eq, "=") != 0) { ok = false; continue; }
    if (!sink(key, val)) ok = false;
  }
  fclose(fp);
  return ok;
}

static bool params_sane(const input_param* in)
{
  return !(in->itsolver_tol <= 0 || in->itsolver_maxit <= 0 || in->AMG_levels < 0 || in->AMG_coarse_dof <= 0);
}

extern "C" {

void fasp_param_input(const char* filenm, input_param* in)
{
  fasp_param_input_init(in);
  strncpy(in->inifile, filenm, sizeof(in->inifile) - 1);
  const bool ok = read_dat(filenm, "fasp_param_input", [&](const char* key, const char* val) {
    const int rc = set_key(in, key, val);
    if (rc == 2) fprintf(stderr, "### WARNING: Unknown input keyword %s!\n", key);   // ignored with a note, like upstream
    return rc != 1;
  });
  if (!ok || !params_sane(in)) fasp_chkerr(ERROR_INPUT_PAR, "fasp_param_input");
}

// benchmarks/pnp_stokes/main.cpp:85: ns.dat holds three parameter sets in one file — the outer
// Stokes iteration (plain keys), the velocity block (keys ending in _v) and the pressure block (_p)
void fasp_ns_param_input(const char* filenm, input_ns_param* in)
{
  fasp_param_input_init(&in->ns); fasp_param_input_init(&in->v); fasp_param_input_init(&in->p);
  strncpy(in->ns.inifile, filenm, sizeof(in->ns.inifile) - 1);
  const bool ok = read_dat(filenm, "fasp_ns_param_input", [&](const char* key, const char* val) {
    std::string k(key);
    input_param* dst = &in->ns;
    if (k.size() > 2 && k.compare(k.size() - 2, 2, "_v") == 0) { dst = &in->v; k.resize(k.size() - 2); }
    else if (k.size() > 2 && k.compare(k.size() - 2, 2, "_p") == 0) { dst = &in->p; k.resize(k.size() - 2); }
    const int rc = set_key(dst, k.c_str(), val);
    if (rc == 2) fprintf(stderr, "### WARNING: Unknown input keyword %s!\n", key);
    return rc != 1;
  });
  in->v.print_level = in->p.print_level = in->ns.print_level;
  if (!ok || !params_sane(&in->ns) || !params_sane(&in->v) || !params_sane(&in->p)) fasp_chkerr(ERROR_INPUT_PAR, "fasp_ns_param_input");
}

// benchmarks/pnp_stokes/main.cpp:86
void fasp_ns_param_init(input_ns_param* in, itsolver_ns_param* its, AMG_ns_param* amg, ILU_param* ilu, Schwarz_param* sch)
{
  if (its) {
    memset(its, 0, sizeof(*its));
    itsolver_param t;
    fasp_param_init(in ? &in->ns : nullptr, &t, nullptr, nullptr, nullptr);
    its->itsolver_type = t.itsolver_type; its->precond_type = t.precond_type; its->stop_type = t.stop_type;
    its->maxit = t.maxit; its->tol = t.tol; its->restart = t.restart; its->print_level = t.print_level;
    fasp_param_init(in ? &in->v : nullptr, &t, nullptr, nullptr, nullptr);
    its->itsolver_type_v = t.itsolver_type; its->precond_type_v = t.precond_type; its->pre_maxit_v = t.maxit;
    its->pre_tol_v = t.tol; its->pre_restart_v = t.restart;
    fasp_param_init(in ? &in->p : nullptr, &t, nullptr, nullptr, nullptr);
    its->itsolver_type_p = t.itsolver_type; its->precond_type_p = t.precond_type; its->pre_maxit_p = t.maxit;
    its->pre_tol_p = t.tol; its->pre_restart_p = t.restart;
  }
  if (amg) {
    fasp_param_init(in ? &in->v : nullptr, nullptr, &amg->param_v, nullptr, nullptr);
    fasp_param_init(in ? &in->p : nullptr, nullptr, &amg->param_p, nullptr, nullptr);
  }
  fasp_param_init(in ? &in->ns : nullptr, nullptr, nullptr, ilu, sch);
}

void fasp_param_solver_init(itsolver_param* p)
{
  p->print_level = 0; p->itsolver_type = SOLVER_CG; p->precond_type = PREC_AMG; p->stop_type = STOP_REL_RES;
  p->maxit = 500; p->restart = 25; p->tol = 1e-6;
}

void fasp_param_amg_init(AMG_param* p)
{
  memset(p, 0, sizeof(*p));
  // defaults = benchmarks/pnp_exact_solution/bsr.dat:52-99 (the configuration the hot path runs)
  p->AMG_type = UA_AMG; p->print_level = 0; p->maxit = 1; p->tol = 1e-6; p->max_levels = 20; p->coarse_dof = 100;
  p->cycle_type = V_CYCLE; p->quality_bound = 8.0; p->smoother = SMOOTHER_GS; p->smooth_order = 1;
  p->presmooth_iter = 1; p->postsmooth_iter = 1; p->relaxation = 1.1; p->polynomial_degree = 3;
  p->coarse_solver = 0; p->coarse_scaling = OFF; p->amli_degree = 3; p->amli_coef = nullptr; p->nl_amli_krylov_type = 6;
  p->coarsening_type = 1; p->aggregation_type = VMB; p->interpolation_type = 1; p->strong_threshold = 0.25;
  p->max_row_sum = 0.9; p->truncation_threshold = 0.4; p->aggressive_level = 0; p->aggressive_path = 1; p->pair_number = 3;
  p->strong_coupled = 0.08; p->max_aggregation = 20; p->tentative_smooth = 0.67; p->smooth_filter = OFF;
  p->ILU_levels = 0; p->ILU_type = 2; p->ILU_lfil = 3; p->ILU_droptol = 0.01; p->ILU_relax = 0.9; p->ILU_permtol = 0.001;
  p->Schwarz_levels = 0; p->Schwarz_mmsize = 200; p->Schwarz_maxlvl = 2; p->Schwarz_type = 1; p->Schwarz_blksolver = 0;
}

void fasp_param_init(const input_param* in, itsolver_param* its, AMG_param* amg, ILU_param* ilu, Schwarz_param* sch)
{
  if (its) {
    fasp_param_solver_init(its);
    if (in) {
      its->print_level = in->print_level; its->itsolver_type = in->solver_type; its->precond_type = in->precond_type;
      its->stop_type = in->stop_type; its->restart = in->restart; its->tol = in->itsolver_tol; its->maxit = in->itsolver_maxit;
    }
  }
  if (amg) {
    fasp_param_amg_init(amg);
    if (in) {
      amg->AMG_type = in->AMG_type; amg->print_level = in->print_level;
      amg->cycle_type = in->AMG_cycle_type; amg->smoother = in->AMG_smoother; amg->smooth_order = in->AMG_smooth_order;
      amg->relaxation = in->AMG_relaxation; amg->coarse_solver = in->AMG_coarse_solver;
      amg->polynomial_degree = in->AMG_polynomial_degree; amg->presmooth_iter = in->AMG_presmooth_iter;
      amg->postsmooth_iter = in->AMG_postsmooth_iter; amg->coarse_dof = in->AMG_coarse_dof; amg->max_levels = in->AMG_levels;
      amg->tol = in->AMG_tol; amg->maxit = in->AMG_maxit; amg->coarse_scaling = in->AMG_coarse_scaling;
      amg->amli_degree = in->AMG_amli_degree; amg->nl_amli_krylov_type = in->AMG_nl_amli_krylov_type;
      amg->coarsening_type = in->AMG_coarsening_type; amg->interpolation_type = in->AMG_interpolation_type;
      amg->strong_threshold = in->AMG_strong_threshold; amg->truncation_threshold = in->AMG_truncation_threshold;
      amg->max_row_sum = in->AMG_max_row_sum; amg->aggressive_level = in->AMG_aggressive_level;
      amg->aggressive_path = in->AMG_aggressive_path; amg->aggregation_type = in->AMG_aggregation_type;
      amg->pair_number = in->AMG_pair_number; amg->quality_bound = in->AMG_quality_bound;
      amg->strong_coupled = in->AMG_strong_coupled; amg->max_aggregation = in->AMG_max_aggregation;
      amg->tentative_smooth = in->AMG_tentative_smooth; amg->smooth_filter = in->AMG_smooth_filter;
      amg->ILU_levels = in->AMG_ILU_levels; amg->ILU_type = in->ILU_type; amg->ILU_lfil = in->ILU_lfil;
      amg->ILU_droptol = in->ILU_droptol; amg->ILU_relax = in->ILU_relax; amg->ILU_permtol = in->ILU_permtol;
      amg->Schwarz_levels = in->AMG_Schwarz_levels; amg->Schwarz_mmsize = in->Schwarz_mmsize;
      amg->Schwarz_maxlvl = in->Schwarz_maxlvl; amg->Schwarz_type = in->Schwarz_type; amg->Schwarz_blksolver = in->Schwarz_blksolver;
    }
  }
  if (ilu) {
    ilu->print_level = in ? in->print_level : 0; ilu->ILU_type = in ? in->ILU_type : 1; ilu->ILU_lfil = in ? in->ILU_lfil : 0;
    ilu->ILU_droptol = in ? in->ILU_droptol : 0.001; ilu->ILU_relax = in ? in->ILU_relax : 0; ilu->ILU_permtol = in ? in->ILU_permtol : 0.0;
  }
  if (sch) {
    sch->print_level = in ? in->print_level : 0; sch->Schwarz_type = in ? (SHORT)in->Schwarz_type : 1;
    sch->Schwarz_maxlvl = in ? in->Schwarz_maxlvl : 2; sch->Schwarz_mmsize = in ? in->Schwarz_mmsize : 200;
    sch->Schwarz_blksolver = in ? in->Schwarz_blksolver : 0;
  }
}

}

// TEST INFRASTRUCTURE ONLY (oracle/) — see pnp_oracle.h.
// CPU restatement of the FASP path behind fasp_solver_dbsr_krylov_amg
// (call site benchmarks/pnp_exact_solution/linear_pnp.cpp:101-107). FASP itself is an
// un-vendored dependency (libfasp.a, CMakeLists.txt:40-48, version unpinned, API ~1.8-1.9);
// this follows its published algorithms as summarised in SURVEY.md Appendix B:
//   fasp_amg_setup_ua_bsr  : diaginv, L-inf condensed scalar matrix, VMB aggregation,
//                            boolean block prolongator, R = P^T, A_{l+1} = R A_l P
//   fasp_solver_mgcycle_bsr: V(nu1,nu2), block Gauss-Seidel ascending/descending,
//                            coarsest level by GMRES(25) on the CSR form
//   fasp_solver_dbsr_pvfgmres: right-preconditioned flexible GMRES, adaptive restart,
//                            modified Gram-Schmidt, Givens rotations
// "Mirror mode" (ext_* arguments, coarse_dense, ortho_cgs) replays the hierarchy, smoother
// order, coarse solve and orthogonalisation of the CUDA path so its kernels can be checked
// one by one against this code.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <functional>
#include "pnp_oracle.h"
#include "oracle_internal.hpp"

using namespace orc;

namespace {

inline void inv3(const double* a, double* o)
{
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
  const double id = 1.0 / det;
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

void bsr_spmv(int n, const int* IA, const int* JA, const double* val, double alpha, const double* x,
              double beta, double* y)
{
  for (int i = 0; i < n; ++i) {
    double s[3] = {0, 0, 0};
    for (int p = IA[i]; p < IA[i + 1]; ++p) {
      const double* a = val + 9 * (size_t)p;
      const double* xj = x + 3 * (size_t)JA[p];
      s[0] += a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2];
      s[1] += a[3] * xj[0] + a[4] * xj[1] + a[5] * xj[2];
      s[2] += a[6] * xj[0] + a[7] * xj[1] + a[8] * xj[2];
    }
    for (int k = 0; k < 3; ++k) y[3 * (size_t)i + k] = alpha * s[k] + (beta == 0.0 ? 0.0 : beta * y[3 * (size_t)i + k]);
  }
}

void bsr_diaginv(int n, const int* IA, const int* JA, const double* val, double* dinv)
{
  for (int i = 0; i < n; ++i)
    for (int p = IA[i]; p < IA[i + 1]; ++p)
      if (JA[p] == i) inv3(val + 9 * (size_t)p, dinv + 9 * (size_t)i);
}

// fasp_smoother_dbsr_gs_ascend / _descend / _order1 [EXT]:
// x_i = D_i^{-1} (b_i - sum_{j != i} A_ij x_j), rows visited in `order`
void bsr_gs(int n, const int* IA, const int* JA, const double* val, const double* dinv,
            const double* b, double* x, const int* order, int norder, bool reverse)
{
  const int cnt = order ? norder : n;
  for (int t = 0; t < cnt; ++t) {
    const int tt = reverse ? cnt - 1 - t : t;
    const int i = order ? order[tt] : tt;
    double r[3] = {b[3 * (size_t)i], b[3 * (size_t)i + 1], b[3 * (size_t)i + 2]};
    for (int p = IA[i]; p < IA[i + 1]; ++p) {
      const int j = JA[p];
      if (j == i) continue;
      const double* a = val + 9 * (size_t)p;
      const double* xj = x + 3 * (size_t)j;
      r[0] -= a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2];
      r[1] -= a[3] * xj[0] + a[4] * xj[1] + a[5] * xj[2];
      r[2] -= a[6] * xj[0] + a[7] * xj[1] + a[8] * xj[2];
    }
    const double* d = dinv + 9 * (size_t)i;
    double* xi = x + 3 * (size_t)i;
    xi[0] = d[0] * r[0] + d[1] * r[1] + d[2] * r[2];
    xi[1] = d[3] * r[0] + d[4] * r[1] + d[5] * r[2];
    xi[2] = d[6] * r[0] + d[7] * r[1] + d[8] * r[2];
  }
}

// fasp_dbsr_Linfinity_dcsr [EXT]: one scalar per block = its infinity norm
void condense_linf(int n, const int* IA, const double* val, std::vector<double>& pp)
{
  pp.resize(IA[n]);
  for (int p = 0; p < IA[n]; ++p) {
    const double* a = val + 9 * (size_t)p;
    double m = 0.0;
    for (int r = 0; r < 3; ++r) m = std::max(m, std::fabs(a[3 * r]) + std::fabs(a[3 * r + 1]) + std::fabs(a[3 * r + 2]));
    pp[p] = m;
  }
}

// aggregation_vmb [EXT] (Vanek-Mandel-Brezina greedy aggregation, sequential, natural order)
int vmb_aggregate(int n, const int* IA, const int* JA, const double* pp, double strong, int max_agg,
                  std::vector<int>& vert)
{
  const int UNPT = -1, G0PT = -2;
  std::vector<double> diag(n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int p = IA[i]; p < IA[i + 1]; ++p) if (JA[p] == i) diag[i] = pp[p];
  const double s2 = strong * strong;
  // strongly coupled neighbourhoods (diagonal always kept)
  std::vector<int> NIA(n + 1, 0), NJA;
  NJA.reserve(IA[n]);
  for (int i = 0; i < n; ++i) {
    for (int p = IA[i]; p < IA[i + 1]; ++p) {
      const int j = JA[p];
      if (j == i || pp[p] * pp[p] >= s2 * std::fabs(diag[i] * diag[j])) NJA.push_back(j);
    }
    NIA[i + 1] = (int)NJA.size();
  }
  vert.assign(n, G0PT);
  int nagg = 0, left = n;
  // Step 1: nodes whose whole strong neighbourhood is free seed an aggregate
  for (int i = 0; i < n; ++i) {
    if (IA[i + 1] - IA[i] == 1) { vert[i] = UNPT; --left; continue; }
    bool subset = true;
    for (int p = NIA[i]; p < NIA[i + 1]; ++p) if (vert[NJA[p]] >= UNPT) { subset = false; break; }
    if (!subset) continue;
    int count = 0;
    vert[i] = nagg; --left; ++count;
    for (int p = NIA[i]; p < NIA[i + 1]; ++p)
      if (NJA[p] != i && count < max_agg) { vert[NJA[p]] = nagg; --left; ++count; }
    ++nagg;
  }
  // Step 2: leftovers join a neighbouring step-1 aggregate that is not full
  std::vector<int> tempC(vert), num(nagg, 0);
  for (int i = 0; i < n; ++i) if (vert[i] >= 0) num[vert[i]]++;
  for (int i = 0; i < n; ++i) {
    if (vert[i] >= UNPT) continue;
    for (int p = NIA[i]; p < NIA[i + 1]; ++p) {
      const int a = tempC[NJA[p]];
      if (a > UNPT && num[a] < max_agg) { vert[i] = a; --left; num[a]++; break; }
    }
  }
  // Step 3: the rest form new aggregates with their free strong neighbours
  while (left > 0) {
    for (int i = 0; i < n; ++i) {
      if (vert[i] >= UNPT) continue;
      int count = 0;
      vert[i] = nagg; --left; ++count;
      for (int p = NIA[i]; p < NIA[i + 1]; ++p)
        if (NJA[p] != i && vert[NJA[p]] < UNPT && count < max_agg) { vert[NJA[p]] = nagg; --left; ++count; }
      ++nagg;
    }
  }
  return nagg;
}

// A_c = P^T A P with boolean block P (fasp_blas_dbsr_rap specialised), sorted columns
void rap_agg(const BSR& A, const std::vector<int>& agg, int nagg, BSR& C)
{
  std::vector<std::vector<int>> members(nagg);
  for (int i = 0; i < A.ROW; ++i) if (agg[i] >= 0) members[agg[i]].push_back(i);
  C.ROW = C.COL = nagg;
  C.IA.assign(nagg + 1, 0);
  C.JA.clear(); C.val.clear();
  std::vector<int> mark(nagg, -1);
  for (int I = 0; I < nagg; ++I) {
    std::vector<int> cols;
    for (int i : members[I])
      for (int p = A.IA[i]; p < A.IA[i + 1]; ++p) {
        const int J = agg[A.JA[p]];
        if (J >= 0 && mark[J] != I) { mark[J] = I; cols.push_back(J); }
      }
    std::sort(cols.begin(), cols.end());
    const int base = (int)C.JA.size();
    C.JA.insert(C.JA.end(), cols.begin(), cols.end());
    C.val.resize(9 * C.JA.size(), 0.0);
    for (int i : members[I])
      for (int p = A.IA[i]; p < A.IA[i + 1]; ++p) {
        const int J = agg[A.JA[p]];
        if (J < 0) continue;
        const int slot = base + (int)(std::lower_bound(cols.begin(), cols.end(), J) - cols.begin());
        for (int t = 0; t < 9; ++t) C.val[9 * (size_t)slot + t] += A.val[9 * (size_t)p + t];
      }
    C.IA[I + 1] = (int)C.JA.size();
  }
  C.NNZ = (int)C.JA.size();
}

// plain restarted GMRES on a small system given as BSR (fasp_coarse_itsolver [EXT]:
// GMRES(25), relative residual tol, bounded iteration count), zero initial guess allowed
void coarse_gmres(const BSR& A, const double* b, double* x, double tol, int maxit, int restart)
{
  const int n = 3 * A.ROW;
  std::vector<std::vector<double>> V(restart + 1, std::vector<double>(n));
  std::vector<double> H((restart + 1) * restart, 0.0), cs(restart), sn(restart), g(restart + 1), w(n);
  double bn = 0.0;
  for (int i = 0; i < n; ++i) bn += b[i] * b[i];
  bn = std::sqrt(bn);
  if (bn == 0.0) { std::fill(x, x + n, 0.0); return; }
  int iter = 0;
  while (iter < maxit) {
    bsr_spmv(A.ROW, A.IA.data(), A.JA.data(), A.val.data(), 1.0, x, 0.0, w.data());
    double rn = 0.0;
    for (int i = 0; i < n; ++i) { V[0][i] = b[i] - w[i]; rn += V[0][i] * V[0][i]; }
    rn = std::sqrt(rn);
    if (rn <= tol * bn) return;
    for (int i = 0; i < n; ++i) V[0][i] /= rn;
    std::fill(g.begin(), g.end(), 0.0); g[0] = rn;
    int k = 0;
    for (; k < restart && iter < maxit; ++k, ++iter) {
      bsr_spmv(A.ROW, A.IA.data(), A.JA.data(), A.val.data(), 1.0, V[k].data(), 0.0, V[k + 1].data());
      for (int j = 0; j <= k; ++j) {
        double h = 0.0;
        for (int i = 0; i < n; ++i) h += V[j][i] * V[k + 1][i];
        H[j * restart + k] = h;
        for (int i = 0; i < n; ++i) V[k + 1][i] -= h * V[j][i];
      }
      double hn = 0.0;
      for (int i = 0; i < n; ++i) hn += V[k + 1][i] * V[k + 1][i];
      hn = std::sqrt(hn);
      H[(k + 1) * restart + k] = hn;
      if (hn != 0.0) for (int i = 0; i < n; ++i) V[k + 1][i] /= hn;
      for (int j = 0; j < k; ++j) {
        const double t = H[j * restart + k];
        H[j * restart + k] = cs[j] * t + sn[j] * H[(j + 1) * restart + k];
        H[(j + 1) * restart + k] = -sn[j] * t + cs[j] * H[(j + 1) * restart + k];
      }
      double gam = std::hypot(H[k * restart + k], H[(k + 1) * restart + k]);
      if (gam == 0.0) gam = 1e-20;
      cs[k] = H[k * restart + k] / gam; sn[k] = H[(k + 1) * restart + k] / gam;
      H[k * restart + k] = gam;
      g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
      if (std::fabs(g[k + 1]) <= tol * bn) { ++k; ++iter; break; }
    }
    std::vector<double> y(k);
    for (int i = k - 1; i >= 0; --i) {
      double t = g[i];
      for (int j = i + 1; j < k; ++j) t -= H[i * restart + j] * y[j];
      y[i] = t / H[i * restart + i];
    }
    for (int j = 0; j < k; ++j) for (int i = 0; i < n; ++i) x[i] += y[j] * V[j][i];
  }
}

// dense LU with partial pivoting (mirror of the CUDA path's coarsest-level solve)
struct DenseLU {
  int n = 0; std::vector<double> a; std::vector<int> piv;
  void factor(const BSR& A)
  {
    n = 3 * A.ROW; a.assign((size_t)n * n, 0.0); piv.resize(n);
    for (int i = 0; i < A.ROW; ++i)
      for (int p = A.IA[i]; p < A.IA[i + 1]; ++p)
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
          a[(size_t)(3 * i + r) * n + 3 * A.JA[p] + c] = A.val[9 * (size_t)p + 3 * r + c];
    for (int k = 0; k < n; ++k) {
      int pv = k; double mx = std::fabs(a[(size_t)k * n + k]);
      for (int i = k + 1; i < n; ++i) if (std::fabs(a[(size_t)i * n + k]) > mx) { mx = std::fabs(a[(size_t)i * n + k]); pv = i; }
      piv[k] = pv;
      if (pv != k) for (int j = 0; j < n; ++j) std::swap(a[(size_t)k * n + j], a[(size_t)pv * n + j]);
      const double d = a[(size_t)k * n + k];
      for (int i = k + 1; i < n; ++i) {
        const double l = a[(size_t)i * n + k] / d;
        a[(size_t)i * n + k] = l;
        if (l != 0.0) for (int j = k + 1; j < n; ++j) a[(size_t)i * n + j] -= l * a[(size_t)k * n + j];
      }
    }
  }
  void solve(const double* b, double* x) const
  {
    std::vector<double> y(b, b + n);
    for (int k = 0; k < n; ++k) if (piv[k] != k) std::swap(y[k], y[piv[k]]);
    for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) y[i] -= a[(size_t)i * n + k] * y[k];
    for (int i = n - 1; i >= 0; --i) { double t = y[i]; for (int j = i + 1; j < n; ++j) t -= a[(size_t)i * n + j] * y[j]; y[i] = t / a[(size_t)i * n + i]; }
    std::copy(y.begin(), y.end(), x);
  }
};

struct Level {
  BSR A;
  std::vector<double> dinv, x, b, w;
  std::vector<int> agg; int nagg = 0;
  std::vector<int> order;     // GS row order (empty: natural)
};

} // namespace

struct orc_amg {
  std::vector<Level> lv;
  orc_solver_param p;
  DenseLU lu;
};

extern "C" {

void orc_default_param(orc_solver_param* p)
{
  // benchmarks/pnp_exact_solution/bsr.dat:25-28,57-60,75-77,97-99
  p->tol = 1e-6; p->maxit = 100; p->restart = 30;
  p->max_levels = 20; p->coarse_dof = 100;
  p->strong_coupled = 0.08; p->max_aggregation = 20; p->tentative_smooth = 0.67;
  p->amg_tol = 1e-6; p->presmooth = 1; p->postsmooth = 1; p->print_level = 0;
  p->coarse_dense = 0; p->ortho_cgs = 0; p->cycle_type = 1; p->kcycle_levels = 3;
}

void orc_bsr_spmv(int n, const int* IA, const int* JA, const double* val, double alpha,
                  const double* x, double beta, double* y)
{ bsr_spmv(n, IA, JA, val, alpha, x, beta, y); }

void orc_bsr_diaginv(int n, const int* IA, const int* JA, const double* val, double* dinv)
{ bsr_diaginv(n, IA, JA, val, dinv); }

void orc_bsr_gs(int n, const int* IA, const int* JA, const double* val, const double* dinv,
                const double* b, double* x, const int* order, int norder)
{ bsr_gs(n, IA, JA, val, dinv, b, x, order, norder, false); }

int orc_vmb_aggregate(int n, const int* IA, const int* JA, const double* val, double strong,
                      int max_agg, int* agg)
{
  std::vector<double> pp; std::vector<int> v;
  condense_linf(n, IA, val, pp);
  const int na = vmb_aggregate(n, IA, JA, pp.data(), strong, max_agg, v);
  std::copy(v.begin(), v.end(), agg);
  return na;
}

int orc_rap_agg_count(int n, const int* IA, const int* JA, const int* agg, int nagg, int* cIA)
{
  BSR A; A.ROW = A.COL = n; A.IA.assign(IA, IA + n + 1); A.JA.assign(JA, JA + IA[n]);
  A.val.assign(9 * (size_t)IA[n], 0.0);
  BSR C; rap_agg(A, std::vector<int>(agg, agg + n), nagg, C);
  std::copy(C.IA.begin(), C.IA.end(), cIA);
  return C.NNZ;
}

void orc_rap_agg_fill(int n, const int* IA, const int* JA, const double* val, const int* agg,
                      int nagg, const int* cIA, int* cJA, double* cval)
{
  (void)cIA;
  BSR A; A.ROW = A.COL = n; A.IA.assign(IA, IA + n + 1); A.JA.assign(JA, JA + IA[n]);
  A.val.assign(val, val + 9 * (size_t)IA[n]);
  BSR C; rap_agg(A, std::vector<int>(agg, agg + n), nagg, C);
  std::copy(C.JA.begin(), C.JA.end(), cJA);
  std::copy(C.val.begin(), C.val.end(), cval);
}

orc_amg* orc_amg_setup(int n, const int* IA, const int* JA, const double* val,
                       const orc_solver_param* p, int ext_nlevels, const int* const* ext_agg,
                       const int* ext_nagg, const int* const* ext_order)
{
  orc_amg* h = new orc_amg;
  h->p = *p;
  h->lv.emplace_back();
  {
    Level& L = h->lv[0];
    L.A.ROW = L.A.COL = n; L.A.NNZ = IA[n];
    L.A.IA.assign(IA, IA + n + 1); L.A.JA.assign(JA, JA + IA[n]); L.A.val.assign(val, val + 9 * (size_t)IA[n]);
  }
  const int min_cdof = std::max(p->coarse_dof, 50);
  const int max_levels = ext_nlevels > 0 ? ext_nlevels : p->max_levels;
  int lvl = 0;
  while (true) {
    Level& L = h->lv[lvl];
    L.dinv.resize(9 * (size_t)L.A.ROW);
    bsr_diaginv(L.A.ROW, L.A.IA.data(), L.A.JA.data(), L.A.val.data(), L.dinv.data());
    if (ext_order && ext_order[lvl]) L.order.assign(ext_order[lvl], ext_order[lvl] + L.A.ROW);
    if (ext_nlevels > 0) { if (lvl >= ext_nlevels - 1) break; }
    else if (!(L.A.ROW > min_cdof && lvl < max_levels - 1)) break;
    int nagg;
    if (ext_nlevels > 0) {
      L.agg.assign(ext_agg[lvl], ext_agg[lvl] + L.A.ROW); nagg = ext_nagg[lvl];
    } else {
      std::vector<double> pp;
      condense_linf(L.A.ROW, L.A.IA.data(), L.A.val.data(), pp);
      const double sc = p->tentative_smooth > 1e-20 ? p->strong_coupled * std::pow(0.5, lvl) : p->strong_coupled;
      nagg = vmb_aggregate(L.A.ROW, L.A.IA.data(), L.A.JA.data(), pp.data(), sc, p->max_aggregation, L.agg);
      if (nagg < 20 /* MIN_CDOF */ || nagg >= L.A.ROW) { L.agg.clear(); break; }
    }
    L.nagg = nagg;
    h->lv.emplace_back();
    rap_agg(h->lv[lvl].A, h->lv[lvl].agg, nagg, h->lv[lvl + 1].A);
    ++lvl;
  }
  for (auto& L : h->lv) { L.x.assign(3 * (size_t)L.A.ROW, 0.0); L.b = L.x; L.w = L.x; }
  if (p->coarse_dense) h->lu.factor(h->lv.back().A);
  if (p->print_level > 0)
    for (size_t l = 0; l < h->lv.size(); ++l)
      printf("  [oracle amg] level %zu: rows %d nnzb %d\n", l, h->lv[l].A.ROW, h->lv[l].A.NNZ);
  return h;
}

int orc_amg_nlevels(const orc_amg* h) { return (int)h->lv.size(); }

void orc_amg_level_info(const orc_amg* h, int l, int* nrow, int* nnzb, int* nagg)
{ *nrow = h->lv[l].A.ROW; *nnzb = h->lv[l].A.NNZ; *nagg = h->lv[l].nagg; }

void orc_amg_level_matrix(const orc_amg* h, int l, int* IA, int* JA, double* val)
{
  const BSR& A = h->lv[l].A;
  std::copy(A.IA.begin(), A.IA.end(), IA); std::copy(A.JA.begin(), A.JA.end(), JA);
  std::copy(A.val.begin(), A.val.end(), val);
}

void orc_amg_level_agg(const orc_amg* h, int l, int* agg)
{ std::copy(h->lv[l].agg.begin(), h->lv[l].agg.end(), agg); }

void orc_amg_free(orc_amg* h) { delete h; }

// fasp_precond_dbsr_amg + fasp_solver_mgcycle_bsr [EXT]: one cycle from a zero guess.
// cycle_type 1 = V-cycle (bsr.dat). Any other value = K-cycle (FASP's nonlinear AMLI cycle in
// Notay's two-step GCR form) on coarse levels 1..kcycle_levels — the mirror of the CUDA path.
static void solve_level(orc_amg* h, int l, const std::vector<double>& b, std::vector<double>& x);

static void mg_cycle(orc_amg* h, int l, const std::vector<double>& b, std::vector<double>& x)
{
  const orc_solver_param& p = h->p;
  Level& L = h->lv[l];
  Level& C = h->lv[l + 1];
  auto smooth = [&](bool reverse) {
    bsr_gs(L.A.ROW, L.A.IA.data(), L.A.JA.data(), L.A.val.data(), L.dinv.data(), b.data(), x.data(),
           L.order.empty() ? nullptr : L.order.data(), (int)L.order.size(), reverse);
  };
  std::fill(x.begin(), x.end(), 0.0);
  for (int s = 0; s < p.presmooth; ++s) smooth(false);
  L.w = b;
  bsr_spmv(L.A.ROW, L.A.IA.data(), L.A.JA.data(), L.A.val.data(), -1.0, x.data(), 1.0, L.w.data());
  std::fill(C.b.begin(), C.b.end(), 0.0);
  for (int i = 0; i < L.A.ROW; ++i)
    if (L.agg[i] >= 0) for (int k = 0; k < 3; ++k) C.b[3 * (size_t)L.agg[i] + k] += L.w[3 * (size_t)i + k];
  solve_level(h, l + 1, C.b, C.x);
  for (int i = 0; i < L.A.ROW; ++i)
    if (L.agg[i] >= 0) for (int k = 0; k < 3; ++k) x[3 * (size_t)i + k] += C.x[3 * (size_t)L.agg[i] + k];
  for (int s = 0; s < p.postsmooth; ++s) smooth(true);
}

static void solve_level(orc_amg* h, int l, const std::vector<double>& b, std::vector<double>& x)
{
  const orc_solver_param& p = h->p;
  const int nl = (int)h->lv.size();
  Level& L = h->lv[l];
  if (l == nl - 1) {
    if (p.coarse_dense) h->lu.solve(b.data(), x.data());
    else {
      const int n = 3 * L.A.ROW;
      std::fill(x.begin(), x.end(), 0.0);
      coarse_gmres(L.A, b.data(), x.data(), p.amg_tol, std::min(n * n, 200), 25);
    }
    return;
  }
  const bool kcyc = p.cycle_type != 1 && p.cycle_type != 0 && l <= p.kcycle_levels;
  if (!kcyc) { mg_cycle(h, l, b, x); return; }
  const size_t m = 3 * (size_t)L.A.ROW;
  std::vector<double> c(m), v(m), rt(m), d(m), w(m);
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& bb) { double t = 0; for (size_t i = 0; i < m; ++i) t += a[i] * bb[i]; return t; };
  mg_cycle(h, l, b, c);
  bsr_spmv(L.A.ROW, L.A.IA.data(), L.A.JA.data(), L.A.val.data(), 1.0, c.data(), 0.0, v.data());
  const double al1 = dot(v, b), r1 = dot(v, v);
  const double a1 = r1 != 0.0 ? al1 / r1 : 0.0;
  for (size_t i = 0; i < m; ++i) rt[i] = b[i] - a1 * v[i];
  mg_cycle(h, l, rt, d);
  bsr_spmv(L.A.ROW, L.A.IA.data(), L.A.JA.data(), L.A.val.data(), 1.0, d.data(), 0.0, w.data());
  const double g = dot(w, v), be = dot(w, w), al2 = dot(w, rt);
  const double r2 = r1 != 0.0 ? be - g * g / r1 : 0.0;
  double k1 = a1, k2 = 0.0;
  if (r2 > 0.0 && r1 != 0.0) { k2 = al2 / r2; k1 = a1 - g * k2 / r1; }
  for (size_t i = 0; i < m; ++i) x[i] = k1 * c[i] + k2 * d[i];
}

void orc_amg_vcycle(orc_amg* h, const double* r, double* z)
{
  const int nl = (int)h->lv.size();
  const orc_solver_param& p = h->p;
  Level& F = h->lv[0];
  std::copy(r, r + 3 * (size_t)F.A.ROW, F.b.begin());
  if (nl == 1) {   // degenerate: single level, smoothing only
    std::fill(F.x.begin(), F.x.end(), 0.0);
    for (int s = 0; s < p.presmooth; ++s)
      bsr_gs(F.A.ROW, F.A.IA.data(), F.A.JA.data(), F.A.val.data(), F.dinv.data(), F.b.data(), F.x.data(),
             F.order.empty() ? nullptr : F.order.data(), (int)F.order.size(), false);
  } else {
    std::vector<double> bb(F.b);
    mg_cycle(h, 0, bb, F.x);
  }
  std::copy(F.x.begin(), F.x.end(), z);
}

} // extern "C"

namespace orc {
// fasp_solver_dbsr_pvfgmres restated [EXT]; pc = the preconditioner z = M^{-1} r (null: none)
int pvfgmres_impl(int nrow, const int* IA, const int* JA, const double* val, const double* b,
                  double* x, const std::function<void(const double*, double*)>* pc, const orc_solver_param* prm, double* hist, int hist_len)
{
  const int n = 3 * nrow;
  const int restart_max = std::max(1, std::min(prm->restart, prm->maxit)), restart_min = 3, dstep = 3;
  const double cr_max = 0.99, cr_min = 0.174, epsmac = 1e-20;
  const int MaxIt = prm->maxit;
  int Restart = restart_max, iter = 0;
  double cr = 1.0, r_norm, r_norm_old;
  std::vector<std::vector<double>> pv(restart_max + 1, std::vector<double>(n)), zv(restart_max, std::vector<double>(n));
  std::vector<std::vector<double>> hh(restart_max + 1, std::vector<double>(restart_max, 0.0));
  std::vector<double> c(restart_max), s(restart_max), rs(restart_max + 1), r(n);
  auto nrm2 = [&](const std::vector<double>& v) { double t = 0; for (int i = 0; i < n; ++i) t += v[i] * v[i]; return std::sqrt(t); };
  auto dotp = [&](const std::vector<double>& a, const std::vector<double>& bb) { double t = 0; for (int i = 0; i < n; ++i) t += a[i] * bb[i]; return t; };
  double b_norm = 0; for (int i = 0; i < n; ++i) b_norm += b[i] * b[i]; b_norm = std::sqrt(b_norm);
  // r = b - A x
  std::copy(b, b + n, pv[0].begin());
  bsr_spmv(nrow, IA, JA, val, -1.0, x, 1.0, pv[0].data());
  r_norm = nrm2(pv[0]);
  const double den = b_norm > 0.0 ? b_norm : r_norm;
  const double epsilon = prm->tol * den;
  if (r_norm <= epsilon) return 0;
  while (iter < MaxIt) {
    rs[0] = r_norm; r_norm_old = r_norm;
    if (r_norm == 0.0) return iter;
    // adaptive restart length
    if (cr > cr_max || iter == 0) Restart = restart_max;
    else if (cr < cr_min) { /* keep */ }
    else { if (Restart - dstep > restart_min) Restart -= dstep; else Restart = restart_max; }
    double t = 1.0 / r_norm;
    for (int i = 0; i < n; ++i) pv[0][i] *= t;
    int i = 0;
    while (i < Restart && iter < MaxIt) {
      ++i; ++iter;
      if (pc) (*pc)(pv[i - 1].data(), zv[i - 1].data()); else zv[i - 1] = pv[i - 1];
      bsr_spmv(nrow, IA, JA, val, 1.0, zv[i - 1].data(), 0.0, pv[i].data());
      if (!prm->ortho_cgs) {
        for (int j = 0; j < i; ++j) {           // modified Gram-Schmidt
          hh[j][i - 1] = dotp(pv[j], pv[i]);
          for (int k = 0; k < n; ++k) pv[i][k] -= hh[j][i - 1] * pv[j][k];
        }
      } else {                                   // classical GS + DGKS re-orthogonalisation
        const double before = nrm2(pv[i]);
        std::vector<double> hcol(i);
        for (int j = 0; j < i; ++j) hcol[j] = dotp(pv[j], pv[i]);
        for (int j = 0; j < i; ++j) for (int k = 0; k < n; ++k) pv[i][k] -= hcol[j] * pv[j][k];
        if (nrm2(pv[i]) < 0.1 * before) {
          std::vector<double> h2(i);
          for (int j = 0; j < i; ++j) h2[j] = dotp(pv[j], pv[i]);
          for (int j = 0; j < i; ++j) { for (int k = 0; k < n; ++k) pv[i][k] -= h2[j] * pv[j][k]; hcol[j] += h2[j]; }
        }
        for (int j = 0; j < i; ++j) hh[j][i - 1] = hcol[j];
      }
      t = nrm2(pv[i]);
      hh[i][i - 1] = t;
      if (t != 0.0) { t = 1.0 / t; for (int k = 0; k < n; ++k) pv[i][k] *= t; }
      for (int j = 1; j < i; ++j) {
        t = hh[j - 1][i - 1];
        hh[j - 1][i - 1] = s[j - 1] * hh[j][i - 1] + c[j - 1] * t;
        hh[j][i - 1] = -s[j - 1] * t + c[j - 1] * hh[j][i - 1];
      }
      t = hh[i][i - 1] * hh[i][i - 1] + hh[i - 1][i - 1] * hh[i - 1][i - 1];
      double gamma = std::sqrt(t);
      if (gamma == 0.0) gamma = epsmac;
      c[i - 1] = hh[i - 1][i - 1] / gamma;
      s[i - 1] = hh[i][i - 1] / gamma;
      rs[i] = -s[i - 1] * rs[i - 1];
      rs[i - 1] = c[i - 1] * rs[i - 1];
      hh[i - 1][i - 1] = s[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
      r_norm = std::fabs(rs[i]);
      if (hist && iter - 1 < hist_len) hist[iter - 1] = r_norm / den;
      if (prm->print_level > 1) printf("  [oracle fgmres] it %3d  relres %.6e\n", iter, r_norm / den);
      if (r_norm <= epsilon) break;
    }
    // solve the triangular system and update x with the preconditioned basis
    rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
    for (int k = i - 2; k >= 0; --k) {
      t = 0.0;
      for (int j = k + 1; j < i; ++j) t -= hh[k][j] * rs[j];
      t += rs[k];
      rs[k] = t / hh[k][k];
    }
    for (int k = 0; k < n; ++k) r[k] = zv[i - 1][k] * rs[i - 1];
    for (int j = i - 2; j >= 0; --j) for (int k = 0; k < n; ++k) r[k] += rs[j] * zv[j][k];
    for (int k = 0; k < n; ++k) x[k] += r[k];
    if (r_norm <= epsilon) {
      std::copy(b, b + n, r.begin());
      bsr_spmv(nrow, IA, JA, val, -1.0, x, 1.0, r.data());
      r_norm = nrm2(r);
      if (r_norm <= epsilon) break;
      pv[0] = r; i = 0;                          // false convergence: restart from the true residual
    }
    // residual vector for the next cycle (recurrence on the Arnoldi basis)
    for (int j = i; j > 0; --j) { rs[j - 1] = -s[j - 1] * rs[j]; rs[j] = c[j - 1] * rs[j]; }
    if (i) for (int k = 0; k < n; ++k) pv[i][k] *= rs[i];
    for (int j = i - 1; j > 0; --j) for (int k = 0; k < n; ++k) pv[i][k] += rs[j] * pv[j][k];
    if (i) { for (int k = 0; k < n; ++k) pv[0][k] = rs[0] * pv[0][k] + pv[i][k]; }
    cr = r_norm / r_norm_old;
  }
  if (iter >= MaxIt && r_norm > epsilon) return -1;   // ERROR_SOLVER_MAXIT
  return iter;
}
} // namespace orc

extern "C" {

int orc_pvfgmres(int nrow, const int* IA, const int* JA, const double* val, const double* b,
                 double* x, orc_amg* pc, const orc_solver_param* prm, double* hist, int hist_len)
{
  if (!pc) return orc::pvfgmres_impl(nrow, IA, JA, val, b, x, nullptr, prm, hist, hist_len);
  std::function<void(const double*, double*)> f = [pc](const double* r, double* z) { orc_amg_vcycle(pc, r, z); };
  return orc::pvfgmres_impl(nrow, IA, JA, val, b, x, &f, prm, hist, hist_len);
}

int orc_solver_dbsr_krylov_amg(int nrow, const int* IA, const int* JA, const double* val,
                               const double* b, double* x, const orc_solver_param* p)
{
  orc_amg* h = orc_amg_setup(nrow, IA, JA, val, p, 0, nullptr, nullptr, nullptr);
  const int its = orc_pvfgmres(nrow, IA, JA, val, b, x, h, p, nullptr, 0);
  orc_amg_free(h);
  return its;
}

}
